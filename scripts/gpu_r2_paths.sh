#!/bin/bash
# round-2 path runs: the BASELINE configs through bench_path.py, realism runs of bench.py
tag=${1:-r02p}
set -x
python bench_path.py --config C2 --cpu-check --sklearn-graphs > gpurun_out/${tag}_C2sklearn.json 2> gpurun_out/${tag}_C2sklearn.err
python bench_path.py --config C3 > gpurun_out/${tag}_C3.json 2> gpurun_out/${tag}_C3.err
python bench_path.py --config C3 --combined > gpurun_out/${tag}_C3combined.json 2> gpurun_out/${tag}_C3combined.err
python bench_path.py --config C4 --cpu-check > gpurun_out/${tag}_C4.json 2> gpurun_out/${tag}_C4.err
python bench_path.py --config C5 > gpurun_out/${tag}_C5.json 2> gpurun_out/${tag}_C5.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --genome-size 2000 > gpurun_out/${tag}_bench_genome2000.json 2> gpurun_out/${tag}_bench_genome2000.err
if [ -f scratch/emb_global_1000000.npy ]; then
python bench.py --steps 3 --warmup 3 --embedding-file scratch/emb_global_1000000.npy > gpurun_out/${tag}_bench_globalmodel.json 2> gpurun_out/${tag}_bench_globalmodel.err
fi
for f in gpurun_out/${tag}_*.json; do echo "== $f"; head -c 2500 $f; echo; done
for f in gpurun_out/${tag}_*.err; do tail -n 3 "$f"; done
