#!/bin/bash
# realism runs of bench.py: the embedding a shipped model produces (scratch/emb_global_1000000.npy, made by
# scripts/make_real_embeddings.py in the build container) and the large-genome mix
tag=${1:-real}
if [ -f scratch/emb_global_1000000.npy ]; then
python bench.py --steps 3 --warmup 3 --embedding-file scratch/emb_global_1000000.npy > gpurun_out/${tag}_bench_globalmodel.json 2> gpurun_out/${tag}_bench_globalmodel.err
tail -2 gpurun_out/${tag}_bench_globalmodel.err
fi
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --genome-size 2000 > gpurun_out/${tag}_bench_genome2000.json 2> gpurun_out/${tag}_bench_genome2000.err
for f in gpurun_out/${tag}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k: j[k] for k in ('value', 'ms_per_step', 'stages_ms', 'knn_stats')}, j['e2e']['ms_per_step'], j['roofline']['frac'], j.get('cpu_baseline'))
except Exception as e:
    print('ERR', e)
PY
done
