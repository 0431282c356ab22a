#!/bin/bash
# usage: scripts/gpu_ncu_path.sh <tag>   (run under gpurun): ncu --set full of the HBM-bound kernels of the edge-list
# and long-read paths (achieved GB/s against the measured HBM peak)
tag=${1:-path}
set -x
python bench_path.py --config C3 > gpurun_out/${tag}_C3_plain.json 2> gpurun_out/${tag}_C3_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:'edges_|threshold_count' -c 8 -o gpurun_out/${tag}_edges -f \
    python bench_path.py --config C3 > gpurun_out/${tag}_ncu_edges.log 2>&1
python bench_path.py --config C4 > gpurun_out/${tag}_C4_plain.json 2> gpurun_out/${tag}_C4_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:'radius_sweep|label_prop' -c 8 -o gpurun_out/${tag}_longread -f \
    python bench_path.py --config C4 > gpurun_out/${tag}_ncu_longread.log 2>&1
ls -la gpurun_out/${tag}_*
