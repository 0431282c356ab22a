#!/bin/bash
# usage: scripts/gpu_profile.sh <tag>   (run under gpurun)  -> gpurun_out/<tag>_*
# plain bench first; then (one ncu session per gpurun call) the launch list and the --set full captures of the filter
# and of the warp rerank, each after the same command line has exited 0 without ncu
tag=${1:-prof}
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || exit 1
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_launches.log 2>&1
$CMD > gpurun_out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_kernel -s 9 -c 1 -o gpurun_out/${tag}_filter -f $CMD > gpurun_out/${tag}_ncu_filter.log 2>&1
$CMD > gpurun_out/${tag}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'rerank_warp|tc_transpose|tc_sort_tiles' -s 7 -c 4 -o gpurun_out/${tag}_rerank -f $CMD > gpurun_out/${tag}_ncu_rerank.log 2>&1
ls -la gpurun_out/${tag}_*
