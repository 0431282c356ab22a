#!/bin/bash
# usage: scripts/gpu_ncu_filter.sh <tag>  -> gpurun_out/<tag>_filter.ncu-rep (one mid-run filter launch, full set + source)
tag=${1:-f}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_kernel -s 9 -c 1 -o gpurun_out/${tag}_filter -f $CMD > gpurun_out/${tag}_ncu_filter.log 2>&1
tail -3 gpurun_out/${tag}_ncu_filter.log
