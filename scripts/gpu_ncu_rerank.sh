#!/bin/bash
tag=${1:-r}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rerank_kernel -s 2 -c 1 -o gpurun_out/${tag}_rerank -f $CMD > gpurun_out/${tag}_ncu_rerank.log 2>&1
tail -2 gpurun_out/${tag}_ncu_rerank.log
