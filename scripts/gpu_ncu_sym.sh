#!/bin/bash
# ncu capture of one launch of the symmetric filter kernel (200k rows: short launches, same per-tile behaviour)
tag=${1:-ncus}
n=${2:-200000}
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'tc_kernel<\(int\)6' -s 6 -c 1 -o gpurun_out/${tag}_sym -f python scripts/compare_filters.py $n > gpurun_out/${tag}_ncu.log 2>&1
tail -5 gpurun_out/${tag}_ncu.log
ls -la gpurun_out/${tag}_sym*
