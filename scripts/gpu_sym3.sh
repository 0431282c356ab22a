#!/bin/bash
tag=${1:-sym3}
timeout 300 python scripts/dbg_sym.py 200000 1000000 > gpurun_out/${tag}_sym.log 2>&1; grep '"n"' gpurun_out/${tag}_sym.log | cut -c1-420
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_launches.log 2>&1
