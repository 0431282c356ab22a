#!/bin/bash
tag=${1:-dbgs}
timeout 300 python scripts/dbg_sym.py 20000 200000 1000000 > gpurun_out/${tag}_sym.log 2>&1
cat gpurun_out/${tag}_sym.log | cut -c1-1500
