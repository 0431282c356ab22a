#!/bin/bash
tag=${1:-dbgm}
python scripts/dbg_multi_e2e.py > gpurun_out/${tag}_multi.log 2>&1
cat gpurun_out/${tag}_multi.log
