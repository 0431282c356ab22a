#!/bin/bash
tag=${1:-dbgs2}
SBK_LIB_VARIANT=probe SBK_SYM_DEBUG=1 timeout 300 python scripts/dbg_sym.py 1000000 > gpurun_out/${tag}_symdbg.log 2>&1
grep -E "sym launch|\"n\"" gpurun_out/${tag}_symdbg.log | cut -c1-400 | head -40
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'tc_kernel<\(int\)6' -s 7 -c 1 -o gpurun_out/${tag}_sym -f python scripts/dbg_sym.py 1000000 > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log | cut -c1-300
