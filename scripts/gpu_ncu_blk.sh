#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/blk_launches.csv python scripts/dbg_blk1.py > gpurun_out/blk_ncu.log 2>&1
tail -2 gpurun_out/blk_ncu.log | cut -c1-300
