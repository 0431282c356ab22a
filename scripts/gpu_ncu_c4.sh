#!/bin/bash
tag=${1:-c4l}
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ig_|DeviceRadixSort' -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench_path.py --config C4 > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log | cut -c1-300
