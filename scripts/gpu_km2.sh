#!/bin/bash
tag=${1:-km2}
timeout 600 python bench_path.py --config N4 > gpurun_out/${tag}_N4.json 2> gpurun_out/${tag}_N4.err; tail -3 gpurun_out/${tag}_N4.err; cat gpurun_out/${tag}_N4.json
