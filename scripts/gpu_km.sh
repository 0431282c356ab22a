#!/bin/bash
tag=${1:-km1}
timeout 600 python -m pytest tests/test_gpu_kmeans.py tests/test_gpu_boundary.py -m gpu -x -q --durations=5 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -30 gpurun_out/${tag}_pytest.log
