#!/bin/bash
tag=${1:-sym5}
timeout 900 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -14 gpurun_out/${tag}_pytest.log
