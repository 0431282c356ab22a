#!/bin/bash
tag=${1:-ig1}
timeout 600 python -m pytest tests/test_gpu_longread.py tests/test_gpu_boundary.py -m gpu -x -q --durations=4 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -8 gpurun_out/${tag}_pytest.log
timeout 300 python bench_path.py --config C4 > gpurun_out/${tag}_C4.json 2> gpurun_out/${tag}_C4.err; tail -3 gpurun_out/${tag}_C4.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/ig1_C4.json')); print({k:j[k] for k in ('knn_ms','sweep_ms','labels_ms','integrate_ms','integrate_bins','integrate_binned_contigs','total_ms')})
PY
