#!/bin/bash
# host-buffer paths (block-row schedule + zero-copy scatter): their GPU tests, then the default bench line
tag=${1:-hostpath}
timeout 600 python -m pytest tests/test_gpu_scale.py tests/test_gpu_knn.py -m gpu -x -q --durations=5 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -25 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err
python - gpurun_out/${tag}_bench.json <<'PY'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: j[k] for k in ('value', 'ms_per_step', 'stages_ms')}, {k: v for k, v in j['e2e'].items() if k != 'api'}, j.get('cpu_baseline'))
PY
