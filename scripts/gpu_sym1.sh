#!/bin/bash
tag=${1:-sym1}
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -25 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -5 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --algo tensor_onesided > gpurun_out/${tag}_bench_onesided.json 2> gpurun_out/${tag}_bench_onesided.err; tail -5 gpurun_out/${tag}_bench_onesided.err
for f in gpurun_out/${tag}_bench.json gpurun_out/${tag}_bench_onesided.json; do python - "$f" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:j[k] for k in ('value','ms_per_step','stages_ms','knn_stats')}, j['e2e']['ms_per_step'], j.get('cpu_baseline'))
except Exception as e: print('ERR',e)
PY
done
