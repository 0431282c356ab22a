#!/bin/bash
tag=${1:-sym4}
timeout 900 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -14 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err
timeout 600 python bench_path.py --config C4 > gpurun_out/${tag}_C4.json 2> gpurun_out/${tag}_C4.err; tail -3 gpurun_out/${tag}_C4.err
timeout 600 python bench_path.py --config C3 > gpurun_out/${tag}_C3.json 2> gpurun_out/${tag}_C3.err; tail -3 gpurun_out/${tag}_C3.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --genome-size 2000 > gpurun_out/${tag}_bench_genome2000.json 2> gpurun_out/${tag}_bench_genome2000.err; tail -3 gpurun_out/${tag}_bench_genome2000.err
for f in gpurun_out/${tag}_bench.json gpurun_out/${tag}_bench_genome2000.json; do python - "$f" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:j[k] for k in ('value','ms_per_step','stages_ms','knn_stats')}, j['e2e']['ms_per_step'], j.get('cpu_baseline'), j['roofline']['frac'])
except Exception as e: print('ERR',e)
PY
done
head -c 1500 gpurun_out/${tag}_C4.json; echo; head -c 1800 gpurun_out/${tag}_C3.json; echo
