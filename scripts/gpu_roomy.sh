#!/bin/bash
python scripts/compare_filters.py 1000000 2>&1 | grep '"n"' | cut -c1-420
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --genome-size 2000 > gpurun_out/roomy_g2000.json 2> gpurun_out/roomy_g2000.err; tail -2 gpurun_out/roomy_g2000.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/roomy_g2000.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','stages_ms','knn_stats')}, j['e2e']['ms_per_step'])
PY
