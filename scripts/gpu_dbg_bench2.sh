#!/bin/bash
tag=${1:-dbgb}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench2.json 2> gpurun_out/${tag}_bench2.err
tail -3 gpurun_out/${tag}_bench2.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/dbgb_bench2.json').read().strip().splitlines()[-1])
print(j['ms_per_step'], json.dumps(j['e2e'], indent=0)[:3000])
PY
