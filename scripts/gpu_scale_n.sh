#!/bin/bash
# one torchrun bench line on N GPUs (the driver's launch line)
n=${1:-8}; tag=${2:-scale}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
tail -2 gpurun_out/${tag}_bench_${n}gpu.err | cut -c1-300
python - gpurun_out/${tag}_bench_${n}gpu.json <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: j[k] for k in ('value', 'ms_per_step', 'stages_ms')}, j['config'].get('gather'), {k: v for k, v in j.get('e2e', {}).items() if k in ('value', 'ms_per_step')},
          j['e2e']['spmd_host_buffers']['ms_per_step'], (j.get('parity') or {}).get('parity_rows_bad'), j['clocks'])
except Exception as e:
    print('ERR', e)
PY
