#!/bin/bash
tag=${1:-r02w2}
python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${tag}_smoke.log | cut -c1-300
python -m pytest tests/test_gpu_scale.py -m gpu -q -x -k "devices or peer" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
for mode in p2p nccl; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --gather $mode > gpurun_out/${tag}_bench_${mode}.json 2> gpurun_out/${tag}_bench_${mode}.err
tail -2 gpurun_out/${tag}_bench_${mode}.err
python - gpurun_out/${tag}_bench_${mode}.json <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:j[k] for k in ('value','ms_per_step','stages_ms')}, j['config'].get('gather'), {k:v for k,v in j.get('e2e',{}).items() if k in('value','ms_per_step')}, j['e2e']['spmd_host_buffers']['ms_per_step'], j.get('cpu_baseline',{}).get('parity_rows_bad'))
except Exception as e: print('ERR',e)
PY
done
