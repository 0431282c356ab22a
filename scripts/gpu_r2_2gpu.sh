#!/bin/bash
# round-2 two-GPU call: single-process multi-device drop-ins, P2P mirror stores, torchrun bench with both gather modes
tag=${1:-r02g2}
set -x
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
python -m pytest tests/test_gpu_scale.py -m gpu -q -x -k "devices or peer" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
for mode in p2p nccl; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --gather $mode > gpurun_out/${tag}_bench_${mode}.json 2> gpurun_out/${tag}_bench_${mode}.err
tail -2 gpurun_out/${tag}_bench_${mode}.err
done
python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err
for f in gpurun_out/${tag}_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:j[k] for k in ('value','ms_per_step','stages_ms')}, j['config'].get('gather'), {k:v for k,v in j.get('e2e',{}).items() if k not in('api','spmd_host_buffers')}, j.get('parity') or j.get('cpu_baseline'))
except Exception as e: print('ERR',e)
PY
done
