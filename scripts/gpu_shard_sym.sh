#!/bin/bash
# sharded symmetric build on N GPUs: the multi-GPU tests, the 1M check against the row-block sharding, one bench line
n=${1:-4}; tag=${2:-r02shard}
python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/${tag}_pytest_${n}gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_${n}gpu.log | cut -c1-600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 scripts/shard_sym_check.py 1000000 > gpurun_out/${tag}_check_${n}gpu.jsonl 2> gpurun_out/${tag}_check_${n}gpu.err
echo "check rc=$?"; head -2 gpurun_out/${tag}_check_${n}gpu.jsonl | cut -c1-700; tail -2 gpurun_out/${tag}_check_${n}gpu.err | cut -c1-300
timeout 420 bash scripts/gpu_scale_n.sh $n $tag
