#!/bin/bash
# round-2 call B: all GPU tests (new scale + boundary tests), bench through kneighbors_graph, A/B rerank, ncu rerank
tag=${1:-r02b}
set -x
python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -30 gpurun_out/${tag}_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench.err
SBK_LIB_VARIANT=probe SBK_RERANK_BLOCK=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_blockrr.json 2> gpurun_out/${tag}_bench_blockrr.err
SBK_LIB_VARIANT=probe SBK_NO_ROW_ORDER=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_noorder.json 2> gpurun_out/${tag}_bench_noorder.err
SBK_LIB_VARIANT=probe SBK_TC_NH=2 python -m pytest tests/test_gpu_tc.py tests/test_gpu_knn.py -m gpu -q -x > gpurun_out/${tag}_pytest_nh2.log 2>&1; echo "pytest nh2 rc=$?" >> gpurun_out/${tag}_pytest_nh2.log
tail -4 gpurun_out/${tag}_pytest_nh2.log
SBK_LIB_VARIANT=probe SBK_TC_NH=2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_nh2.json 2> gpurun_out/${tag}_bench_nh2.err
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'rerank_warp' -s 1 -c 1 -o gpurun_out/${tag}_rerank -f $CMD > gpurun_out/${tag}_ncu_rerank.log 2>&1
tail -3 gpurun_out/${tag}_ncu_rerank.log
for f in gpurun_out/${tag}_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:j[k] for k in ('value','ms_per_step','stages_ms','knn_stats')}, {k:v for k,v in j.get('e2e',{}).items() if k!='api'}, j.get('cpu_baseline'))
except Exception as e: print('ERR',e)
PY
done
