"""torchrun check of the sharded symmetric build: identical to the single-GPU result, and timing against the row-block
(one-sided) sharding.  torchrun --nproc-per-node G scripts/shard_sym_check.py [n ...]"""
import os, sys, time, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
from semibin_b200 import synth, dist as sdist
from semibin_b200.neighbors import knn
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ.get('LOCAL_RANK', rank))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
sizes = [int(a) for a in sys.argv[1:]] or [41000, 1000000]
for n in sizes:
    k = 50 if n < 100000 else 200
    X = torch.from_numpy(synth.embeddings(n, seed=77)).cuda()
    bufs = sdist.PeerResult(n, k, torch.float32, X.device)
    shard = sdist.SymShard(n, X.shape[1], k, torch.float32, X.device)
    for it in range(3):
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        d, i, st = sdist.knn_sharded_sym(X, k, bufs, shard, return_stats=True)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
    d = d.clone(); i = i.clone()
    for it in range(3):
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        d1, i1, st1 = sdist.knn_sharded_p2p(X, k, bufs, return_stats=True)
        torch.cuda.synchronize(); ms1 = (time.perf_counter() - t0) * 1e3
    same = bool(torch.equal(d, d1) and torch.equal(i, i1))
    ref_ok = None
    if n <= 100000:
        d0, i0 = knn(X, k, algo='exact')
        ref_ok = bool(torch.equal(d0, d) and torch.equal(i0, i))
    keep = lambda s: {q: (round(v, 2) if isinstance(v, float) else v) for q, v in s.items() if q.startswith('ms_') or q in ('n_fallback_rows', 'n_candidates', 'symmetric', 'n_filter_launches')}
    print(json.dumps({'rank': rank, 'n': n, 'sym_ms': round(ms, 1), 'rowblock_ms': round(ms1, 1), 'sym_equals_rowblock': same, 'sym_equals_exact': ref_ok,
                      'sym': keep(st), 'rowblock': keep(st1)}), flush=True)
    del bufs, shard
dist.destroy_process_group()
