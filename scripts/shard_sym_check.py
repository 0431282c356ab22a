"""torchrun check of the sharded symmetric build: identical to the row-block (one-sided) sharding and, for small n, to the
exact path; timing of both.  torchrun --nproc-per-node G scripts/shard_sym_check.py [--starve] [--iters I] [n ...]
--starve: mailboxes of 1024 records, so the overflow vote must send every rank to the row-block build."""
import argparse, os, sys, time, json
import torch, torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from semibin_b200 import synth, dist as sdist, _lib
from semibin_b200.neighbors import knn
ap = argparse.ArgumentParser()
ap.add_argument('sizes', nargs='*', type=int)
ap.add_argument('--starve', action='store_true')
ap.add_argument('--iters', type=int, default=3)
args = ap.parse_args()
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ.get('LOCAL_RANK', rank))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
if args.starve:
    _lib.load().sbk_debug_shard_mail_cap(1024)
keep = lambda s: {q: (round(v, 2) if isinstance(v, float) else v) for q, v in s.items()
                  if q.startswith('ms_') or q in ('n_fallback_rows', 'n_candidates', 'symmetric', 'n_filter_launches', 'sym_shard_fallback')}
for n in (args.sizes or [41000, 1000000]):
    k = 50 if n < 100000 else 200
    X = torch.from_numpy(synth.embeddings(n, seed=77)).cuda()
    bufs = sdist.PeerResult(n, k, torch.float32, X.device)
    shard = sdist.SymShard(n, X.shape[1], k, torch.float32, X.device)
    for it in range(args.iters):
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        d, i, st = sdist.knn_sharded_sym(X, k, bufs, shard, return_stats=True)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
    d = d.clone(); i = i.clone()
    for it in range(args.iters):
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        d1, i1, st1 = sdist.knn_sharded_p2p(X, k, bufs, return_stats=True)
        torch.cuda.synchronize(); ms1 = (time.perf_counter() - t0) * 1e3
    same = bool(torch.equal(d, d1) and torch.equal(i, i1))
    ref_ok = None
    if n <= 100000:
        d0, i0 = knn(X, k, algo='exact')
        ref_ok = bool(torch.equal(d0, d) and torch.equal(i0, i))
    print(json.dumps({'rank': rank, 'world': world, 'n': n, 'sym_ms': round(ms, 1), 'rowblock_ms': round(ms1, 1), 'sym_equals_rowblock': same,
                      'sym_equals_exact': ref_ok, 'sym': keep(st), 'rowblock': keep(st1)}), flush=True)
    del bufs, shard
dist.destroy_process_group()
