"""Symmetric against one-sided filter at growing sizes: stage times, candidate counts, fallback rows and bitwise
equality of the two results (run under gpurun: python scripts/compare_filters.py 200000 1000000)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, '.')
from semibin_b200.neighbors import knn
from semibin_b200 import synth
sizes = [int(a) for a in sys.argv[1:]] or [20000, 200000, 1000000]
for n in sizes:
    X = torch.from_numpy(synth.embeddings(n)).cuda()
    k = 200
    res = {}
    for algo in ('tensor', 'tensor_onesided'):
        for it in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            d, i, st = knn(X, k, algo=algo, return_stats=True)
            torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
        res[algo] = (d, i)
        keep = {q: (round(v, 2) if isinstance(v, float) else v) for q, v in st.items() if q.startswith('ms_') or q.startswith('n_') or q == 'symmetric'}
        print(json.dumps({'n': n, 'algo': algo, 'ms': round(ms, 1), **keep}), flush=True)
        if st['n_fallback_rows'] > 0.05 * n:
            print('too many fallback rows, stopping'); sys.exit(1)
    same_i = bool((res['tensor'][1] == res['tensor_onesided'][1]).all()); same_d = bool((res['tensor'][0] == res['tensor_onesided'][0]).all())
    print(json.dumps({'n': n, 'identical_idx': same_i, 'identical_dist': same_d}), flush=True)
