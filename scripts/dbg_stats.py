"""Full kNN stage statistics for one tensor-path run (debug helper).  N=... K=... env."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semibin_b200 import synth
from semibin_b200.neighbors import knn
n, k = int(os.environ.get("N", 1000000)), int(os.environ.get("K", 200))
X = torch.from_numpy(synth.embeddings(n, seed=20242)).cuda()
for it in range(2):
    d, i, st = knn(X, k, algo='tensor', return_stats=True)
    if os.environ.get('FULL'):
        print({a: (round(b, 2) if isinstance(b, float) else b) for a, b in st.items()}, flush=True)
    else:
        print('probe=%s filter %.2f rerank %.2f sample %.2f fallback %.2f (%d rows) cand %d reranked %d' % (
            os.environ.get('SBK_TC_PROBE'), st['ms_filter'], st['ms_rerank'], st['ms_sample'], st['ms_fallback'],
            st['n_fallback_rows'], st['n_candidates'], st['n_reranked']), flush=True)
