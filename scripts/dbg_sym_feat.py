import sys, json
import numpy as np, torch
sys.path.insert(0, '.')
from semibin_b200.neighbors import knn
from semibin_b200 import synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 450000
F = synth.kmer_features(N, seed=20271)
X = np.ascontiguousarray(np.concatenate([F, synth.depth_table(N, 5, seed=20271)], axis=1).astype(np.float64))
Xd = torch.from_numpy(X).cuda()
for algo in ('tensor_onesided', 'tensor_symmetric'):
    d, i, st = knn(Xd, 50, algo=algo, return_stats=True)
    print(algo, json.dumps({q: (round(v, 2) if isinstance(v, float) else v) for q, v in st.items()}), flush=True)
