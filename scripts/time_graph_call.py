import sys, time, json
import numpy as np, torch
sys.path.insert(0, '.')
from semibin_b200.neighbors import kneighbors_graph, knn_host
from semibin_b200 import synth
n, k = 1000000, 200
X = torch.from_numpy(synth.embeddings(n)).pin_memory().numpy()
for algo in ('tensor_symmetric', 'tensor_onesided'):
    for it in range(3):
        t0 = time.perf_counter()
        g, st = kneighbors_graph(X, k, algo=algo, return_stats=True)
        ms = (time.perf_counter() - t0) * 1e3
        print(algo, round(ms, 1), {q: round(v, 1) for q, v in st.items() if q.startswith('ms_')}, st['n_chunks'], st['n_filter_launches'], st['n_launches'], st.get('host_ms'), flush=True)
        g = None
