import sys, time, json
import numpy as np, torch
sys.path.insert(0, '.')
from semibin_b200.neighbors import kneighbors_graph
from semibin_b200 import synth
n, k = 1000000, 200
X = torch.from_numpy(synth.embeddings(n)).pin_memory().numpy()
g, st = kneighbors_graph(X, k, algo='tensor_symmetric', return_stats=True)
print({q: round(v, 1) for q, v in st.items() if q.startswith('ms_')}, st['n_candidates'], st['n_reranked'])
