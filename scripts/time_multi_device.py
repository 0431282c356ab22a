"""Timing split of the single-process multi-device kneighbors_graph: python scripts/time_multi_device.py 1 2 8
(run under gpurun --gpus N)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, '.')
from semibin_b200.neighbors import kneighbors_graph
from semibin_b200.synth import embeddings as synth_embeddings
n, k = 1000000, 200
X = torch.from_numpy(synth_embeddings(n)).pin_memory().numpy()
counts = [int(a) for a in sys.argv[1:]] or [1, 2]
for devs in [list(range(c)) for c in counts]:
    for it in range(3):
        t0 = time.perf_counter()
        g, st = kneighbors_graph(X, k, mode='distance', p=2, devices=devs, return_stats=True)
        ms = (time.perf_counter() - t0) * 1e3
        per = st.get('per_device', {devs[0]: st})
        print(json.dumps({'devs': devs, 'it': it, 'ms': round(ms, 1), 'host_ms': st.get('host_ms'),
                          'per_device': {d: {q: round(v, 1) for q, v in s.items() if (q.endswith('_ms') or q.startswith('ms_')) and isinstance(v, (int, float))} for d, s in per.items()}}), flush=True)
        g = None
