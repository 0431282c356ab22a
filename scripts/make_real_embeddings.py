#!/usr/bin/env python
"""Realism check: embeddings produced by a SHIPPED SemiBin model.

Runs in the build container only (it reads /root/reference): synthetic k-mer features of N contigs (synth.kmer_features:
per-genome Dirichlet 4-mer profiles, per-contig Poisson counts, lognormal genome sizes) go through the reference's own
`model_load` + `model.embedding` (SemiBin/semi_supervised_model.py, models/global_model.pt, CPU) and the float32 [N, 100]
result is written to scratch/ (git-ignored; it travels to the GPU box with the next gpurun snapshot), where
`python bench.py --embedding-file scratch/emb_global_<N>.npy` times the kNN graph on it.

usage: python scripts/make_real_embeddings.py [N=1000000] [genome_mean_size=200]
"""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
from semibin_b200 import synth                       # noqa: E402
from SemiBin.semi_supervised_model import model_load  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
mean_size = int(sys.argv[2]) if len(sys.argv) > 2 else 200
model = model_load('/root/reference/SemiBin/models/global_model.pt', torch.device('cpu'), warn_on_old_format=False)
model.eval()
rng = np.random.default_rng(20242)
genome, _ = synth._genome_assignment(rng, n, mean_size)
feat = synth.kmer_features(n, seed=20242, genome=genome)
out = np.empty((n, 100), np.float32)
with torch.no_grad():
    for s in range(0, n, 65536):
        out[s:s + 65536] = model.embedding(torch.from_numpy(feat[s:s + 65536]).float()).numpy()
os.makedirs(os.path.join(ROOT, 'scratch'), exist_ok=True)
path = os.path.join(ROOT, 'scratch', f'emb_global_{n}.npy')
np.save(path, out)
d = np.linalg.norm(out[:2000, None, :] - out[None, :2000, :], axis=2)
print(path, out.shape, 'abs mean', float(np.abs(out).mean()), 'pairwise distance quantiles (2000 rows)',
      np.quantile(d[d > 0], [0.001, 0.01, 0.1, 0.5, 0.9]).round(4).tolist())
