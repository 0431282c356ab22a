"""Calls into the UNMODIFIED reference (TEST ORACLE; only usable where
/root/reference exists, i.e. in the build container -- never on the GPU box).

run_embed_infomap imports igraph (absent here) at cluster.py:72; a stub module
captures the exact (edges, edge_weights, vertex_weights, trials) hand-off.
"""
import sys
import types
import logging
import numpy as np

REFERENCE_ROOT = '/root/reference'


def _import_reference():
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import SemiBin  # noqa: F401
    return SemiBin


class _Capture(dict):
    pass


def _stub_igraph(capture, n_labels_fn=None):
    mod = types.ModuleType('igraph')

    class Graph:
        def add_vertices(self, v):
            capture['n_vertices'] = len(v)

        def add_edges(self, edges):
            capture['edges'] = list(edges)

        def community_infomap(self, edge_weights=None, vertex_weights=None, trials=None):
            capture['edge_weights'] = np.asarray(edge_weights).copy()
            capture['vertex_weights'] = np.asarray(vertex_weights).copy()
            capture['trials'] = trials
            return [list(range(capture['n_vertices']))]

    mod.Graph = Graph
    return mod


def capture_edge_handoff(model, data, *, max_edges, max_node, is_combined, n_sample, contig_dict,
                         device='cpu'):
    """Run SemiBin.cluster.run_embed_infomap unmodified; return the captured
    hand-off #1 (cluster.py:150-163) as (X int32, Y int32, V f64, embedding)."""
    _import_reference()
    cap = _Capture()
    saved = sys.modules.get('igraph')
    sys.modules['igraph'] = _stub_igraph(cap)
    try:
        from SemiBin.cluster import run_embed_infomap
        emb, _ = run_embed_infomap(logging.getLogger('oracle'), model, data, device=device,
                                   max_edges=max_edges, max_node=max_node, is_combined=is_combined,
                                   n_sample=n_sample, contig_dict=contig_dict, num_process=1,
                                   random_seed=None)
    finally:
        if saved is None:
            del sys.modules['igraph']
        else:
            sys.modules['igraph'] = saved
    e = cap['edges']
    X = np.array([a for a, _ in e], np.int32)
    Y = np.array([b for _, b in e], np.int32)
    return X, Y, cap['edge_weights'].astype(np.float64), emb, cap


class _IdentityModel:
    """model.embedding() that returns a precomputed embedding, so that the
    unmodified run_embed_infomap can be driven with synthetic embeddings."""
    def __init__(self, emb):
        import torch
        self._e = torch.from_numpy(np.asarray(emb, np.float32))

    def eval(self):
        return self

    def embedding(self, x):
        return self._e


def reference_edge_list(emb, features, depth, *, max_edges, max_node, is_combined, n_sample):
    """Drive the unmodified run_embed_infomap on given arrays.  `features` is
    N x 136 float64 (or N x (136+S) when combined); depth N x 2S float32."""
    import pandas as pd
    N = emb.shape[0]
    if is_combined:
        vals = np.asarray(features, np.float64)
    else:
        vals = np.concatenate([np.asarray(features, np.float64), np.asarray(depth, np.float64)], axis=1)
    data = pd.DataFrame(vals, index=[f'c{i}' for i in range(N)])
    contig_dict = {f'c{i}': 'A' for i in range(N)}
    X, Y, V, _, _ = capture_edge_handoff(_IdentityModel(emb), data, max_edges=max_edges,
                                         max_node=max_node, is_combined=is_combined,
                                         n_sample=n_sample, contig_dict=contig_dict)
    return X, Y, V


def reference_cal_kl(m, v):
    _import_reference()
    from SemiBin.cluster import cal_kl
    return cal_kl(m, v, use_ne='no')


def reference_dbscan_sweep(embedding_new, lengths, k=200):
    """long_read_cluster.py:108-120 verbatim on given arrays."""
    import warnings
    from sklearn.cluster import dbscan
    from sklearn.neighbors import kneighbors_graph
    dist_matrix = kneighbors_graph(embedding_new, n_neighbors=min(k, embedding_new.shape[0] - 1),
                                   mode='distance', p=2, n_jobs=1)
    out = []
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for eps_value in [0.01, 0.05, 0.1, 0.15, 0.2, 0.25, 0.3, 0.35, 0.4, 0.45, 0.5, 0.55]:
            _, labels = dbscan(dist_matrix, eps=eps_value, min_samples=5, n_jobs=1,
                               metric='precomputed', sample_weight=lengths)
            out.append(labels)
    return dist_matrix, out
