"""CPU restatement of the long-read eps sweep (TEST ORACLE):
SemiBin/long_read_cluster.py:115-120 -> sklearn/cluster/_dbscan.py:397-474 and
_dbscan_inner.pyx on a *precomputed sparse kNN graph*.

  nb(i)   = {i} U { j in kNN(i) : d_ij <= eps }        (_dbscan.py:430-436, neighbors/_base.py:339-381)
  core(i) = sum_{j in nb(i)} weight_j >= min_samples   (_dbscan.py:451-463)
  labels  : for i ascending, if unlabelled and core: depth-first expansion;
            every reached unlabelled point takes the label, only core points
            expand; label += 1                          (_dbscan_inner.pyx)

The expansion result does not depend on the traversal order: label(v) is the
rank of  seed(v) = min{ u core : u reaches v through core points }  among the
distinct seeds.  `dbscan_restated` is the literal DFS; `dbscan_minprop` is the
order-free characterisation the CUDA kernel implements (both are checked
against live sklearn in the tests).
"""
import numpy as np

EPS_GRID = [0.01, 0.05, 0.1, 0.15, 0.2, 0.25, 0.3, 0.35, 0.4, 0.45, 0.5, 0.55]  # long_read_cluster.py:116


def neighbourhood_sizes(dist, eps):
    """Rows are sorted ascending -> the radius query is a prefix cut."""
    return (dist <= eps).sum(axis=1)


def core_mask(dist, idx, eps, min_samples, sample_weight):
    N = dist.shape[0]
    w = np.ones(N, np.float64) if sample_weight is None else np.asarray(sample_weight, np.float64)
    inside = dist <= eps
    tot = w + (w[idx] * inside).sum(axis=1)
    return tot >= min_samples


def dbscan_restated(dist, idx, eps, min_samples=5, sample_weight=None):
    """Literal DFS of _dbscan_inner.pyx. Returns (core_idx, labels)."""
    N, k = idx.shape
    core = core_mask(dist, idx, eps, min_samples, sample_weight)
    plen = neighbourhood_sizes(dist, eps)
    labels = np.full(N, -1, np.intp)
    label_num = 0
    for i in range(N):
        if labels[i] != -1 or not core[i]:
            continue
        stack = []
        while True:
            if labels[i] == -1:
                labels[i] = label_num
                if core[i]:
                    # neighbourhood order: self is part of it; order irrelevant
                    for v in idx[i, :plen[i]]:
                        if labels[v] == -1:
                            stack.append(v)
            if not stack:
                break
            i = stack.pop()
        label_num += 1
    return np.where(core)[0], labels


def dbscan_minprop(dist, idx, eps, min_samples=5, sample_weight=None):
    """Order-free restatement: iterate L[j] = min(L[j], L[u]) over edges u->j
    with u core, from L[u] = u for core u; then rank the seeds."""
    N, k = idx.shape
    core = core_mask(dist, idx, eps, min_samples, sample_weight)
    inside = dist <= eps
    INF = np.iinfo(np.int64).max
    L = np.where(core, np.arange(N), INF)
    src = np.repeat(np.arange(N), k).reshape(N, k)
    e_src = src[inside & core[:, None]]
    e_dst = idx[inside & core[:, None]]
    while True:
        new = L.copy()
        np.minimum.at(new, e_dst, L[e_src])
        if (new == L).all():
            break
        L = new
    labels = np.full(N, -1, np.intp)
    seeds = np.unique(L[L != INF])
    has = L != INF
    labels[has] = np.searchsorted(seeds, L[has])
    return np.where(core)[0], labels


def dbscan_sklearn(dist, idx, eps, min_samples=5, sample_weight=None):
    """The reference call itself (long_read_cluster.py:117-119)."""
    import warnings
    from sklearn.cluster import dbscan
    from .knn import to_csr
    g = to_csr(dist, idx, dist.shape[0])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        core, labels = dbscan(g, eps=eps, min_samples=min_samples, metric='precomputed',
                              sample_weight=sample_weight)
    return core, labels
