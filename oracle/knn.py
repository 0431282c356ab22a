"""CPU restatement of sklearn's brute-force Euclidean kNN graph (TEST ORACLE).

The algorithm lives in scikit-learn 1.9.0 (third-party, pinned by the
reference at pixi.lock:4152-4153), reached from SemiBin/cluster.py:94-105 and
SemiBin/long_read_cluster.py:108-113 via
sklearn/neighbors/_graph.py:140-152 -> _base.py:756-1041 ->
metrics/_pairwise_distances_reduction/_argkmin.pyx.tp.

Restated semantics:
  norms_i = <x_i, x_i> in float64 (float32 inputs are up-cast first)
  d2_ij   = max(0, norms_i - 2 <x_i, y_j> + norms_j)        _argkmin.pyx.tp:494-502
  keep the k+1 smallest per row; a later column replaces the heap max only if
  strictly smaller -> exact ties at the boundary go to the LOWER column index
                                                               _heap.pyx:46
  rows sorted ascending by d2                                  _sorting.pyx
  float32 input : dist = f64(f32(sqrt(f64(f32(d2)))))          _argkmin.pyx.tp:285-295
  float64 input : dist = sqrt(d2)
  self removed by INDEX; if self is not among the k+1, column 0 is dropped
                                                               neighbors/_base.py:943-957
  CSR: data f64[N*k], indices i64[N*k] (distance order), indptr = arange*k
                                                               neighbors/_base.py:1031-1041
"""
import numpy as np


def _finalize(d2, idx, k, is_f32):
    """d2, idx: (n, k+1) sorted ascending by (d2, idx).  Applies the rounding
    recipe and the self-removal rule."""
    n = d2.shape[0]
    if is_f32:
        dist = np.sqrt(d2.astype(np.float32).astype(np.float64)).astype(np.float32).astype(np.float64)
    else:
        dist = np.sqrt(d2)
    return dist, idx


def _drop_self(dist, idx, q_begin):
    n, k1 = idx.shape
    rows = np.arange(q_begin, q_begin + n)[:, None]
    mask = idx != rows
    absent = mask.all(axis=1)
    mask[absent, 0] = False          # neighbors/_base.py:946-950
    # exactly one False per row
    dist = dist[mask].reshape(n, k1 - 1)
    idx = idx[mask].reshape(n, k1 - 1)
    return dist, idx


def knn_restated(X, k, q_begin=0, q_end=None, direct=False, chunk=1024):
    """Exact top-k Euclidean neighbours of rows [q_begin, q_end) of X among all
    rows of X, self excluded, following sklearn (see module docstring).

    direct=False : expanded form in float64 (what sklearn computes)
    direct=True  : direct-difference form sum((x-y)^2) in float64 (what the
                   CUDA rerank computes); differs from the expanded form only
                   by float64 rounding noise.
    Returns (dist float64 [n,k], idx int64 [n,k]) ordered by (distance, index).
    """
    X = np.ascontiguousarray(X)
    is_f32 = X.dtype == np.float32
    N = X.shape[0]
    if q_end is None:
        q_end = N
    if not (0 < k < N):
        raise ValueError("Expected n_neighbors < n_samples_fit")
    Xd = X.astype(np.float64)
    norms = np.einsum('ij,ij->i', Xd, Xd)
    out_d = np.empty((q_end - q_begin, k + 1), np.float64)
    out_i = np.empty((q_end - q_begin, k + 1), np.int64)
    for s in range(q_begin, q_end, chunk):
        e = min(s + chunk, q_end)
        if direct:
            d2 = np.empty((e - s, N), np.float64)
            for r in range(s, e):
                diff = Xd - Xd[r]
                d2[r - s] = np.einsum('ij,ij->i', diff, diff)
        else:
            d2 = norms[s:e, None] - 2.0 * (Xd[s:e] @ Xd.T) + norms[None, :]
            np.maximum(d2, 0.0, out=d2)
        # (d2, idx) lexicographic == stable argsort on d2 because columns are
        # already in ascending index order
        part = np.argsort(d2, axis=1, kind='stable')[:, :k + 1]
        out_i[s - q_begin:e - q_begin] = part
        out_d[s - q_begin:e - q_begin] = np.take_along_axis(d2, part, axis=1)
    dist, idx = _finalize(out_d, out_i, k, is_f32)
    return _drop_self(dist, idx, q_begin)


def knn_sklearn(X, k, q_begin=0, q_end=None):
    """The reference's own path: live scikit-learn (same call stack as
    kneighbors_graph(X, k, mode='distance', p=2)).  For a row block it uses
    NearestNeighbors.kneighbors on the block with k+1 and removes self the way
    kneighbors(X=None) does."""
    from sklearn.neighbors import NearestNeighbors, kneighbors_graph
    N = X.shape[0]
    if q_begin == 0 and (q_end is None or q_end == N):
        g = kneighbors_graph(X, n_neighbors=k, mode='distance', p=2)
        return g.data.reshape(N, k).copy(), g.indices.reshape(N, k).astype(np.int64)
    nn = NearestNeighbors(n_neighbors=k + 1, algorithm='brute').fit(X)
    dist, idx = nn.kneighbors(X[q_begin:q_end])
    return _drop_self(dist, idx.astype(np.int64), q_begin)


def to_csr(dist, idx, n_cols):
    from scipy.sparse import csr_matrix
    n, k = idx.shape
    return csr_matrix((dist.ravel().astype(np.float64), idx.ravel().astype(np.int64),
                       np.arange(0, n * k + 1, k, dtype=np.int64)), shape=(n, n_cols))
