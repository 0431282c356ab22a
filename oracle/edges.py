"""Sparse CPU restatement of SemiBin/cluster.py:13-60 (cal_kl) and :109-151
(intersect -> 1-min(d,1) -> max_node threshold -> KL factor -> edge list).
TEST ORACLE.  The reference evaluates the KL factor as a dense N x N float32
matrix (cluster.py:129-141); this restatement evaluates the same float32
expression, in the numpy evaluator's operation order (cluster.py:49-60), only at
the surviving edges, so it scales to N >= 200k.
"""
import numpy as np


def kl_edges(m, v, I, J):
    """float32 kl[i,j] of cal_kl (cluster.py:13-60, numpy branch :49-60) at the
    index pairs (I, J).  m1/v1 vary along columns (j), m2/v2 along rows (i):
        kl[i,j] = (log v_j - log v_i)/2 + ((m_j - m_i)^2 + v_i) / (2 v_j) - 0.5
    every step rounded to float32 exactly as numpy does, then clipped to
    [1e-6, 1-1e-6]."""
    m = np.clip(np.asarray(m, np.float32), 1e-6, None)
    v = np.clip(np.asarray(v, np.float32), 1.0, None)
    assert m.dtype == np.float32 and v.dtype == np.float32
    lv = np.log(v)                       # float32 log
    v_div = lv[J] - lv[I]
    v_div /= np.float32(2.0)
    m_dif = m[J] - m[I]
    m_dif **= 2
    m_dif += v[I]
    m_dif /= (2 * v[J])
    v_div += m_dif
    v_div -= np.float32(0.5)
    np.clip(v_div, 1e-6, 1 - 1e-6, out=v_div)
    assert v_div.dtype == np.float32
    return v_div


def threshold_sequence():
    """The values `threshold` takes in cluster.py:118-125 (repeated += 0.05 in
    float64; the drift is part of the semantics)."""
    out = []
    t = 0
    while t < 1:
        t += 0.05
        out.append(t)
    return out


def max_node_threshold(rowmax, num_contigs, max_node):
    """cluster.py:118-126."""
    threshold = 0
    while threshold < 1:
        threshold += 0.05
        n_above = np.sum(rowmax > threshold)
        if round(n_above / num_contigs, 2) < max_node:
            break
    threshold -= 0.05
    return threshold


def edge_list_restated(emb_d, emb_i, kmer_d, kmer_i, depth, *, max_node, is_combined, n_sample):
    """emb_* / kmer_*: (N,k) kNN results (distance order) of the embedding
    graph (cluster.py:94-99) and the feature graph (:100-105).
    depth: float32 (N, 2*n_sample) interleaved [mean_s, var_s] (cluster.py:88).
    Returns (X int32, Y int32, V float64) == what cluster.py:144-151 hands to
    igraph, and the threshold used."""
    N, k = emb_i.shape
    rows = np.repeat(np.arange(N, dtype=np.int64), k)
    # :109-113  intersect; explicit zeros are eliminated on both sides
    key_b = rows * N + kmer_i.ravel()
    key_b = key_b[kmer_d.ravel() != 0]
    key_a = rows * N + emb_i.ravel()
    da = emb_d.ravel().astype(np.float64)
    keep = np.isin(key_a, key_b) & (da != 0)
    I = rows[keep]; J = emb_i.ravel()[keep].astype(np.int64); d = da[keep]
    # :115-116
    w = 1 - np.minimum(d, 1.0)
    # :118-128  (row max over the sparse row; implicit zeros count, w >= 0)
    rowmax = np.zeros(N, np.float64)
    np.maximum.at(rowmax, I, w)
    thr = max_node_threshold(rowmax, N, max_node)
    keep = ~(w <= thr)
    keep &= (w != 0)                     # eliminate_zeros after the clip (d >= 1)
    I, J, w = I[keep], J[keep], w[keep]
    if not is_combined:
        acc = None
        for s in range(n_sample):
            kl = kl_edges(depth[:, 2 * s], depth[:, 2 * s + 1], I, J)
            if acc is None:
                kl *= -1
                acc = kl
            else:
                acc -= kl
        acc += n_sample
        acc /= n_sample
        assert acc.dtype == np.float32
        w = w * acc.astype(np.float64)
    keep = ~(w <= 1e-6)
    I, J, w = I[keep], J[keep], w[keep]
    up = J > I
    I, J, w = I[up], J[up], w[up]
    order = np.lexsort((J, I))
    return I[order].astype(np.int32), J[order].astype(np.int32), w[order], thr
