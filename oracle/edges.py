"""Sparse CPU restatement of SemiBin/cluster.py:13-60 (cal_kl) and :109-151
(intersect -> 1-min(d,1) -> max_node threshold -> KL factor -> edge list).
TEST ORACLE.  The reference evaluates the KL factor as a dense N x N float32
matrix (cluster.py:129-141); this restatement evaluates the same float32
expression, in the numpy evaluator's operation order (cluster.py:49-60), only at
the surviving edges, so it scales to N >= 200k.
"""
import numpy as np


def kl_edges(m, v, I, J, log_mode='numpy'):
    """float32 kl[i,j] of cal_kl (cluster.py:13-60, numpy branch :49-60) at the
    index pairs (I, J).  m1/v1 vary along columns (j), m2/v2 along rows (i):
        kl[i,j] = (log v_j - log v_i)/2 + ((m_j - m_i)^2 + v_i) / (2 v_j) - 0.5
    every step rounded to float32 exactly as numpy does, then clipped to
    [1e-6, 1-1e-6].
    log_mode='numpy'  : np.log on float32, what the reference evaluates (SIMD kernel, not correctly rounded)
    log_mode='rounded': the correctly rounded float32 log (float64 log rounded once), what the CUDA kernel evaluates;
                        used by the CPU tests to exercise the derived bound kl_factor_bound without a GPU."""
    m = np.clip(np.asarray(m, np.float32), 1e-6, None)
    v = np.clip(np.asarray(v, np.float32), 1.0, None)
    assert m.dtype == np.float32 and v.dtype == np.float32
    lv = np.log(v) if log_mode == 'numpy' else np.log(v.astype(np.float64)).astype(np.float32)
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


def edge_list_restated(emb_d, emb_i, kmer_d, kmer_i, depth, *, max_node, is_combined, n_sample, log_mode='numpy'):
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
            kl = kl_edges(depth[:, 2 * s], depth[:, 2 * s + 1], I, J, log_mode)
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


# ---------------------------------------------------------------------------
# derived tolerance of the KL factor (the one evaluator-dependent operation of the path)
# ---------------------------------------------------------------------------
# Every float32 operation of cal_kl's numpy branch (cluster.py:49-60) is IEEE and is reproduced bit for bit by the
# CUDA kernel EXCEPT `np.log` on float32: numpy's SIMD log is not correctly rounded (its own documentation states a
# maximum error of 3.83 ulp for the AVX2/AVX512F kernels; tests/test_oracle.py::test_numpy_log_f32_ulp measures it on
# the running CPU) and differs between CPU dispatch targets and from numexpr's (the reference's other evaluator,
# cluster.py:32-45).  The kernel takes the correctly rounded value (<= 0.5 ulp).  The two evaluations of log v can
# therefore differ by up to LOG_ULPS float32 ulps of log v, and that difference propagates through the remaining
# (identical, individually rounded) operations as bounded below.
LOG_ULPS = 3.83 + 0.5


def _ulp32(x):
    return np.spacing(np.abs(np.asarray(x, np.float64)).astype(np.float32)).astype(np.float64)


def kl_factor_bound(depth, n_sample, I, J, log_ulps=LOG_ULPS):
    """Upper bound of |acc_a - acc_b| for the float32 factor acc = (S - sum_s kl_s)/S of cluster.py:129-140 at the
    index pairs (I, J) when the two evaluations differ ONLY in the float32 log (by <= log_ulps ulp each).
    Per sample:  kl = clip(((lvj - lvi)/2 + m_dif) - 0.5):
        input difference of the subtraction   e0 = log_ulps * (ulp(lvj) + ulp(lvi))
        after rn(lvj - lvi)                   e1 = e0 + ulp(lvj - lvi)           (rounding is monotone: two results
        after /2 (exact)                      e1 / 2                              differ by at most the input
        after + m_dif                         e2 = e1/2 + ulp(v_div + m_dif)      difference plus one ulp of the
        after - 0.5                           e3 = e2 + ulp(result)               result)
        clip is 1-Lipschitz.
    Accumulation (acc = -kl_0; acc -= kl_s; acc += S; acc /= S): every step adds one ulp of its result."""
    I = np.asarray(I, np.int64); J = np.asarray(J, np.int64)
    tot = np.zeros(len(I), np.float64)
    part = np.zeros(len(I), np.float64)
    for s in range(n_sample):
        m = np.clip(np.asarray(depth[:, 2 * s], np.float32), 1e-6, None).astype(np.float64)
        v = np.clip(np.asarray(depth[:, 2 * s + 1], np.float32), 1.0, None).astype(np.float64)
        lv = np.log(v)
        e0 = log_ulps * (_ulp32(lv[J]) + _ulp32(lv[I]))
        a = lv[J] - lv[I]
        e1 = e0 + _ulp32(a)
        m_dif = ((m[J] - m[I]) ** 2 + v[I]) / (2 * v[J])
        s1 = a / 2 + m_dif
        e2 = e1 / 2 + _ulp32(np.abs(s1) + e1)
        s2 = s1 - 0.5
        e3 = e2 + _ulp32(np.abs(s2) + e2)
        kl = np.clip(s2, 1e-6, 1 - 1e-6)
        # both evaluations clipped to the same end -> no difference at all
        same_end = ((s2 - e3 >= 1 - 1e-6) | (s2 + e3 <= 1e-6))
        e3 = np.where(same_end, 0.0, e3)
        part = part + kl
        tot = tot + e3 + (_ulp32(part) if s > 0 else 0.0)
    S = float(n_sample)
    tot = tot + _ulp32(S)                       # += n_sample
    return tot / S + _ulp32(1.0)                # /= n_sample


def kl_weight_bound(depth, n_sample, I, J, w_emb=None, log_ulps=LOG_ULPS):
    """|V_a - V_b| <= w_emb * |acc_a - acc_b| with w_emb = 1 - min(d, 1) <= 1 (cluster.py:141)."""
    b = kl_factor_bound(depth, n_sample, I, J, log_ulps)
    return b if w_emb is None else b * np.asarray(w_emb, np.float64)
