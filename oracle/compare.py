"""Tie-aware parity comparators (TEST ORACLE).

north_star contract: neighbour index sets identical up to exact-distance ties;
distances / edge weights within 1e-5 relative; DBSCAN labels bit-exact.
"""
import numpy as np


def compare_knn(ref_d, ref_i, got_d, got_i, rtol=1e-5, tie_rtol=5e-7, atol=0.0):
    """Compare two (dist, idx) kNN results, rows ordered by ascending distance.

    Returns a dict with:
      n_rows_exact   rows whose index arrays are identical position by position
      n_rows_tied    rows that differ only inside groups of tied distances
                     (|d_a - d_b| <= tie_rtol * d, i.e. equal after float32
                     rounding up to a couple of ulps) or at a tied k-th boundary
      n_rows_bad     everything else (must be 0 for parity)
      max_rel        max relative distance error
    """
    ref_d = np.asarray(ref_d, np.float64); got_d = np.asarray(got_d, np.float64)
    ref_i = np.asarray(ref_i, np.int64); got_i = np.asarray(got_i, np.int64)
    assert ref_d.shape == got_d.shape == ref_i.shape == got_i.shape
    denom = np.maximum(np.abs(ref_d), 1e-300)
    rel = np.abs(ref_d - got_d) / denom
    rel[(ref_d == got_d)] = 0.0
    if atol:
        rel[np.abs(ref_d - got_d) <= atol] = 0.0
    max_rel = float(rel.max()) if rel.size else 0.0
    bad_dist_rows = np.where((rel > rtol).any(axis=1))[0]
    diff_rows = np.where((ref_i != got_i).any(axis=1))[0]
    n_tied = 0
    bad = []
    for r in diff_rows:
        d = ref_d[r]; a = ref_i[r]; b = got_i[r]
        k = len(d)
        # tie groups over the reference distances
        tol = tie_rtol * np.maximum(d[1:], 1e-300) + atol
        brk = np.concatenate(([True], (d[1:] - d[:-1]) > tol))
        gid = np.cumsum(brk) - 1
        ok = True
        last = gid[-1]
        for g in np.unique(gid[a != b]):
            sel = gid == g
            sa = set(a[sel].tolist()); sb = set(b[sel].tolist())
            if sa == sb:
                continue
            if g == last:
                # boundary group: members may be swapped for other refs at the
                # same (tied) distance that did not make the cut
                continue
            ok = False
            break
        if ok:
            n_tied += 1
        else:
            bad.append(int(r))
    bad_rows = sorted(set(bad) | set(bad_dist_rows.tolist()))
    return dict(n_rows=ref_i.shape[0],
                n_rows_exact=int(ref_i.shape[0] - len(diff_rows)),
                n_rows_tied=int(n_tied), n_rows_bad=len(bad_rows),
                bad_rows=bad_rows[:16], max_rel=max_rel)


def assert_knn_parity(ref_d, ref_i, got_d, got_i, rtol=1e-5, max_tied_frac=1.0, **kw):
    res = compare_knn(ref_d, ref_i, got_d, got_i, rtol=rtol, **kw)
    assert res['n_rows_bad'] == 0, res
    assert res['n_rows_tied'] <= max_tied_frac * max(res['n_rows'], 1), res
    return res


def compare_edges(ref, got, rtol=1e-5, atol=0.0, cut=1e-6, bound_fn=None, explain=0):
    """ref/got: (X int, Y int, V float64) edge lists sorted row-major.

    Weights must agree within  rtol*|v| + bound(edge)  where `bound_fn(X, Y)` is the DERIVED per-edge bound of the
    one evaluator-dependent operation on the path (oracle.edges.kl_weight_bound: the float32 log of cal_kl); with
    bound_fn=None (combined mode: no KL factor) the weights must agree within rtol alone.  `atol` is kept only for
    callers that compare against goldens of unknown provenance and defaults to 0.
    Edges must be the same set except edges whose weight sits within that same tolerance of the `<= 1e-6` survival
    cut (cluster.py:143) and edges removed/kept because of an exact-zero distance (tie class).
    explain=n lists up to n offending edges (x, y, v_ref, v_got, allowed)."""
    rx, ry, rv = [np.asarray(a) for a in ref]
    gx, gy, gv = [np.asarray(a) for a in got]
    rk = rx.astype(np.int64) << 32 | ry.astype(np.int64)
    gk = gx.astype(np.int64) << 32 | gy.astype(np.int64)
    assert (np.diff(rk) > 0).all(), "reference edges not strictly row-major sorted"
    sorted_ok = bool((np.diff(gk) > 0).all())
    common, ri, gi = np.intersect1d(rk, gk, assume_unique=True, return_indices=True)
    only_r = np.setdiff1d(np.arange(len(rk)), ri)
    only_g = np.setdiff1d(np.arange(len(gk)), gi)

    def bnd(x, y):
        if bound_fn is None or len(x) == 0:
            return np.zeros(len(x), np.float64)
        return np.asarray(bound_fn(x, y), np.float64)

    dv = np.abs(rv[ri] - gv[gi])
    allowed = rtol * np.abs(rv[ri]) + atol + bnd(rx[ri], ry[ri])
    bad_w = dv > allowed
    near_cut_r = np.abs(rv[only_r] - cut) <= rtol * cut + atol + bnd(rx[only_r], ry[only_r])
    near_cut_g = np.abs(gv[only_g] - cut) <= rtol * cut + atol + bnd(gx[only_g], gy[only_g])
    out = dict(n_ref=len(rk), n_got=len(gk), n_common=len(common), sorted_ok=sorted_ok,
               n_only_ref=int(len(only_r)), n_only_got=int(len(only_g)),
               n_only_ref_unexplained=int((~near_cut_r).sum()),
               n_only_got_unexplained=int((~near_cut_g).sum()),
               n_bad_weight=int(bad_w.sum()),
               n_weight_differs=int((dv > 0).sum()),
               n_beyond_rtol=int((dv > rtol * np.abs(rv[ri])).sum()),
               max_rel_w=float((dv / np.maximum(np.abs(rv[ri]), 1e-300)).max()) if len(ri) else 0.0,
               max_excess=float((dv / np.maximum(allowed, 1e-300)).max()) if len(ri) else 0.0)
    if explain:
        # the worst edges relative to what the derived bound allows (offenders first)
        order = np.argsort(-(dv / np.maximum(allowed, 1e-300)))[:explain]
        out['worst'] = [dict(x=int(rx[ri][t]), y=int(ry[ri][t]), v_ref=float(rv[ri][t]), v_got=float(gv[gi][t]),
                             diff=float(dv[t]), allowed=float(allowed[t])) for t in order]
    return out


def assert_edges_parity(ref, got, **kw):
    res = compare_edges(ref, got, **kw)
    assert res['sorted_ok'], res
    assert res['n_only_ref_unexplained'] == 0 and res['n_only_got_unexplained'] == 0, res
    assert res['n_bad_weight'] == 0, res
    return res
