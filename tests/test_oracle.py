"""CPU tier: the oracle restatements against the committed golden vectors
(outputs of the live reference, see oracle/make_golden.py) and against live
scikit-learn when it is importable."""
import numpy as np
import pytest

from oracle import knn, edges, dbscan, compare
from semibin_b200 import synth


def _sk_available():
    try:
        import sklearn  # noqa: F401
        return True
    except Exception:
        return False


def test_knn_restated_bin_data(golden):
    g = golden('bin_data')
    for k in (20, 39):
        d, i = knn.knn_restated(g['embedding'], k)
        compare.assert_knn_parity(g[f'emb_k{k}_d'], g[f'emb_k{k}_i'], d, i, rtol=0)
        d, i = knn.knn_restated(np.ascontiguousarray(g['values'][:, :136]), k)
        # one duplicated k-mer row in the fixture -> a d == 0 pair (tie class)
        compare.assert_knn_parity(g[f'kmer_k{k}_d'], g[f'kmer_k{k}_i'], d, i, rtol=1e-9, atol=1e-7)


@pytest.mark.parametrize('direct', [False, True])
def test_knn_restated_synth(golden, direct):
    g = golden('synth_knn')
    X = synth.embeddings(4096, seed=101)
    d, i = knn.knn_restated(X, 50, direct=direct)
    res = compare.assert_knn_parity(g['emb4096_k50_d'], g['emb4096_k50_i'], d, i, rtol=0)
    assert res['n_rows_exact'] == 4096
    F = synth.kmer_features(4096, seed=101)
    d, i = knn.knn_restated(F, 50, direct=direct)
    compare.assert_knn_parity(g['kmer4096_k50_d'], g['kmer4096_k50_i'], d, i, rtol=1e-6, max_tied_frac=0.001)
    X = synth.embeddings(300, seed=103)
    d, i = knn.knn_restated(X, 299, direct=direct)
    compare.assert_knn_parity(g['emb300_k299_d'], g['emb300_k299_i'], d, i, rtol=0)
    Xl, _ = synth.long_read_case(3000, n_sample=3, seed=104)
    d, i = knn.knn_restated(Xl, 64, direct=direct)
    compare.assert_knn_parity(g['long3000_k64_d'], g['long3000_k64_i'], d, i, rtol=0)


def test_knn_row_block(golden):
    g = golden('synth_knn')
    X = synth.embeddings(1500, seed=102)
    d, i = knn.knn_restated(X, 200, q_begin=500, q_end=900)
    compare.assert_knn_parity(g['emb1500_k200_d'][500:900], g['emb1500_k200_i'][500:900], d, i, rtol=0)


def test_knn_ties(golden):
    g = golden('ties')
    for name, kw, dt in [('grid', dict(n=700, n_dup=0), np.float32), ('dup5', dict(n=700, n_dup=5), np.float32),
                         ('dup40', dict(n=700, n_dup=40), np.float32), ('grid64', dict(n=700, n_dup=5), np.float64)]:
        X = synth.tie_grid(dtype=dt, **kw)
        d, i = knn.knn_restated(X, 20)
        res = compare.assert_knn_parity(g[name + '_d'], g[name + '_i'], d, i, rtol=0)
        assert res['max_rel'] == 0.0          # the distance multiset is identical


def test_knn_errors():
    X = synth.embeddings(10, seed=1)
    with pytest.raises(ValueError):
        knn.knn_restated(X, 10)


def test_cal_kl(golden):
    g = golden('cal_kl')
    for t in (0, 1):
        m, v, kl = g[f'm{t}'], g[f'v{t}'], g[f'kl{t}']
        I, J = np.meshgrid(np.arange(128), np.arange(128), indexing='ij')
        got = edges.kl_edges(m, v, I.ravel(), J.ravel()).reshape(128, 128)
        assert got.dtype == np.float32
        np.testing.assert_array_equal(got, kl)
    # the reference's own known-answer test (test/test_cal_kl.py:4-26)
    rng = np.random.default_rng(0)
    for _ in range(8):
        m = rng.random(128).astype(np.float32); v = rng.random(128).astype(np.float32)
        mc = np.clip(m, 1e-6, None).astype(np.float64); vc = np.clip(v, 1.0, None).astype(np.float64)
        slow = np.log(np.sqrt(vc[None, :] / vc[:, None])) + ((mc[None, :] - mc[:, None]) ** 2 + vc[:, None]) / (2 * vc[None, :]) - 0.5
        slow = np.clip(slow, 1e-6, 1 - 1e-6)
        I, J = np.meshgrid(np.arange(128), np.arange(128), indexing='ij')
        got = edges.kl_edges(m, v, I.ravel(), J.ravel()).reshape(128, 128)
        assert np.allclose(slow, got, rtol=1e-4, atol=2e-6)


def test_threshold_sequence():
    seq = edges.threshold_sequence()
    assert len(seq) == 20 and seq[2] == 0.05 + 0.05 + 0.05 and seq[-1] >= 1


def test_edges_bin_data(golden):
    g = golden('bin_data')
    vals = g['values']
    depth = vals[:, 136:].astype(np.float32)
    for k in (20, 39):
        X, Y, V, thr = edges.edge_list_restated(g[f'emb_k{k}_d'], g[f'emb_k{k}_i'].astype(np.int64),
                                                g[f'kmer_k{k}_d'], g[f'kmer_k{k}_i'].astype(np.int64),
                                                depth, max_node=1, is_combined=bool(g['is_combined']),
                                                n_sample=int(g['n_sample']))
        np.testing.assert_array_equal(X, g[f'edges_k{k}_X'])
        np.testing.assert_array_equal(Y, g[f'edges_k{k}_Y'])
        np.testing.assert_array_equal(V, g[f'edges_k{k}_V'])
    assert len(g['edges_k20_X']) == 279           # SURVEY.md Appendix B probe


@pytest.mark.parametrize('S,comb', [(1, False), (3, False), (10, False), (10, True)])
def test_edges_synth(golden, S, comb):
    g = golden('synth_edges')
    emb, feat, depth, L = synth.short_read_case(3000, n_sample=S, seed=200 + S)
    f = feat
    if comb:
        f = np.concatenate([feat, depth[:, 0::2].astype(np.float64) / 100.0], axis=1)
        # cluster.py:81-86 (norm_abundance is True for these widths/sums? replicate utils.py:630-636)
        n = f.shape[1] - 136
        if n >= 20 or (n >= 5 and np.mean(np.sum(f[:, 136:], axis=1)) > 2):
            f = f / np.sum(f, axis=0)
            f = f / np.abs(f).sum(axis=1, keepdims=True)
    ad, ai = knn.knn_restated(emb, 50)
    bd, bi = knn.knn_restated(f, 50)
    for max_node in (1, 0.6):
        X, Y, V, thr = edges.edge_list_restated(ad, ai, bd, bi, depth, max_node=max_node,
                                                is_combined=comb, n_sample=S)
        tag = f'S{S}_{"comb" if comb else "kl"}_mn{max_node}'
        ref = (g[tag + '_X'], g[tag + '_Y'], g[tag + '_V'])
        res = compare.assert_edges_parity(ref, (X, Y, V), rtol=1e-12, atol=0)
        assert res['n_only_ref'] == 0 and res['n_only_got'] == 0


def test_dbscan(golden):
    g = golden('synth_dbscan')
    X, L = synth.long_read_case(4000, n_sample=2, seed=300)
    d, i = knn.knn_restated(X, 60)
    for e, eps in enumerate(dbscan.EPS_GRID):
        c, lab = dbscan.dbscan_minprop(d, i, eps, 5, L)
        np.testing.assert_array_equal(lab, g['w_labels'][e])
        c, lab = dbscan.dbscan_minprop(d, i, eps, 5, None)
        np.testing.assert_array_equal(lab, g['u_labels'][e])
        assert len(c) == g['u_ncore'][e]
    for e in (5, 6):
        c, lab = dbscan.dbscan_restated(d, i, dbscan.EPS_GRID[e], 5, None)
        np.testing.assert_array_equal(lab, g['u_labels'][e])


def test_dbscan_bin_data(golden):
    g = golden('bin_data')
    d, i = knn.knn_restated(g['long_emb_new'], 39)
    compare.assert_knn_parity(g['long_d'], g['long_i'], d, i, rtol=0)
    for e, eps in enumerate(dbscan.EPS_GRID):
        c, lab = dbscan.dbscan_minprop(d, i, eps, 5, g['lengths'])
        np.testing.assert_array_equal(lab, g['long_labels'][e])


@pytest.mark.skipif(not _sk_available(), reason='scikit-learn not importable')
def test_live_sklearn_agrees():
    X = synth.embeddings(1200, seed=77)
    rd, ri = knn.knn_sklearn(X, 30)
    d, i = knn.knn_restated(X, 30)
    compare.assert_knn_parity(rd, ri, d, i, rtol=0)
    rd, ri = knn.knn_sklearn(X, 30, q_begin=100, q_end=400)
    d, i = knn.knn_restated(X, 30, q_begin=100, q_end=400)
    compare.assert_knn_parity(rd, ri, d, i, rtol=0)


def test_integrate_restated_matches_reference_loop(golden):
    """long_read_cluster.py:12-49,122-150: the restatement against outputs of the reference's own
    get_best_bin loop (oracle/make_golden.py::integrate)."""
    from oracle import integrate as oi
    g = golden('integrate')
    for seed in range(24):
        labels, lengths, markers, minfasta = oi.random_case(seed)
        lab, n_ext = oi.integrate_restated(labels, lengths, markers, minfasta)
        np.testing.assert_array_equal(lab, g[f'rand{seed}'])
        assert n_ext == int(g[f'rand{seed}'].max()) + 1
    labels, lengths, markers = oi.synth_case(1500, 5)
    lab, _ = oi.integrate_restated(labels, lengths, markers, 20000)
    np.testing.assert_array_equal(lab, g['synth1500'])


def test_integrate_edge_cases():
    from oracle import integrate as oi
    # total length below minfasta: nothing is extracted (long_read_cluster.py:123)
    lab, n = oi.integrate_restated(np.zeros((2, 3), np.int64), [10, 10, 10], [np.array([0])] * 3, 100)
    assert n == 0 and (lab == -1).all()
    # a single contig that is long enough becomes a bin without any marker (:124-126)
    lab, n = oi.integrate_restated(np.zeros((2, 1), np.int64), [500], [np.array([], np.int32)], 100)
    assert n == 1 and lab[0] == 0
    # no bin with a marker: loop ends (:133-134)
    lab, n = oi.integrate_restated(np.zeros((1, 4), np.int64), [500] * 4, [np.array([], np.int32)] * 4, 100)
    assert n == 0


def test_numpy_log_f32_ulp():
    """The evaluator-dependent operation of the path: np.log on float32 (cal_kl, cluster.py:49-60) is not correctly
    rounded.  Measure its error on this CPU; oracle.edges.LOG_ULPS (numpy's documented 3.83 ulp + 0.5 for the
    kernel's correctly rounded log) must cover it."""
    rng = np.random.default_rng(0)
    x = np.exp(rng.uniform(0, np.log(1e9), size=2_000_000)).astype(np.float32)
    a = np.log(x)
    b = np.log(x.astype(np.float64))
    ulp = np.spacing(np.abs(b).astype(np.float32)).astype(np.float64)
    err = np.abs(a.astype(np.float64) - b) / np.maximum(ulp, 1e-300)
    assert err.max() + 0.5 <= edges.LOG_ULPS, err.max()


@pytest.mark.parametrize('S', [1, 3, 10])
def test_kl_bound_covers_log_evaluators(S):
    """Derived weight tolerance (oracle.edges.kl_factor_bound): the factor evaluated with numpy's float32 log and
    with the correctly rounded log (the CUDA kernel's) differ by less than the bound on every pair -- including
    pairs near saturation, where the relative difference is far above 1e-5 -- and the bound is not vacuous."""
    n = 60000
    emb, genome = synth.embeddings(n, seed=20241, return_genome=True)
    dep = synth.depth_table(n, S, 20241, genome)
    rng = np.random.default_rng(1)
    order = np.argsort(genome, kind='stable')
    I = np.concatenate([np.repeat(order, 8), rng.integers(0, n, 400000)])
    J = np.concatenate([np.roll(np.repeat(order, 8), 5), rng.integers(0, n, 400000)])

    def factor(mode):
        acc = None
        for s in range(S):
            kl = edges.kl_edges(dep[:, 2 * s], dep[:, 2 * s + 1], I, J, mode)
            if acc is None:
                kl *= -1
                acc = kl
            else:
                acc -= kl
        acc += S
        acc /= S
        return acc.astype(np.float64)
    a, b = factor('numpy'), factor('rounded')
    bound = edges.kl_factor_bound(dep, S, I, J)
    diff = np.abs(a - b)
    assert (diff > 0).any()                                   # the evaluators do differ
    assert (diff <= bound).all(), float((diff / bound).max())
    assert np.median(bound) < 1e-5 and bound.max() < 1e-4     # a few float32 ulps, not a blanket tolerance


def test_kmeans_restatement_matches_sklearn_golden(golden):
    """Row N4: the seeded, weighted Lloyd restatement against scikit-learn's own labels / centres / iteration counts
    (tests/golden/kmeans_cases.npz, KMeans called as SemiBin/cluster.py:236-241 calls it)."""
    from oracle import kmeans as okm
    g = golden('kmeans_cases')
    for name in ('separated', 'overlapping', 'many', 'dupseeds', 'tiny'):
        X, w, seeds = g[name + '_X'], g[name + '_w'], g[name + '_seeds']
        lab, cen, it = okm.lloyd_seeded(X, w, X[seeds])
        res = okm.compare_labels(X, g[name + '_centers'], g[name + '_labels'], lab)
        assert res['n_unexplained'] == 0 and res['n_differ'] == 0, (name, res)
        assert it == int(g[name + '_n_iter'])
        np.testing.assert_allclose(cen, g[name + '_centers'], rtol=0, atol=2e-6)
