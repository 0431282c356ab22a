"""GPU tier, row N4: the batched seeded / weighted KMeans of the re-clustering step (SemiBin/cluster.py:229-246)
through the C-ABI (ctypes -> libsbk.so: sbk_kmeans_batch) against scikit-learn's own results (committed goldens, and
live scikit-learn where importable) and the oracle restatement."""
import numpy as np
import pytest

from oracle import kmeans as okm

pytestmark = pytest.mark.gpu
CASES = ('separated', 'overlapping', 'many', 'dupseeds', 'tiny')


def _run(problems):
    """problems: list of (X f32 [n,d], w, seed_rows) with equal d -> per-problem (labels, n_iter, centers)."""
    import torch
    from semibin_b200.recluster import kmeans_batch
    X = np.concatenate([p[0] for p in problems]).astype(np.float32)
    w = np.concatenate([np.asarray(p[1], np.float32) for p in problems])
    C0 = np.concatenate([p[0][p[2]] for p in problems]).astype(np.float32)
    po = np.concatenate([[0], np.cumsum([len(p[0]) for p in problems])]).astype(np.int64)
    co = np.concatenate([[0], np.cumsum([len(p[2]) for p in problems])]).astype(np.int32)
    lab, nit, cen = kmeans_batch(torch.from_numpy(X).cuda(), torch.from_numpy(w).cuda(), po, torch.from_numpy(C0).cuda(), co,
                                 return_centers=True)
    lab, nit, cen = lab.cpu().numpy(), nit.cpu().numpy(), cen.cpu().numpy()
    return [(lab[po[t]:po[t + 1]], int(nit[t]), cen[co[t]:co[t + 1]]) for t in range(len(problems))]


def test_golden_cases_one_batch_and_alone(golden):
    """The five committed scikit-learn cases (2 to 44 Lloyd iterations, 3 to 12 clusters, duplicated seeds): labels,
    iteration counts and centres, solved together in one launch and one at a time."""
    g = golden('kmeans_cases')
    probs = []
    for name in CASES:
        X = g[name + '_X']
        pad = 110 - X.shape[1]                                   # one launch needs one width: zero columns change nothing
        probs.append((np.pad(X, ((0, 0), (0, pad))), g[name + '_w'], g[name + '_seeds']))
    for res, name in zip(_run(probs), CASES):
        lab, nit, cen = res
        d = g[name + '_X'].shape[1]
        cmp = okm.compare_labels(g[name + '_X'], g[name + '_centers'], g[name + '_labels'], lab)
        assert cmp['n_unexplained'] == 0 and cmp['n_differ'] <= 0.001 * cmp['n'] + 1, (name, cmp)
        assert nit == int(g[name + '_n_iter']), (name, nit, int(g[name + '_n_iter']))
        np.testing.assert_allclose(cen[:, :d], g[name + '_centers'], rtol=0, atol=2e-5)
    for name in ('dupseeds', 'tiny'):
        (lab, nit, cen), = _run([(g[name + '_X'], g[name + '_w'], g[name + '_seeds'])])
        cmp = okm.compare_labels(g[name + '_X'], g[name + '_centers'], g[name + '_labels'], lab)
        assert cmp['n_unexplained'] == 0 and cmp['n_differ'] == 0, (name, cmp)


def test_batch_of_bins_vs_live_sklearn_and_oracle():
    """60 bins of 30 to 4000 contigs (D = 103, 2 to 9 seeds each) in one launch against live scikit-learn called the way
    the reference calls it, and against the oracle (which must agree with the kernel bit for bit in labels)."""
    sk = pytest.importorskip('sklearn.cluster')
    from semibin_b200.synth import recluster_problem as kmeans_problem
    rng = np.random.default_rng(77)
    probs = []
    for t in range(60):
        n = int(rng.integers(30, 4000)); K = int(rng.integers(2, 10))
        probs.append(kmeans_problem(1000 + t, n, 103, K, float(rng.uniform(0.05, 0.4)), float(rng.uniform(0.04, 0.5)),
                                    int(rng.integers(0, 3))))
    res = _run(probs)
    n_diff = 0
    for (X, w, seeds), (lab, nit, cen) in zip(probs, res):
        km = sk.KMeans(n_clusters=len(seeds), init=X[seeds], n_init=1, random_state=0).fit(X, sample_weight=w)
        cmp = okm.compare_labels(X, km.cluster_centers_, km.labels_, lab)
        assert cmp['n_unexplained'] == 0, cmp
        n_diff += cmp['n_differ']
        assert abs(nit - km.n_iter_) <= 1, (nit, km.n_iter_)
        ol, oc, oit = okm.lloyd_seeded(X, w, X[seeds])
        np.testing.assert_array_equal(ol, lab)
        assert oit == nit
    assert n_diff <= 5, n_diff


def test_recluster_labels_bookkeeping():
    """cluster.py:221-253 around the batched launch: bins with one seed / below minfasta keep one label, split bins get
    len(seed) labels, labels handed out in bin order -- against the oracle's restatement of the same loop."""
    from semibin_b200.recluster import recluster_labels
    from semibin_b200.synth import recluster_problem as kmeans_problem
    rng = np.random.default_rng(5)
    Xs, labs, seeds_by_bin, lens = [], [], {}, []
    off = 0
    for b in range(25):
        n = int(rng.integers(5, 600)); K = int(rng.integers(1, 5))
        X, w, seeds = kmeans_problem(300 + b, max(n, K + 1), 103, K, 0.2, 0.3)
        Xs.append(X); labs.append(np.full(len(X), b)); lens.append(w)
        if b % 7 != 3:
            seeds_by_bin[b] = list(off + seeds)
        off += len(X)
    E = np.concatenate(Xs); cl = np.concatenate(labs); L = np.concatenate(lens)
    perm = rng.permutation(len(E))                              # bins interleaved, as contig order is
    inv = np.argsort(perm)
    E, cl, L = E[perm], cl[perm], L[perm]
    seeds_by_bin = {b: [int(inv[s]) for s in v] for b, v in seeds_by_bin.items()}
    total = {b: int(L[cl == b].sum()) for b in range(25)}
    minfasta = int(np.median(list(total.values())) * 0.3)
    got = recluster_labels(E, cl, seeds_by_bin, L, total, minfasta)
    ref = okm.recluster_labels(E, cl, seeds_by_bin, L, total, minfasta)
    np.testing.assert_array_equal(got, ref)
    assert got.min() == 0 and len(np.unique(got)) == got.max() + 1


def test_errors_and_empty():
    import torch
    from semibin_b200.recluster import kmeans_batch
    X = torch.zeros((3, 8), device='cuda'); w = torch.ones(3, device='cuda')
    with pytest.raises(ValueError, match='n_samples=3 should be >= n_clusters=4'):
        kmeans_batch(X, w, [0, 3], torch.zeros((4, 8), device='cuda'), [0, 4])
    lab, nit = kmeans_batch(X, w, [0], torch.zeros((0, 8), device='cuda'), [0])
    assert lab.numel() == 3 and nit.numel() == 0
    # identical points, two seeds: nothing to relocate (max distance 0, _k_means_common.pyx:192-195), all in cluster 0
    lab, nit = kmeans_batch(X, w, [0, 3], torch.zeros((2, 8), device='cuda'), [0, 2])
    assert lab.cpu().tolist() == [0, 0, 0]
