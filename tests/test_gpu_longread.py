"""GPU tier: eps sweep + DBSCAN labels (hand-off #2) -- bit-exact."""
import numpy as np
import pytest

from oracle import dbscan as od, knn as ok
from semibin_b200 import synth

pytestmark = pytest.mark.gpu


def test_bin_data_labels(golden):
    from semibin_b200.long_read import dbscan_sweep
    g = golden('bin_data')
    labs = dbscan_sweep((g['long_d'], g['long_i']), od.EPS_GRID, min_samples=5, sample_weight=g['lengths'])
    for e in range(12):
        np.testing.assert_array_equal(labs[e], g['long_labels'][e])
        assert labs[e].dtype == np.intp


def test_synth_labels_weighted_and_unweighted(golden):
    from semibin_b200.long_read import long_read_labels, dbscan_sweep, dbscan
    g = golden('synth_dbscan')
    X, L = synth.long_read_case(4000, n_sample=2, seed=300)
    emb, dep = X[:, :100], None
    # full hot path: kNN on the GPU, then sweep (weights = contig lengths)
    labs, (dist, idx) = long_read_labels(X, None, L, n_sample=2, is_combined=True, n_neighbors=60,
                                         return_graph=True)
    for e in range(12):
        np.testing.assert_array_equal(labs[e], g['w_labels'][e])
    # unit weights: non-core points and noise
    labs, cores = dbscan_sweep((dist, idx), od.EPS_GRID, min_samples=5, sample_weight=None, return_core=True)
    for e in range(12):
        np.testing.assert_array_equal(labs[e], g['u_labels'][e])
        assert len(cores[e]) == g['u_ncore'][e]
    c, lab = dbscan((dist, idx), eps=0.25, min_samples=5, metric='precomputed')
    np.testing.assert_array_equal(lab, g['u_labels'][5])


def test_dbscan_csr_input_and_50k():
    """50k contigs: labels vs the order-free oracle restatement on the same graph."""
    from semibin_b200.neighbors import kneighbors_graph
    from semibin_b200.long_read import dbscan_sweep
    X, L = synth.long_read_case(50000, n_sample=1, seed=301)
    G = kneighbors_graph(X, 100, mode='distance', p=2)
    labs = dbscan_sweep(G, od.EPS_GRID, min_samples=5, sample_weight=L)
    d = G.data.reshape(50000, 100); i = G.indices.reshape(50000, 100)
    for e in (0, 4, 6, 11):
        c, ref = od.dbscan_minprop(d, i, od.EPS_GRID[e], 5, L)
        np.testing.assert_array_equal(labs[e], ref)
    # idempotence / label properties
    for lab in labs:
        u = np.unique(lab[lab >= 0])
        assert (u == np.arange(len(u))).all()


def test_integrate_golden(golden):
    """Ensemble integration (long_read_cluster.py:12-49,122-150) against outputs of the reference's own loop."""
    from oracle import integrate as oi
    from semibin_b200.long_read import integrate_labels
    g = golden('integrate')
    for seed in range(24):
        labels, lengths, markers, minfasta = oi.random_case(seed)
        lab, n_ext = integrate_labels(labels, lengths, markers, minfasta)
        np.testing.assert_array_equal(lab, g[f'rand{seed}'])
        assert n_ext == int(g[f'rand{seed}'].max()) + 1
    for n, seed, minfasta in ((1500, 5, 20000), (2500, 6, 200000)):
        labels, lengths, markers = oi.synth_case(n, seed)
        lab, _ = integrate_labels(labels, lengths, markers, minfasta)
        np.testing.assert_array_equal(lab, g[f'synth{n}'])


def test_integrate_reference_objects_and_edge_cases():
    from collections import defaultdict
    from oracle import integrate as oi
    from semibin_b200.long_read import integrate_bins, integrate_labels
    labels, lengths, markers, minfasta = oi.random_case(7)
    names = [f'c{i}' for i in range(labels.shape[1])]
    cd = {nm: 'A' * int(l) for nm, l in zip(names, lengths)}
    c2m = defaultdict(list, {nm: [f'm{j}' for j in m] for nm, m in zip(names, markers) if len(m)})
    ext = integrate_bins([l.tolist() for l in labels], c2m, names, cd, minfasta)
    ref, _ = oi.integrate_restated(labels, lengths, markers, minfasta)
    got = np.full(len(names), -1)
    for b, members in enumerate(ext):
        assert members == sorted(members, key=names.index)
        got[[names.index(m) for m in members]] = b
    np.testing.assert_array_equal(got, ref)
    lab, n = integrate_labels(np.zeros((2, 3), np.int64), [10, 10, 10], [np.array([0])] * 3, 100)
    assert n == 0 and (lab == -1).all()
    lab, n = integrate_labels(np.zeros((2, 1), np.int64), [500], [np.array([], np.int32)], 100)
    assert n == 1 and lab[0] == 0
    lab, n = integrate_labels(np.zeros((1, 4), np.int64), [500] * 4, [np.array([], np.int32)] * 4, 100)
    assert n == 0


def test_integrate_20k_against_restatement():
    """Full hot path at 20k contigs: GPU kNN + sweep + integration vs the CPU restatement on the same labels."""
    from oracle import integrate as oi
    from semibin_b200.long_read import long_read_labels, integrate_labels
    n = 20000
    X, L = synth.long_read_case(n, n_sample=1, seed=77)
    _, genome = synth.embeddings(n, 77, return_genome=True)
    markers = synth.marker_table(genome, L, seed=77)
    labs = long_read_labels(X, None, L, n_sample=1, is_combined=True, n_neighbors=100)
    labels = np.stack(labs)
    got, n_got = integrate_labels(labels, L, markers, 200000)
    ref, n_ref = oi.integrate_restated(labels, L, markers, 200000)
    assert n_got == n_ref and n_ref > 20
    np.testing.assert_array_equal(got, ref)
