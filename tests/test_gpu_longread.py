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
