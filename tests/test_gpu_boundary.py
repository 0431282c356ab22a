"""GPU tier: the drop-in ENTRY POINTS themselves on hardware -- run_embed_infomap (SemiBin/cluster.py:64-168),
cluster_long_read (long_read_cluster.py:51-161) and patch.install() -- the boundary the reference pins with
test/test_bin.py:77-95.  igraph and the SemiBin package are not on the GPU box: a stub `igraph` captures hand-off #1,
a minimal fake `SemiBin` package provides the four helpers the long-read drop-in calls back into (marker search and
bin writing stay the reference's job); the expected values come from the committed goldens (outputs of the unmodified
reference) and from the oracle."""
import logging
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import knn as ok, edges as oe, dbscan as od, integrate as oi, compare
from semibin_b200 import synth

pytestmark = pytest.mark.gpu


class IdentityModel:
    """Stands in for the trained network (north_star: the MLP forward stays in the reference): returns the rows of a
    fixed embedding, on the device it is asked for."""

    def __init__(self, emb):
        self.emb = emb

    def eval(self):
        return self

    def embedding(self, x):
        import torch
        assert x.shape[0] == self.emb.shape[0]
        return torch.from_numpy(self.emb).to(x.device)


def _stub_igraph(monkeypatch, cap, labels_fn):
    mod = types.ModuleType('igraph')

    class Graph:
        def add_vertices(self, v):
            cap['n_vertices'] = len(v)

        def add_edges(self, es):
            cap['edges'] = np.asarray(es).reshape(-1, 2).copy()

        def community_infomap(self, edge_weights=None, vertex_weights=None, trials=None):
            cap['edge_weights'] = np.asarray(edge_weights).copy()
            cap['vertex_weights'] = np.asarray(vertex_weights).copy()
            cap['trials'] = trials
            return labels_fn(cap['n_vertices'])
    mod.Graph = Graph
    monkeypatch.setitem(sys.modules, 'igraph', mod)


@pytest.mark.parametrize('device', ['cuda', 'cpu'])
def test_run_embed_infomap_bin_data(golden, monkeypatch, device):
    """The reference's own fixture through the drop-in entry point: the hand-off igraph receives equals the one
    captured from the unmodified run_embed_infomap (tests/golden/bin_data.npz), the labels are mapped as
    cluster.py:164-167 does, and the embedding comes back unchanged.  device='cpu' is the reference test's own
    argument (test/test_bin.py:32): it only governs the network, the graph is still built on the GPU."""
    import pandas as pd
    from semibin_b200.cluster import run_embed_infomap
    g = golden('bin_data')
    n = 40
    names = [f'contig_{i}' for i in range(n)]
    data = pd.DataFrame(g['values'], index=names)
    contig_dict = {nm: 'A' * int(l) for nm, l in zip(names, g['lengths'])}
    cap = {}
    _stub_igraph(monkeypatch, cap, lambda nv: [list(range(0, nv, 4)), list(range(1, nv, 4)), list(range(2, nv, 4)), list(range(3, nv, 4))])
    emb, labels = run_embed_infomap(logging.getLogger('t'), IdentityModel(g['embedding']), data, device=device,
                                    max_edges=20, max_node=1, is_combined=bool(g['is_combined']),
                                    n_sample=int(g['n_sample']), contig_dict=contig_dict, num_process=1, random_seed=None)
    np.testing.assert_array_equal(emb, g['embedding'])
    assert emb.dtype == np.float32 and labels.shape == (n,)
    np.testing.assert_array_equal(labels, np.arange(n) % 4)                 # cluster.py:164-167
    depth = g['values'][:, 136:].astype(np.float32)
    S = int(g['n_sample'])
    res = compare.assert_edges_parity((g['edges_k20_X'], g['edges_k20_Y'], g['edges_k20_V']),
                                      (cap['edges'][:, 0], cap['edges'][:, 1], cap['edge_weights']),
                                      bound_fn=lambda x, y: oe.kl_weight_bound(depth, S, x, y))
    assert res['n_only_ref'] <= 1 and res['n_only_got'] <= 1, res           # duplicated k-mer row (d == 0 tie class)
    np.testing.assert_array_equal(cap['vertex_weights'], g['vertex_weights'])
    assert cap['trials'] == 10 and cap['n_vertices'] == n


def _fake_semibin(monkeypatch, markers_by_name, written):
    """Minimal `SemiBin` package: what semibin_b200.long_read / patch call back into."""
    import pandas as pd
    pkg = types.ModuleType('SemiBin'); pkg.__path__ = []
    cl = types.ModuleType('SemiBin.cluster')
    lr = types.ModuleType('SemiBin.long_read_cluster')
    ut = types.ModuleType('SemiBin.utils')
    mk = types.ModuleType('SemiBin.markers')
    cl.run_embed_infomap = lambda *a, **k: 'reference run_embed_infomap'
    cl.recluster_bins = lambda *a, **k: 'reference recluster_bins'
    lr.kneighbors_graph = lambda *a, **k: 'reference kneighbors_graph'
    lr.dbscan = lambda *a, **k: 'reference dbscan'
    lr.cluster_long_read = lambda *a, **k: 'reference cluster_long_read'

    def norm_abundance(data):                               # SemiBin/utils.py:630-636
        n = data.shape[1] - 136
        return bool(n >= 20 or (n >= 5 and np.mean(np.sum(data[:, 136:], axis=1)) > 2))

    def write_bins(namelist, contig_labels, output, contig_dict, *, minfasta, output_tag=None, output_compression='none'):
        written['labels'] = list(contig_labels)
        bins = sorted(set(l for l in contig_labels if l >= 0))
        return pd.DataFrame({'filename': [f'bin.{b}.fa' for b in bins], 'nbps': [0] * len(bins), 'nr_contigs': [0] * len(bins)})

    def estimate_seeds(fasta, binned_length, num_process, output=None, orf_finder=None, prodigal_output_faa=None):
        open(os.path.join(output, 'markers.hmmout'), 'wt').close()

    def get_marker(hmmout, orf_finder=None, min_contig_len=None, fasta_path=None, contig_to_marker=False):
        assert contig_to_marker
        return markers_by_name
    ut.norm_abundance, ut.write_bins = norm_abundance, write_bins
    mk.estimate_seeds, mk.get_marker = estimate_seeds, get_marker
    pkg.cluster, pkg.long_read_cluster, pkg.utils, pkg.markers = cl, lr, ut, mk
    for name, m in (('SemiBin', pkg), ('SemiBin.cluster', cl), ('SemiBin.long_read_cluster', lr), ('SemiBin.utils', ut),
                    ('SemiBin.markers', mk)):
        monkeypatch.setitem(sys.modules, name, m)
    return pkg


def test_cluster_long_read_entry_point(tmp_path, monkeypatch):
    """cluster_long_read(...) with the reference signature: the bins written (contig_bins.tsv, the labels handed to
    write_bins) equal the restatement of long_read_cluster.py:72-150 on the same inputs: features, kNN graph (k =
    min(200, N-1)), 12 weighted DBSCAN runs, greedy marker-gene integration."""
    import pandas as pd
    from collections import defaultdict
    n, S = 3000, 1
    emb, feat, depth, L = synth.short_read_case(n, n_sample=S, seed=77)
    _, genome = synth.embeddings(n, 77, return_genome=True)
    marks = synth.marker_table(genome, L, seed=77)
    names = [f'c{i}' for i in range(n)]
    table = np.concatenate([feat, depth.astype(np.float64)], axis=1)
    data = pd.DataFrame(table, index=names)
    contig_dict = {nm: 'A' * int(l) for nm, l in zip(names, L)}
    c2m = defaultdict(list, {nm: [f'm{j}' for j in m] for nm, m in zip(names, marks) if len(m)})
    written = {}
    _fake_semibin(monkeypatch, c2m, written)
    from semibin_b200.long_read import cluster_long_read
    out = str(tmp_path)
    args = SimpleNamespace(num_process=1, orf_finder='prodigal', prodigal_output_faa=None, output_tag='t', output_compression='none')
    minfasta = 200000
    cluster_long_read(logging.getLogger('t'), IdentityModel(emb), data, 'cuda', False, S, out, contig_dict,
                      binned_length=1000, args=args, minfasta=minfasta)
    # expected: the oracle restatement of the same steps
    x = np.concatenate((emb, np.log(np.clip(depth[:, [0]], 1e-6, None))), axis=1)      # long_read_cluster.py:75-81
    assert x.dtype == np.float32
    d, i = ok.knn_restated(x, min(200, n - 1))
    labels = np.stack([od.dbscan_minprop(d, i, eps, 5, L)[1] for eps in od.EPS_GRID])
    ref, n_ref = oi.integrate_restated(labels, L, marks, minfasta)
    got = pd.read_csv(os.path.join(out, 'contig_bins.tsv'), sep='\t')
    assert list(got['contig']) == names
    np.testing.assert_array_equal(got['bin'].values, ref)
    np.testing.assert_array_equal(np.array(written['labels']), ref)
    assert n_ref > 3 and os.path.exists(os.path.join(out, 'bins_info.tsv')) and os.path.isdir(os.path.join(out, 'output_bins'))


def test_patch_install_serves_reference_call_sites(golden, monkeypatch):
    """patch.install(): the names the reference's cluster_long_read body resolves (long_read_cluster.py:108-120) are
    bound to the GPU drop-ins; the 12 sequential dbscan() calls on one graph are served from ONE sweep on the graph
    still resident on the device, a NEW graph of the same size is never served stale labels, and
    uninstall() restores the originals."""
    import semibin_b200.patch as patch
    from semibin_b200 import neighbors
    g = golden('bin_data')
    pkg = _fake_semibin(monkeypatch, {}, {})
    rl, rc = pkg.long_read_cluster, pkg.cluster
    patch.install()
    try:
        assert rc.run_embed_infomap.__module__ == 'semibin_b200.cluster'
        assert rc.recluster_bins.__module__ == 'semibin_b200.recluster'
        assert rl.cluster_long_read.__module__ == 'semibin_b200.long_read'
        L = g['lengths']
        dist_matrix = rl.kneighbors_graph(g['long_emb_new'], n_neighbors=min(200, 39), mode='distance', p=2, n_jobs=4)
        assert neighbors.device_graph_of(dist_matrix) is not None
        compare.assert_knn_parity(g['long_d'], g['long_i'], dist_matrix.data.reshape(40, 39), dist_matrix.indices.reshape(40, 39),
                                  rtol=1e-7, atol=1e-6)
        for e, eps in enumerate(od.EPS_GRID):                                   # :115-120
            _, labels = rl.dbscan(dist_matrix, eps=eps, min_samples=5, n_jobs=4, metric='precomputed', sample_weight=L)
            np.testing.assert_array_equal(labels, g['long_labels'][e])
        # second sample, same shape: other data, other labels -- must not hit the cache of the first
        X2 = synth.long_read_case(40, n_sample=1, seed=5)[0]
        m2 = rl.kneighbors_graph(X2, n_neighbors=39, mode='distance', p=2, n_jobs=4)
        d2, i2 = ok.knn_restated(X2, 39)
        for eps in (0.3, 0.55):
            _, labels = rl.dbscan(m2, eps=eps, min_samples=5, n_jobs=4, metric='precomputed', sample_weight=L)
            np.testing.assert_array_equal(labels, od.dbscan_minprop(d2, i2, eps, 5, L)[1])
    finally:
        patch.uninstall()
    assert rc.run_embed_infomap() == 'reference run_embed_infomap' and rl.dbscan() == 'reference dbscan'
    assert rl.kneighbors_graph() == 'reference kneighbors_graph' and rl.cluster_long_read() == 'reference cluster_long_read'
