"""Build-container tier (needs /root/reference; skipped elsewhere): the host-side glue of the cluster_long_read
drop-in -- network input, label mapping, files written -- against the UNMODIFIED reference function on its own
fixture test/bin_data.  The GPU stages are replaced by the CPU oracle here (their parity is the -m gpu tier's
job); the marker-gene tools (prodigal / hmmsearch, absent) are stubbed identically for both sides."""
import argparse
import logging
import os
import sys
from collections import defaultdict

import numpy as np
import pytest

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'SemiBin')), reason='reference tree not present')


def _fake_markers(names):
    rng = np.random.default_rng(3)
    table = defaultdict(list)
    for name in names:
        for m in np.unique(rng.integers(0, 30, rng.integers(0, 4))):
            table[name].append(f'marker{m}')
    return table


def test_cluster_long_read_files_match_reference(tmp_path, monkeypatch):
    sys.path.insert(0, REF)
    pd = pytest.importorskip('pandas')
    import SemiBin.long_read_cluster as rl
    import SemiBin.markers as rm
    from SemiBin.main import binning_preprocess
    from SemiBin.fasta import fasta_iter
    from semibin_b200 import long_read as ours
    from oracle import dbscan as od, knn as ok, integrate as oi

    cwd = os.getcwd()
    os.chdir(REF)
    try:
        contig_dict = {h: seq for h, seq in fasta_iter('test/bin_data/input.fasta')}
        is_combined, n_sample, data, model = binning_preprocess(
            data='test/bin_data/data.csv', depth_metabat2=None, model_path='test/bin_data/model.pt',
            environment=None, device='cpu')
    finally:
        os.chdir(cwd)
    names = list(data.index)
    markers = _fake_markers(names)
    for mod in (rl, rm):
        monkeypatch.setattr(mod, 'estimate_seeds', lambda *a, **k: None)
        monkeypatch.setattr(mod, 'get_marker', lambda *a, **k: markers)
    args = argparse.Namespace(num_process=1, orf_finder='prodigal', prodigal_output_faa=None, output_tag=None,
                              output_compression='none')

    ref_out = str(tmp_path / 'ref')
    os.makedirs(ref_out)
    rl.cluster_long_read(logging, model, data, 'cpu', is_combined, n_sample, ref_out, contig_dict,
                         binned_length=1000, args=args, minfasta=0)

    def cpu_labels(embedding, depth, lengths, *, n_sample, is_combined, **kw):
        x = ours.long_read_features(embedding, depth, n_sample, is_combined)
        d, i = ok.knn_sklearn(x, min(200, x.shape[0] - 1))
        return [od.dbscan_minprop(d, i, eps, 5, lengths)[1] for eps in ours.EPS_GRID]

    def cpu_integrate(results, contig2marker, contig_list, cdict, minfasta):
        ids = {}
        per = [np.array(sorted({ids.setdefault(m, len(ids)) for m in contig2marker.get(n, [])}), dtype=np.int64)
               for n in contig_list]
        lab, n_ext = oi.integrate_restated(np.stack(results), [len(cdict[n]) for n in contig_list], per, minfasta)
        ext = [[] for _ in range(n_ext)]
        for n, l in zip(contig_list, lab):
            if l >= 0:
                ext[l].append(n)
        return ext

    monkeypatch.setattr(ours, 'long_read_labels', cpu_labels)
    monkeypatch.setattr(ours, 'integrate_bins', cpu_integrate)
    our_out = str(tmp_path / 'ours')
    os.makedirs(our_out)
    ours.cluster_long_read(logging, model, data, 'cpu', is_combined, n_sample, our_out, contig_dict,
                           binned_length=1000, args=args, minfasta=0)

    a = pd.read_table(os.path.join(ref_out, 'contig_bins.tsv'))
    b = pd.read_table(os.path.join(our_out, 'contig_bins.tsv'))
    pd.testing.assert_frame_equal(a, b)
    assert (a['bin'] >= 0).any()
    ia = pd.read_table(os.path.join(ref_out, 'bins_info.tsv'))
    ib = pd.read_table(os.path.join(our_out, 'bins_info.tsv'))
    assert list(ia.columns) == list(ib.columns) and len(ia) == len(ib)
    num = [c for c in ia.columns if ia[c].dtype.kind in 'if']
    pd.testing.assert_frame_equal(ia[num], ib[num])
    assert sorted(os.listdir(os.path.join(ref_out, 'output_bins'))) == sorted(os.listdir(os.path.join(our_out, 'output_bins')))


def test_run_embed_infomap_handoff_matches_reference(monkeypatch):
    """The run_embed_infomap drop-in (input preparation, embedding forward, igraph hand-off, label vector) against
    the unmodified reference on test/bin_data and on a combined-mode frame, with a recording igraph stub on both
    sides and the CPU oracle standing in for the GPU edge-list build."""
    sys.path.insert(0, REF)
    import types
    from oracle import reference_calls as rc, knn as ok, edges as oe
    from semibin_b200 import cluster as ours, synth
    rc._import_reference()
    from SemiBin.main import binning_preprocess
    from SemiBin.fasta import fasta_iter

    def cpu_edge_list(embedding, features, depth, *, max_edges, max_node, is_combined, n_sample, **kw):
        emb = embedding.cpu().numpy() if hasattr(embedding, 'cpu') else np.asarray(embedding)
        k = min(int(max_edges), emb.shape[0] - 1)
        a_d, a_i = ok.knn_sklearn(emb, k)
        b_d, b_i = ok.knn_sklearn(np.ascontiguousarray(features), k)
        X, Y, V, _ = oe.edge_list_restated(a_d, a_i, b_d, b_i, depth, max_node=max_node, is_combined=is_combined,
                                           n_sample=n_sample)
        return X, Y, V

    monkeypatch.setattr(ours, 'build_edge_list', cpu_edge_list)

    def run_ours(model, data, **kw):
        cap = {}
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
                n = cap['n_vertices']
                return [list(range(0, n, 2)), list(range(1, n, 2))]
        mod.Graph = Graph
        monkeypatch.setitem(sys.modules, 'igraph', mod)
        emb, labels = ours.run_embed_infomap(logging.getLogger('t'), model, data, num_process=1, random_seed=None, **kw)
        return emb, labels, cap

    cwd = os.getcwd()
    os.chdir(REF)
    try:
        contig_dict = {h: seq for h, seq in fasta_iter('test/bin_data/input.fasta')}
        is_combined, n_sample, data, model = binning_preprocess(
            data='test/bin_data/data.csv', depth_metabat2=None, model_path='test/bin_data/model.pt',
            environment=None, device='cpu')
    finally:
        os.chdir(cwd)
    kw = dict(device='cpu', max_edges=20, max_node=1, is_combined=is_combined, n_sample=n_sample, contig_dict=contig_dict)
    X, Y, V, ref_emb, ref_cap = rc.capture_edge_handoff(model, data, **{k: v for k, v in kw.items() if k != 'device'})
    emb, labels, cap = run_ours(model, data, **kw)
    np.testing.assert_array_equal(emb, ref_emb)
    np.testing.assert_array_equal(cap['edges'], np.stack([X, Y], axis=1))
    np.testing.assert_array_equal(cap['edge_weights'], V)        # the oracle stands in for the GPU stages here: bit-equal
    np.testing.assert_array_equal(cap['vertex_weights'], ref_cap['vertex_weights'])
    assert cap['trials'] == ref_cap['trials'] == 10 and cap['n_vertices'] == ref_cap['n_vertices']
    assert labels.shape == (len(data),) and set(labels[0::2]) == {0} and set(labels[1::2]) == {1}

    # combined mode (S = 10 >= 5: column- then row-normalised input, no KL term), driven through the frame
    import pandas as pd
    emb_s, feat, depth, L = synth.short_read_case(600, n_sample=10, seed=21)
    vals = np.concatenate([feat, depth[:, 0::2].astype(np.float64) / 100.0], axis=1)
    frame = pd.DataFrame(vals, index=[f'c{i}' for i in range(600)])
    cdict = {f'c{i}': 'A' * int(l) for i, l in enumerate(L)}
    kw = dict(max_edges=30, max_node=1, is_combined=True, n_sample=10, contig_dict=cdict)
    X, Y, V, _, ref_cap = rc.capture_edge_handoff(rc._IdentityModel(emb_s), frame, **kw)
    _, _, cap = run_ours(rc._IdentityModel(emb_s), frame, device='cpu', **kw)
    np.testing.assert_array_equal(cap['edges'], np.stack([X, Y], axis=1))
    np.testing.assert_array_equal(cap['edge_weights'], V)        # the oracle stands in for the GPU stages here: bit-equal
    np.testing.assert_array_equal(cap['vertex_weights'], ref_cap['vertex_weights'])


def test_recluster_bins_matches_reference(monkeypatch):
    """Row N4: semibin_b200.recluster.recluster_bins against the UNMODIFIED SemiBin.cluster.recluster_bins (live
    scikit-learn KMeans per bin) on a synthetic sample: depth scaling of the features (:185-195), the concatenated-bin
    fasta handed to the seed search (:202-214), which bins are split, label bookkeeping (:221-253).  The seed search
    (ORF finder + hmmsearch, absent) is stubbed identically on both sides; the batched GPU launch is replaced by the
    CPU oracle here (its parity with scikit-learn is tests/test_oracle.py and the -m gpu tier)."""
    import pandas as pd
    sys.path.insert(0, REF)
    import SemiBin.cluster as ref_cluster
    import semibin_b200.recluster as ours
    from semibin_b200 import synth
    from oracle import kmeans as okm
    rng = np.random.default_rng(9)
    n_sample = 3
    embs, bins, names = [], [], []
    seeds = {}
    for b in range(12):
        n = int(rng.integers(4, 400)); K = int(rng.integers(1, 5))
        X, w, s = synth.recluster_problem(500 + b, max(n, K + 1), 100, K, 0.15, 0.3)
        nm = [f'contig_{b}_{i}' for i in range(len(X))]
        embs.append(X); bins += [b] * len(X); names += nm
        if b != 5:
            seeds[f'bin{b:06}'] = [nm[i] for i in s]
    emb = np.concatenate(embs)
    perm = rng.permutation(len(emb))
    emb = emb[perm]; bins = np.asarray(bins)[perm]; names = [names[i] for i in perm]
    lengths = rng.integers(1000, 30000, len(emb))
    contig_dict = {nm: 'A' * int(ell) for nm, ell in zip(names, lengths)}
    depth = np.abs(rng.normal(40, 25, (len(emb), 2 * n_sample)))
    data = pd.DataFrame(np.concatenate([rng.random((len(emb), 136)), depth], axis=1), index=names)
    total = {b: int(lengths[bins == b].sum()) for b in range(12)}
    minfasta = int(sorted(total.values())[2]) + 1              # the two smallest bins stay below minfasta
    seen = {}

    def fake_estimate_seeds(fasta, binned_length, num_process, multi_mode=False, orf_finder=None):
        assert multi_mode
        seen.setdefault('fasta', []).append(open(fasta).read())
        return seeds
    monkeypatch.setattr(ref_cluster, 'estimate_seeds', fake_estimate_seeds)
    kw = dict(n_sample=n_sample, embedding=emb, is_combined=False, contig_labels=bins.copy(), minfasta=minfasta,
              contig_dict=contig_dict, binned_length=1000, orf_finder='prodigal', num_process=1, random_seed=7)
    ref = ref_cluster.recluster_bins(logging.getLogger('t'), data, **kw)

    def cpu_problems(embedding_new, rows, po, seed_rows, co, lengths_, device=None):
        out = np.empty(len(rows), np.int32)
        for t in range(len(po) - 1):
            r = rows[po[t]:po[t + 1]]
            out[po[t]:po[t + 1]], _, _ = okm.lloyd_seeded(embedding_new[r], np.asarray(lengths_)[r], embedding_new[seed_rows[co[t]:co[t + 1]]])
        return out
    monkeypatch.setattr(ours, '_kmeans_problems', cpu_problems)
    got = ours.recluster_bins(logging.getLogger('t'), data, **kw, estimate_seeds=fake_estimate_seeds)
    np.testing.assert_array_equal(np.asarray(ref), got)
    assert seen['fasta'][0] == seen['fasta'][1]                # the same concatenated fasta went to the seed search
    assert got.max() + 1 > 12                                  # some bins were split
    # no seeds at all: the labels come back untouched, as :215-217
    monkeypatch.setattr(ref_cluster, 'estimate_seeds', lambda *a, **k: {})
    r0 = ref_cluster.recluster_bins(logging.getLogger('t'), data, **kw)
    g0 = ours.recluster_bins(logging.getLogger('t'), data, **kw, estimate_seeds=lambda *a, **k: {})
    np.testing.assert_array_equal(np.asarray(r0), g0)
