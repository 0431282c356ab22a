"""Generate tests/golden/*.npz from the LIVE reference (TEST ORACLE tooling).

Run in the build container only (needs /root/reference and scikit-learn 1.9.0):
    python -m oracle.make_golden
Every array below is an output of the unmodified reference code path
(SemiBin.cluster.run_embed_infomap / cal_kl, sklearn kneighbors_graph / dbscan)
on inputs that are either the reference's own fixture test/bin_data or seeded
synthetic arrays from semibin_b200.synth (regenerated from the seed at test
time, so only outputs are stored).
"""
import os
import sys
import warnings
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import knn, reference_calls as rc          # noqa: E402
from semibin_b200 import synth                          # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
REF = rc.REFERENCE_ROOT


def _knn(X, k):
    d, i = knn.knn_sklearn(X, k)
    return d, i.astype(np.int32)


def bin_data():
    """The reference's fixture (test/test_bin.py:77-95 inputs)."""
    import torch
    sys.path.insert(0, REF)
    from SemiBin.main import binning_preprocess
    from SemiBin.fasta import fasta_iter
    cwd = os.getcwd(); os.chdir(REF)
    try:
        contig_dict = {h: seq for h, seq in fasta_iter('test/bin_data/input.fasta')}
        is_combined, n_sample, data, model = binning_preprocess(
            data='test/bin_data/data.csv', depth_metabat2=None,
            model_path='test/bin_data/model.pt', environment=None, device='cpu')
    finally:
        os.chdir(cwd)
    out = dict(values=data.values.copy(), is_combined=is_combined, n_sample=n_sample,
               lengths=np.array([len(contig_dict[n]) for n in data.index], np.int64))
    for k in (20, 39):
        X, Y, V, emb, cap = rc.capture_edge_handoff(model, data, max_edges=k, max_node=1,
                                                    is_combined=is_combined, n_sample=n_sample,
                                                    contig_dict=contig_dict)
        out[f'edges_k{k}_X'] = X; out[f'edges_k{k}_Y'] = Y; out[f'edges_k{k}_V'] = V
        out['embedding'] = emb
        out['vertex_weights'] = cap['vertex_weights']
        d, i = _knn(emb, k); out[f'emb_k{k}_d'] = d; out[f'emb_k{k}_i'] = i
        d, i = _knn(np.ascontiguousarray(data.values[:, :136]), k)
        out[f'kmer_k{k}_d'] = d; out[f'kmer_k{k}_i'] = i
    # long-read path (long_read_cluster.py:75-81,108-120) on the same fixture
    depth = data.values[:, 136:].astype(np.float32)[:, [0]]
    emb_new = np.concatenate((out['embedding'], np.log(np.clip(depth, 1e-6, None))), axis=1)
    G, labels = rc.reference_dbscan_sweep(emb_new, out['lengths'], k=200)
    out['long_emb_new'] = emb_new
    out['long_d'] = G.data.reshape(40, 39); out['long_i'] = G.indices.reshape(40, 39).astype(np.int32)
    out['long_labels'] = np.stack(labels).astype(np.int64)
    np.savez_compressed(os.path.join(OUT, 'bin_data.npz'), **out)


def ties():
    out = {}
    for name, kw, dt in [('grid', dict(n=700, n_dup=0), np.float32),
                         ('dup5', dict(n=700, n_dup=5), np.float32),
                         ('dup40', dict(n=700, n_dup=40), np.float32),     # > k duplicates corner
                         ('grid64', dict(n=700, n_dup=5), np.float64)]:
        X = synth.tie_grid(dtype=dt, **kw)
        d, i = _knn(X, 20)
        out[name + '_d'] = d; out[name + '_i'] = i
    np.savez_compressed(os.path.join(OUT, 'ties.npz'), **out)


def synth_knn():
    out = {}
    X = synth.embeddings(4096, seed=101)
    d, i = _knn(X, 50); out['emb4096_k50_d'] = d.astype(np.float32); out['emb4096_k50_i'] = i
    F = synth.kmer_features(4096, seed=101)
    d, i = _knn(F, 50); out['kmer4096_k50_d'] = d.astype(np.float32); out['kmer4096_k50_i'] = i
    X = synth.embeddings(1500, seed=102)
    d, i = _knn(X, 200); out['emb1500_k200_d'] = d.astype(np.float32); out['emb1500_k200_i'] = i
    X = synth.embeddings(300, seed=103)                  # k = N-1
    d, i = _knn(X, 299); out['emb300_k299_d'] = d.astype(np.float32); out['emb300_k299_i'] = i
    Xl, _ = synth.long_read_case(3000, n_sample=3, seed=104)   # D = 103 (odd width)
    d, i = _knn(Xl, 64); out['long3000_k64_d'] = d.astype(np.float32); out['long3000_k64_i'] = i
    out = {k: (v.astype(np.uint16) if k.endswith('_i') else v) for k, v in out.items()}
    np.savez_compressed(os.path.join(OUT, 'synth_knn.npz'), **out)


def combined_features(feat, depth):
    cov = depth[:, 0::2].astype(np.float64)
    return np.concatenate([feat, cov / 100.0], axis=1)


def synth_edges():
    out = {}
    for S, comb in [(1, False), (3, False), (10, False), (10, True)]:
        emb, feat, depth, L = synth.short_read_case(3000, n_sample=S, seed=200 + S)
        f = combined_features(feat, depth) if comb else feat
        for max_node in (1, 0.6):
            X, Y, V = rc.reference_edge_list(emb, f, depth, max_edges=50, max_node=max_node,
                                             is_combined=comb, n_sample=S)
            tag = f'S{S}_{"comb" if comb else "kl"}_mn{max_node}'
            out[tag + '_X'] = X; out[tag + '_Y'] = Y; out[tag + '_V'] = V
    np.savez_compressed(os.path.join(OUT, 'synth_edges.npz'), **out)


def synth_dbscan():
    out = {}
    X, L = synth.long_read_case(4000, n_sample=2, seed=300)
    G, labels = rc.reference_dbscan_sweep(X, L, k=60)
    out['w_labels'] = np.stack(labels).astype(np.int64)
    # unit weights -> non-core points and noise exist
    from oracle import dbscan as od
    d = G.data.reshape(4000, 60); i = G.indices.reshape(4000, 60)
    labs = []; ncore = []
    for eps in od.EPS_GRID:
        c, lab = od.dbscan_sklearn(d, i, eps, 5, None)
        labs.append(lab); ncore.append(len(c))
    out['u_labels'] = np.stack(labs).astype(np.int64); out['u_ncore'] = np.array(ncore)
    np.savez_compressed(os.path.join(OUT, 'synth_dbscan.npz'), **out)


def cal_kl():
    rng = np.random.default_rng(400)
    out = {}
    for t, scale in enumerate((1.0, 30.0)):
        m = (rng.random(128) * scale).astype(np.float32)
        v = (rng.random(128) * scale * 3).astype(np.float32)
        out[f'm{t}'] = m; out[f'v{t}'] = v
        out[f'kl{t}'] = rc.reference_cal_kl(m, v)
    np.savez_compressed(os.path.join(OUT, 'cal_kl.npz'), **out)


def integrate():
    """long_read_cluster.py:12-49,122-150 run with the reference's own get_best_bin."""
    from collections import defaultdict
    from oracle import integrate as oi
    out = {}

    def run(labels, lengths, markers, minfasta):
        n = labels.shape[1]
        names = [f'c{i}' for i in range(n)]
        cd = {nm: 'A' * int(l) for nm, l in zip(names, lengths)}
        c2m = defaultdict(list, {nm: [f'm{j}' for j in m] for nm, m in zip(names, markers) if len(m)})
        ext, lab = oi.reference_loop([l.tolist() for l in labels], c2m, names, cd, minfasta)
        return np.asarray(lab, dtype=np.int64)

    for seed in range(24):
        labels, lengths, markers, minfasta = oi.random_case(seed)
        out[f'rand{seed}'] = run(labels, lengths, markers, minfasta)
    for n, seed, minfasta in ((1500, 5, 20000), (2500, 6, 200000)):
        labels, lengths, markers = oi.synth_case(n, seed)
        out[f'synth{n}'] = run(labels, lengths, markers, minfasta)
    np.savez_compressed(os.path.join(OUT, 'integrate.npz'), **out)


def kmeans_cases():
    """Live scikit-learn KMeans exactly as the reference calls it (cluster.py:236-241)."""
    from sklearn.cluster import KMeans
    out = {}
    cases = dict(separated=(11, 1500, 103, 4, 0.05, 0.5, 0), overlapping=(12, 2000, 110, 6, 0.35, 0.06, 0),
                 many=(13, 1800, 101, 12, 0.2, 0.05, 0), dupseeds=(14, 900, 103, 3, 0.15, 0.5, 2),
                 tiny=(15, 40, 103, 3, 0.3, 0.3, 0))
    for name, (seed, n, d, K, spread, sep, dup) in cases.items():
        X, w, seeds = synth.recluster_problem(seed, n, d, K, spread, sep, dup)
        km = KMeans(n_clusters=len(seeds), init=X[seeds], n_init=1, random_state=0)
        km.fit(X, sample_weight=w)
        out[f'{name}_X'] = X; out[f'{name}_w'] = w; out[f'{name}_seeds'] = seeds
        out[f'{name}_labels'] = km.labels_.astype(np.int32); out[f'{name}_centers'] = km.cluster_centers_
        out[f'{name}_n_iter'] = np.int64(km.n_iter_)
    np.savez_compressed(os.path.join(OUT, 'kmeans_cases.npz'), **out)


if __name__ == '__main__':
    warnings.simplefilter('ignore')
    os.makedirs(OUT, exist_ok=True)
    import sklearn
    assert sklearn.__version__.startswith('1.9'), sklearn.__version__
    only = sys.argv[1:]
    for fn in (bin_data, ties, synth_knn, synth_edges, synth_dbscan, cal_kl, integrate, kmeans_cases):
        if only and fn.__name__ not in only:
            continue
        fn(); print('wrote', fn.__name__)
