"""GPU tier: parity at (or near) the BASELINE config sizes against the REFERENCE's own code on the box's host
cores -- live scikit-learn kneighbors / dbscan -- and the oracle restatement fed with scikit-learn's graphs, never
with the GPU's own.  Sampled rows keep the CPU side to seconds."""
import numpy as np
import pytest

from oracle import knn as ok, edges as oe, dbscan as od, compare
from semibin_b200 import synth

pytestmark = pytest.mark.gpu


def _sk_rows(X, k, rows):
    """kneighbors_graph rows of live scikit-learn (brute force, self dropped by index, neighbors/_base.py:943-957)."""
    from sklearn.neighbors import NearestNeighbors
    nn = NearestNeighbors(n_neighbors=k + 1, algorithm='brute').fit(X)
    rd, ri = nn.kneighbors(X[rows])
    out_d = np.empty((len(rows), k)); out_i = np.empty((len(rows), k), np.int64)
    for t, r in enumerate(rows):
        m = ri[t] != r
        if m.all():
            m[0] = False
        out_d[t], out_i[t] = rd[t][m], ri[t][m]
    return out_d, out_i


def _combined_features(feat, depth):
    f = np.concatenate([feat, depth[:, 0::2].astype(np.float64) / 100.0], axis=1)
    n = f.shape[1] - 136
    if n >= 20 or (n >= 5 and np.mean(np.sum(f[:, 136:], axis=1)) > 2):      # utils.py:630-636, cluster.py:83-86
        f = f / np.sum(f, axis=0)
        f = f / np.abs(f).sum(axis=1, keepdims=True)
    return np.ascontiguousarray(f)


@pytest.mark.parametrize('width', [136, 146])
def test_feature_graph_200k_vs_live_sklearn(width):
    """a4: float64 feature graph (136 k-mer columns, 146 combined) at 200k through algo='auto' (the tensor path with
    its certificate) against live scikit-learn on 512 sampled rows."""
    import torch
    from semibin_b200.neighbors import knn
    N, k = 200000, 200
    emb, feat, depth, L = synth.short_read_case(N, n_sample=10, seed=20242)
    F = feat if width == 136 else _combined_features(feat, depth)
    assert F.shape[1] == width and F.dtype == np.float64
    d, i, st = knn(torch.from_numpy(F).cuda(), k, algo='auto', return_stats=True)
    assert st['algo_used'] == 2
    d = d.cpu().numpy(); i = i.cpu().numpy().astype(np.int64)
    rows = np.sort(np.random.default_rng(11).choice(N, 512, replace=False))
    rd, ri = _sk_rows(F, k, rows)
    res = compare.assert_knn_parity(rd, ri, d[rows], i[rows], rtol=1e-6)
    print(width, res, {kk: st[kk] for kk in ('n_fallback_rows', 'n_retry_rows', 'n_fail_few', 'n_fail_certificate', 'n_split_cols', 'kpad')})
    assert st['n_fallback_rows'] < 0.01 * N, st
    assert (np.diff(d, axis=1) >= 0).all() and (i != np.arange(N)[:, None]).all() and i.min() >= 0 and i.max() < N


def test_feature_graph_450k_symmetric_vs_live_sklearn():
    """a4 above the size from which a build of all rows takes the symmetric filter: combined features (D = 146, float64,
    split precision columns) at 450k through algo='auto', 384 sampled rows against live scikit-learn."""
    import torch
    from semibin_b200.neighbors import knn
    N, k = 450000, 100
    emb, feat, depth, L = synth.short_read_case(N, n_sample=10, seed=20243)
    F = _combined_features(feat, depth)
    d, i, st = knn(torch.from_numpy(F).cuda(), k, algo='auto', return_stats=True)
    assert st['algo_used'] == 2 and st['symmetric'] == 1, st
    d = d.cpu().numpy(); i = i.cpu().numpy().astype(np.int64)
    rows = np.sort(np.random.default_rng(12).choice(N, 384, replace=False))
    rd, ri = _sk_rows(F, k, rows)
    res = compare.assert_knn_parity(rd, ri, d[rows], i[rows], rtol=1e-6)
    print(res, {kk: st[kk] for kk in ('n_fallback_rows', 'n_retry_rows', 'n_fail_few', 'n_fail_certificate', 'n_split_cols', 'kpad')})
    assert st['n_fallback_rows'] < 0.01 * N, st
    assert (np.diff(d, axis=1) >= 0).all() and (i != np.arange(N)[:, None]).all() and i.min() >= 0 and i.max() < N


def test_c4_labels_300k_vs_live_sklearn():
    """a13 + a14 at scale: long-read kNN graph (D = 101) and DBSCAN labels of 300k contigs; the labels of three eps
    values against live sklearn.cluster.dbscan on the graph of the GPU (the graph itself is checked on sampled rows
    against live scikit-learn first), weighted by contig length as long_read_cluster.py:117-119."""
    from semibin_b200.long_read import long_read_labels
    N, k = 300000, 200
    X, L = synth.long_read_case(N, n_sample=1, seed=20244)
    labs, (dist, idx) = long_read_labels(X, None, L, n_sample=1, is_combined=True, n_neighbors=k, return_graph=True)
    d = dist.cpu().numpy().astype(np.float64); i = idx.cpu().numpy().astype(np.int64)
    rows = np.sort(np.random.default_rng(12).choice(N, 256, replace=False))
    rd, ri = _sk_rows(X, k, rows)
    compare.assert_knn_parity(rd, ri, d[rows], i[rows], rtol=1e-6)
    for e in (1, 5, 11):
        _, ref = od.dbscan_sklearn(d, i, od.EPS_GRID[e], 5, L)
        np.testing.assert_array_equal(labs[e], ref)
    for lab in labs:                                           # label properties at full size
        u = np.unique(lab[lab >= 0])
        assert (u == np.arange(len(u))).all()


@pytest.mark.parametrize('combined', [False, True])
def test_c3_edges_200k_sampled_rows_vs_oracle(combined):
    """a5-a10 at 200k contigs x 10 samples, k = 200: the edge list of 512 sampled rows against the restatement fed
    with LIVE SCIKIT-LEARN graphs of those rows (not the GPU's), weights within the derived bound; row maxima equal,
    so the max_node threshold (a function of the row-maximum counts, replayed on the host) is pinned too."""
    _c3_sampled_rows(200000, combined)


def test_c3_edges_1M_sampled_rows_vs_oracle():
    """The same at the BASELINE size of C3 (1M contigs x 10 samples, KL weights): both kNN graphs come from the
    symmetric filter here, the 28.6 M-edge list is compared on 512 sampled rows."""
    _c3_sampled_rows(1000000, False)


def _c3_sampled_rows(N, combined):
    from semibin_b200.cluster import build_edge_list, threshold_sequence, pick_threshold
    k, S = 200, 10
    emb, feat, depth, L = synth.short_read_case(N, n_sample=S, seed=20243)
    F = _combined_features(feat, depth) if combined else feat
    max_node = 0.9
    (X, Y, V), gr = build_edge_list(emb, F, depth, max_edges=k, max_node=max_node, is_combined=combined, n_sample=S,
                                    return_graphs=True)
    rows = np.sort(np.random.default_rng(13).choice(N, 512, replace=False))
    ad, ai = _sk_rows(emb, k, rows)
    bd, bi = _sk_rows(F, k, rows)
    # the restatement on the sampled rows: intersection, w, row maximum
    rowmax_gpu = gr['rowmax'].cpu().numpy()
    thr = gr['threshold']
    counts = np.array([(rowmax_gpu > t).sum() for t in threshold_sequence()])
    assert pick_threshold(counts, N, max_node) == thr
    I_l, J_l, w_l = [], [], []
    for t, r in enumerate(rows):
        inb = np.isin(ai[t], bi[t][bd[t] != 0]) & (ad[t] != 0)                 # cluster.py:109-113
        w = 1 - np.minimum(ad[t][inb], 1.0)                                    # :115-116
        assert (w.max() if len(w) else 0.0) == rowmax_gpu[r], (r, w.max() if len(w) else 0.0, rowmax_gpu[r])
        J = ai[t][inb]
        keep = ~(w <= thr) & (w != 0)                                          # :127-128
        I_l.append(np.full(int(keep.sum()), r, np.int64)); J_l.append(J[keep]); w_l.append(w[keep])
    I = np.concatenate(I_l); J = np.concatenate(J_l); w = np.concatenate(w_l)
    if not combined:                                                           # :129-141
        acc = None
        for s in range(S):
            kl = oe.kl_edges(depth[:, 2 * s], depth[:, 2 * s + 1], I, J)
            if acc is None:
                kl *= -1
                acc = kl
            else:
                acc -= kl
        acc += S
        acc /= S
        w = w * acc.astype(np.float64)
    keep = ~(w <= 1e-6) & (J > I)                                              # :143-148
    I, J, w = I[keep], J[keep], w[keep]
    o = np.lexsort((J, I))
    RX, RY, RV = I[o], J[o], w[o]
    sel = np.isin(X, rows)
    res = compare.assert_edges_parity((RX, RY, RV), (X[sel], Y[sel], V[sel]), rtol=0.0 if combined else 1e-5,
                                      bound_fn=None if combined else (lambda x, y: oe.kl_weight_bound(depth, S, x, y)))
    print(combined, thr, res)
    assert res['n_common'] > 1000 and res['n_only_ref'] + res['n_only_got'] <= 2
    key = X.astype(np.int64) << 32 | Y
    assert (np.diff(key) > 0).all() and (Y > X).all() and (V > 1e-6).all() and (V <= 1).all()


def test_rerank_warp_and_block_paths_agree():
    """The warp-per-row rerank and the block-per-row kernel it hands wide rows to are interchangeable: k = 200
    (warp kernel, 256-element sort) and k = 300 (512-element sort) against the FP64 exact path at 60k, and a matrix
    with 400-row duplicate groups whose rows overflow the warp kernel's sort and take the retry path."""
    import torch
    from semibin_b200.neighbors import knn
    X = synth.embeddings(60000, seed=5)
    Xd = torch.from_numpy(X).cuda()
    for k in (200, 300, 31):
        d0, i0 = knn(Xd, k, algo='exact')
        d1, i1, st = knn(Xd, k, algo='tensor', return_stats=True)
        assert torch.equal(i0, i1) and torch.equal(d0, d1), (k, st)
    X2 = X.copy()
    for g in range(20):
        X2[1000 * g:1000 * g + 400] = X2[1000 * g]            # 400 identical rows: equal lower bounds, > 256 survivors
    Xd = torch.from_numpy(X2).cuda()
    d0, i0 = knn(Xd, 100, algo='exact')
    d1, i1, st = knn(Xd, 100, algo='tensor', return_stats=True)
    assert torch.equal(i0, i1) and torch.equal(d0, d1), st
    assert st['n_retry_rows'] + st['n_fallback_rows'] >= 20 * 400, st


def test_symmetric_block_row_build_streams_to_host():
    """Host-buffer builds of all rows through the symmetric filter run block row by block row: every block of rows is
    reranked and written straight into the (pinned, mapped) host arrays while the later block rows are multiplied.  Both
    host forms (the csr_matrix arrays of kneighbors_graph; the f32 / i32 buffers of knn_host), float32 and float64
    input, a tile count that is not a multiple of the 8 blocks, rows that need the exact fallback -- equal to the
    device-resident exact path."""
    import torch
    from semibin_b200.neighbors import knn, kneighbors_graph, knn_host
    N, k = 41000, 50
    X = synth.embeddings(N, seed=41)
    d0, i0 = knn(torch.from_numpy(X).cuda(), k, algo='exact')
    g, st = kneighbors_graph(X, k, algo='tensor_symmetric', return_stats=True)
    assert st['symmetric'] == 1 and st['n_chunks'] == 8, st
    np.testing.assert_array_equal(g.indices.reshape(N, k), i0.cpu().numpy())
    np.testing.assert_array_equal(g.data.reshape(N, k), d0.cpu().numpy().astype(np.float64))
    hd, hi, sth = knn_host(torch.from_numpy(X).pin_memory(), k, algo='tensor_symmetric', return_stats=True)
    assert sth['symmetric'] == 1
    np.testing.assert_array_equal(hi.numpy(), i0.cpu().numpy())
    np.testing.assert_array_equal(hd.numpy(), d0.cpu().numpy())
    # near-duplicate rows: lists overflow / certificates fail for some rows, which are recomputed and re-sent at the end
    Xd = X.copy(); Xd[5000:5400] = Xd[0] + 1e-7 * np.arange(400, dtype=np.float32)[:, None]
    d1, i1 = knn(torch.from_numpy(Xd).cuda(), k, algo='exact')
    g1, st1 = kneighbors_graph(Xd, k, algo='tensor_symmetric', return_stats=True)
    np.testing.assert_array_equal(g1.indices.reshape(N, k), i1.cpu().numpy())
    np.testing.assert_array_equal(g1.data.reshape(N, k), d1.cpu().numpy().astype(np.float64))
    F = synth.kmer_features(30000, seed=41)                      # float64: distances are downloaded as they are
    df, jf = knn(torch.from_numpy(F).cuda(), 20, algo='exact')
    gf, stf = kneighbors_graph(F, 20, algo='tensor_symmetric', return_stats=True)
    assert stf['symmetric'] == 1
    np.testing.assert_array_equal(gf.data.reshape(30000, 20), df.cpu().numpy())
    np.testing.assert_array_equal(gf.indices.reshape(30000, 20), jf.cpu().numpy())
    # pageable host outputs cannot be written by the device: the call falls back to the chunked copies
    pd_, pi_ = torch.empty((N, k), dtype=torch.float32), torch.empty((N, k), dtype=torch.int32)
    knn_host(torch.from_numpy(X).pin_memory(), k, algo='tensor_symmetric', out_dist=pd_, out_idx=pi_)
    np.testing.assert_array_equal(pi_.numpy(), i0.cpu().numpy())


def test_kneighbors_graph_host_arrays_and_devices():
    """The sklearn-signature drop-in: csr_matrix over pinned (float64, int64) arrays written by the device, equal to
    the device-resident call; with two devices in ONE process (skipped on a single-GPU box) the result is bytewise
    the same."""
    import torch
    from semibin_b200.neighbors import knn, kneighbors_graph
    N, k = 150000, 64
    X = synth.embeddings(N, seed=31)
    d0, i0 = knn(torch.from_numpy(X).cuda(), k)
    g, st = kneighbors_graph(X, k, mode='distance', p=2, n_blocks=3, return_stats=True)
    assert g.shape == (N, N) and g.data.dtype == np.float64 and g.indices.dtype == np.int64 and g.indptr.dtype == np.int64
    np.testing.assert_array_equal(g.indptr, np.arange(0, N * k + 1, k))
    np.testing.assert_array_equal(g.indices.reshape(N, k), i0.cpu().numpy())
    np.testing.assert_array_equal(g.data.reshape(N, k), d0.cpu().numpy().astype(np.float64))
    assert g.nnz == N * k and (g[5].toarray() != 0).sum() == k
    gc = kneighbors_graph(X, k, mode='connectivity')
    assert (gc.data == 1.0).all() and np.array_equal(gc.indices, g.indices)
    F = synth.kmer_features(30000, seed=31)                      # float64 input: distances are not rounded
    df, jf = knn(torch.from_numpy(F).cuda(), 20)
    gf = kneighbors_graph(F, 20)
    np.testing.assert_array_equal(gf.data.reshape(30000, 20), df.cpu().numpy())
    np.testing.assert_array_equal(gf.indices.reshape(30000, 20), jf.cpu().numpy())
    if torch.cuda.device_count() < 2:
        pytest.skip('single-GPU box: the two-device single-process path needs two GPUs')
    g2 = kneighbors_graph(X, k, devices=2)
    assert np.array_equal(g2.indices, g.indices) and np.array_equal(g2.data, g.data)
    from semibin_b200.cluster import build_edge_list
    emb, feat, depth, L = synth.short_read_case(60000, n_sample=3, seed=9)
    e1 = build_edge_list(emb, feat, depth, max_edges=100, max_node=0.9, is_combined=False, n_sample=3)
    e2 = build_edge_list(emb, feat, depth, max_edges=100, max_node=0.9, is_combined=False, n_sample=3, devices=2)
    for a, b in zip(e1, e2):
        assert a.tobytes() == b.tobytes()


def test_rerank_mirrors_rows_to_peer_gpu():
    """sbk_knn_peers: the rerank kernels store every finished row into the peer GPU's full result buffer (P2P stores)
    as well as into their own; two devices each computing one row block leave BOTH full buffers equal to the
    single-device result.  (Skipped on a single-GPU box; bench.py --gpus N uses the same entry point under torchrun.)"""
    import ctypes as C
    import torch
    if torch.cuda.device_count() < 2 or not torch.cuda.can_device_access_peer(0, 1):
        pytest.skip('needs two GPUs with peer access')
    from semibin_b200 import _lib
    from semibin_b200.neighbors import knn, _workspace
    from semibin_b200.dist import row_block
    lib = _lib.load()
    N, k = 120000, 50
    X = synth.embeddings(N, seed=41)
    d0, i0 = knn(torch.from_numpy(X).cuda(0), k)
    full_d = [torch.zeros((N, k), dtype=torch.float32, device=f'cuda:{g}') for g in range(2)]
    full_i = [torch.full((N, k), -1, dtype=torch.int32, device=f'cuda:{g}') for g in range(2)]
    full_d[1].copy_(full_d[0]); full_d[0].copy_(full_d[1])      # torch enables peer access on the first cross-device copy
    for g in range(2):
        with torch.cuda.device(g):
            Xg = torch.from_numpy(X).cuda(g)
            lo, hi = row_block(N, g, 2)
            ws = _workspace(lib.sbk_knn_workspace_bytes(_lib.SBK_F32, N, 100, hi - lo, k, 0), Xg.device)
            pd = (C.c_void_p * 1)(full_d[1 - g].data_ptr())
            pi = (C.c_void_p * 1)(full_i[1 - g].data_ptr())
            st = _lib.KnnStats()
            rc = lib.sbk_knn_peers(_lib.ptr(Xg), _lib.SBK_F32, N, 100, 100, lo, hi, k, _lib.ptr(full_d[g][lo:hi]),
                                   _lib.ptr(full_i[g][lo:hi]), 1, pd, pi, _lib.ptr(ws), ws.numel(), 0, C.byref(st),
                                   _lib.stream_ptr(Xg.device))
            _lib.check(rc)
            torch.cuda.synchronize(g)
    for g in range(2):
        assert torch.equal(full_d[g].cpu(), d0.cpu()) and torch.equal(full_i[g].cpu(), i0.cpu())
