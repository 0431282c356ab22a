"""GPU tier: kNN graph parity through the C-ABI (ctypes -> libsbk.so) against
the oracle restatement, the committed golden vectors (live sklearn outputs)
and, when importable, live scikit-learn."""
import numpy as np
import pytest

from oracle import knn as ok, compare
from semibin_b200 import synth

pytestmark = pytest.mark.gpu


def _gpu_knn(X, k, algo='exact', **kw):
    import torch
    from semibin_b200.neighbors import knn
    Xd = torch.from_numpy(np.ascontiguousarray(X)).cuda()
    out = knn(Xd, k, algo=algo, **kw)
    d, i = out[0], out[1]
    res = (d.cpu().numpy().astype(np.float64), i.cpu().numpy().astype(np.int64))
    return res + (out[2],) if len(out) == 3 else res


@pytest.mark.parametrize('algo', ['exact', 'tensor'])
def test_bin_data_golden(golden, algo):
    g = golden('bin_data')
    for k in (20, 39):
        d, i = _gpu_knn(g['embedding'], k, algo)
        # rows 16/17 are duplicates: sklearn's expanded form leaves a BLAS residue (2.98e-8), the
        # direct-difference form gives exactly 0 (DESIGN.md "zero-distance tie class") -> atol
        compare.assert_knn_parity(g[f'emb_k{k}_d'], g[f'emb_k{k}_i'], d, i, rtol=1e-7, atol=1e-6)
        d, i = _gpu_knn(np.ascontiguousarray(g['values'][:, :136]), k, algo)
        # the fixture has one duplicated k-mer row: sklearn returns a BLAS residue
        # (or 0), the direct-difference form returns exactly 0 -> atol
        compare.assert_knn_parity(g[f'kmer_k{k}_d'], g[f'kmer_k{k}_i'], d, i, rtol=1e-9, atol=1e-7)
        d, i = _gpu_knn(g['long_emb_new'], 39, algo)
        compare.assert_knn_parity(g['long_d'], g['long_i'], d, i, rtol=1e-7, atol=1e-6)


@pytest.mark.parametrize('algo', ['exact', 'tensor'])
def test_synth_golden(golden, algo):
    g = golden('synth_knn')
    X = synth.embeddings(4096, seed=101)
    d, i = _gpu_knn(X, 50, algo)
    res = compare.assert_knn_parity(g['emb4096_k50_d'], g['emb4096_k50_i'], d, i, rtol=1e-6, max_tied_frac=0.001)
    assert (d == g['emb4096_k50_d']).mean() > 0.9999           # sklearn's float32 rounding recipe reproduced
    F = synth.kmer_features(4096, seed=101)
    d, i = _gpu_knn(F, 50, algo)
    compare.assert_knn_parity(g['kmer4096_k50_d'], g['kmer4096_k50_i'], d, i, rtol=1e-6, max_tied_frac=0.001)
    X = synth.embeddings(1500, seed=102)
    d, i = _gpu_knn(X, 200, algo)
    compare.assert_knn_parity(g['emb1500_k200_d'], g['emb1500_k200_i'], d, i, rtol=1e-6, max_tied_frac=0.001)
    X = synth.embeddings(300, seed=103)                        # k = N - 1
    d, i = _gpu_knn(X, 299, algo)
    compare.assert_knn_parity(g['emb300_k299_d'], g['emb300_k299_i'], d, i, rtol=1e-6)
    Xl, _ = synth.long_read_case(3000, n_sample=3, seed=104)   # D = 103 (odd width, unaligned rows)
    d, i = _gpu_knn(Xl, 64, algo)
    compare.assert_knn_parity(g['long3000_k64_d'], g['long3000_k64_i'], d, i, rtol=1e-6, max_tied_frac=0.001)


@pytest.mark.parametrize('algo', ['exact', 'tensor'])
def test_ties_and_duplicates(golden, algo):
    g = golden('ties')
    for name, kw, dt in [('grid', dict(n=700, n_dup=0), np.float32), ('dup5', dict(n=700, n_dup=5), np.float32),
                         ('dup40', dict(n=700, n_dup=40), np.float32), ('grid64', dict(n=700, n_dup=5), np.float64)]:
        X = synth.tie_grid(dtype=dt, **kw)
        d, i = _gpu_knn(X, 20, algo)
        res = compare.assert_knn_parity(g[name + '_d'], g[name + '_i'], d, i, rtol=0)
        assert res['max_rel'] == 0.0                           # identical distance multiset
        # our own tie rule is deterministic: lower index wins, (distance, index) order
        rd, ri = ok.knn_restated(X, 20)
        np.testing.assert_array_equal(i, ri)
        np.testing.assert_array_equal(d, rd)


@pytest.mark.parametrize('algo', ['exact', 'tensor'])
def test_row_block_and_oracle(algo):
    X = synth.embeddings(6000, seed=7)
    d, i = _gpu_knn(X, 31, algo, q_begin=1234, q_end=3001)
    rd, ri = ok.knn_restated(X, 31, q_begin=1234, q_end=3001)
    res = compare.assert_knn_parity(rd, ri, d, i, rtol=1e-6, max_tied_frac=0.001)
    assert res['n_rows'] == 3001 - 1234


def test_tensor_path_20k_vs_exact():
    """N=20000, k=200: tensor-core filter + certificate == FP64 exact path, row by row."""
    X = synth.embeddings(20000, seed=20241)
    d1, i1 = _gpu_knn(X, 200, 'exact')
    d2, i2, st = _gpu_knn(X, 200, 'tensor', return_stats=True)
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_array_equal(d1, d2)
    assert st['algo_used'] == 2
    assert st['n_fallback_rows'] < 0.05 * 20000, st
    F = synth.kmer_features(20000, seed=20241)
    d1, i1 = _gpu_knn(F, 200, 'exact')
    d2, i2, st = _gpu_knn(F, 200, 'tensor', return_stats=True)
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_array_equal(d1, d2)


@pytest.mark.parametrize('n,k', [(20000, 200), (33000, 64), (50177, 31)])
def test_symmetric_filter_equals_onesided_and_exact(n, k):
    """The symmetric self-join filter (every tile product tested for both of its row sets; the default for builds of
    all rows from ~400k rows on, forced here) must give the one-sided filter's and the FP64 exact path's result, bit
    for bit.  Odd tile counts here (79, 129, 197), an even one below."""
    X = synth.embeddings(n, seed=20260 + n)
    d0, i0 = _gpu_knn(X, k, 'exact')
    d1, i1, st1 = _gpu_knn(X, k, 'tensor_symmetric', return_stats=True)
    d2, i2, st2 = _gpu_knn(X, k, 'tensor_onesided', return_stats=True)
    assert st1['algo_used'] == 2 and st1['symmetric'] == 1, st1
    assert st2['algo_used'] == 2 and st2['symmetric'] == 0, st2
    np.testing.assert_array_equal(i0, i1)
    np.testing.assert_array_equal(d0, d1)
    np.testing.assert_array_equal(i0, i2)
    np.testing.assert_array_equal(d0, d2)
    assert st1['n_fallback_rows'] < 0.05 * n, st1


def test_symmetric_filter_even_tile_count_and_features():
    """Even tile count (the opposite tile is multiplied by both of its pairs and row-tested only) on the float64
    k-mer features (split precision columns), against the exact path."""
    n = 256 * 80
    F = synth.kmer_features(n, seed=20262)
    d0, i0 = _gpu_knn(F, 50, 'exact')
    d1, i1, st = _gpu_knn(F, 50, 'tensor_symmetric', return_stats=True)
    assert st['symmetric'] == 1, st
    np.testing.assert_array_equal(i0, i1)
    np.testing.assert_array_equal(d0, d1)
    X = synth.tie_grid(n, d=24, seed=5, n_dup=300)
    d0, i0 = _gpu_knn(X, 20, 'exact')
    d1, i1, st = _gpu_knn(X, 20, 'tensor_symmetric', return_stats=True)
    assert st['symmetric'] == 1, st
    np.testing.assert_array_equal(i0, i1)
    np.testing.assert_array_equal(d0, d1)


def test_live_sklearn_sampled_rows_200k():
    """Scale check: random query rows of a 200k set against live sklearn."""
    sk = pytest.importorskip('sklearn.neighbors')
    N, k = 200000, 200
    X = synth.embeddings(N, seed=20241)
    rows = np.sort(np.random.default_rng(5).choice(N, 512, replace=False))
    d, i, st = _gpu_knn(X, k, 'auto', return_stats=True)
    assert st['algo_used'] == 2
    nn = sk.NearestNeighbors(n_neighbors=k + 1, algorithm='brute').fit(X)
    rd, ri = nn.kneighbors(X[rows])
    # drop self by index as kneighbors(X=None) does (neighbors/_base.py:943-957)
    for t, r in enumerate(rows):
        m = ri[t] != r
        if m.all():
            m[0] = False
        compare.assert_knn_parity(rd[t][m][None], ri[t][m][None], d[r][None], i[r][None], rtol=1e-6)
    assert st['n_fallback_rows'] < 0.01 * N, st
    # size-independent properties at full size
    assert (np.diff(d, axis=1) >= 0).all()                      # rows sorted
    assert (i != np.arange(N)[:, None]).all()                   # self excluded
    assert i.min() >= 0 and i.max() < N


def test_symmetric_pool_exhaustion_falls_back_to_onesided():
    """A symmetric build whose record pool runs dry (test hook: 48 chunks) has lost group records; the call notices
    (flag checked after the filter) and redoes the build one-sided -- device-resident and through the host-buffer
    block-row schedule, where finished blocks had already been written out."""
    import torch
    from semibin_b200.neighbors import kneighbors_graph
    n, k = 30000, 40
    X = synth.embeddings(n, seed=20290)
    d0, i0 = _gpu_knn(X, k, 'exact')
    d1, i1, st = _gpu_knn(X, k, 'tensor_symmetric_starved', return_stats=True)
    assert st['symmetric'] == 0 and st['algo_used'] == 2, st          # the one-sided rebuild reports itself
    np.testing.assert_array_equal(i0, i1)
    np.testing.assert_array_equal(d0, d1)
    g, stg = kneighbors_graph(X, k, algo='tensor_symmetric_starved', return_stats=True)
    assert stg['symmetric'] == 0, stg
    np.testing.assert_array_equal(g.indices.reshape(n, k), i0)
    np.testing.assert_array_equal(g.data.reshape(n, k), d0)


def test_default_symmetric_build_450k_vs_live_sklearn():
    """A build of all rows above ~400k rows takes the symmetric filter by default: sampled rows against live sklearn
    (float32 embedding, k = 100), plus the full-size properties.  (The float64 feature graph at that size:
    tests/test_gpu_scale.py.)"""
    sk = pytest.importorskip('sklearn.neighbors')
    N, k = 450000, 100
    X = synth.embeddings(N, seed=20270)
    rows = np.sort(np.random.default_rng(7).choice(N, 384, replace=False))
    d, i, st = _gpu_knn(X, k, 'auto', return_stats=True)
    assert st['algo_used'] == 2 and st['symmetric'] == 1, st
    nn = sk.NearestNeighbors(n_neighbors=k + 1, algorithm='brute').fit(X)
    rd, ri = nn.kneighbors(X[rows])
    for t, r in enumerate(rows):
        m = ri[t] != r
        if m.all():
            m[0] = False
        compare.assert_knn_parity(rd[t][m][None], ri[t][m][None], d[r][None], i[r][None], rtol=1e-6)
    assert st['n_fallback_rows'] < 0.01 * N, st
    assert (np.diff(d, axis=1) >= 0).all()
    assert (i != np.arange(N)[:, None]).all()
    assert i.min() >= 0 and i.max() < N


def test_errors():
    import torch
    from semibin_b200.neighbors import knn, kneighbors_graph
    X = torch.zeros((10, 20), device='cuda')
    with pytest.raises(ValueError, match='Expected n_neighbors < n_samples_fit'):
        knn(X, 10)
    with pytest.raises(ValueError):
        kneighbors_graph(np.zeros((10, 20), np.float32), 3, p=1)


def test_kneighbors_graph_csr_layout(golden):
    g = golden('bin_data')
    from semibin_b200.neighbors import kneighbors_graph
    G = kneighbors_graph(g['embedding'], n_neighbors=20, mode='distance', p=2, n_jobs=4)
    assert G.shape == (40, 40) and G.format == 'csr'
    assert G.data.dtype == np.float64 and G.indices.dtype == np.int64 and G.indptr.dtype == np.int64
    np.testing.assert_array_equal(G.indptr, np.arange(0, 801, 20))
    compare.assert_knn_parity(g['emb_k20_d'], g['emb_k20_i'], G.data.reshape(40, 20), G.indices.reshape(40, 20), rtol=1e-7, atol=1e-6)


def test_knn_host_blocks_equal_device_call():
    """Host-buffer entry point (row blocks, download overlapped with the next block) == one device call."""
    import torch
    from semibin_b200.neighbors import knn, knn_host
    X = synth.embeddings(150000, seed=77)
    Xd = torch.from_numpy(X).cuda()
    d0, i0 = knn(Xd, 50, q_begin=1000, q_end=141000)
    hd, hi = knn_host(torch.from_numpy(X).pin_memory(), 50, q_begin=1000, q_end=141000, n_blocks=3, min_block_rows=1)
    assert torch.equal(hd, d0.cpu()) and torch.equal(hi, i0.cpu())


def test_clustered_row_order_320k():
    """Rows sorted by genome (neighbours adjacent in index, as contigs of one assembly are): the reference side is
    staged in a pseudo-random order, so the threshold sample and the between-launch refinement stay unbiased.
    Sampled rows against live sklearn; almost no row may need the exact fallback."""
    sk = pytest.importorskip('sklearn.neighbors')
    N, k = 320000, 100
    X, genome = synth.embeddings(N, seed=9, return_genome=True)
    X = np.ascontiguousarray(X[np.argsort(genome, kind='stable')])
    d, i, st = _gpu_knn(X, k, 'auto', return_stats=True)
    assert st['algo_used'] == 2 and st['n_filter_launches'] >= 2
    assert st['n_fallback_rows'] < 0.002 * N, st
    rows = np.sort(np.random.default_rng(6).choice(N, 256, replace=False))
    nn = sk.NearestNeighbors(n_neighbors=k + 1, algorithm='brute').fit(X)
    rd, ri = nn.kneighbors(X[rows])
    for t, r in enumerate(rows):
        m = ri[t] != r
        if m.all():
            m[0] = False
        compare.assert_knn_parity(rd[t][m][None], ri[t][m][None], d[r][None], i[r][None], rtol=1e-6)
    assert (np.diff(d, axis=1) >= 0).all() and (i != np.arange(N)[:, None]).all()


def test_stream_with_fallback_rows():
    """Rows the tensor path cannot certify (hundreds of exact duplicates: more than a candidate segment holds of one
    value, and the 'self not among the k+1 nearest' corner of neighbors/_base.py:946-950) are recomputed by the
    exact filter at the END of the call; the host-buffer path must carry the recomputed rows, not the chunk
    downloads made before."""
    import torch
    from semibin_b200.neighbors import knn, knn_host
    X = synth.embeddings(40000, seed=123)
    X[100:1100] = X[7]                                   # 1001 identical rows, k = 50
    Xd = torch.from_numpy(X).cuda()
    d0, i0 = knn(Xd, 50, algo='exact')
    d1, i1, st = knn(Xd, 50, algo='tensor', return_stats=True)
    assert torch.equal(d0, d1) and torch.equal(i0, i1)
    hd, hi, sth = knn_host(torch.from_numpy(X).pin_memory(), 50, algo='tensor', n_blocks=4, min_block_rows=1,
                           return_stats=True)
    assert sth['n_chunks'] >= 2
    assert torch.equal(hd, d0.cpu()) and torch.equal(hi, i0.cpu())
    print('fallback rows', st['n_fallback_rows'], sth['n_fallback_rows'])
