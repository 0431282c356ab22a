"""GPU tier: short-read edge list (hand-off #1) parity."""
import numpy as np
import pytest

from oracle import knn as ok, edges as oe, compare
from semibin_b200 import synth

pytestmark = pytest.mark.gpu


def test_bin_data_handoff(golden):
    """The reference's own fixture through the fused path == the captured
    (edges, edge_weights) of the unmodified run_embed_infomap."""
    from semibin_b200.cluster import build_edge_list
    g = golden('bin_data')
    vals = g['values']
    depth = vals[:, 136:].astype(np.float32)
    for k in (20, 39):
        X, Y, V = build_edge_list(g['embedding'], np.ascontiguousarray(vals[:, :136]), depth, max_edges=k,
                                  max_node=1, is_combined=bool(g['is_combined']), n_sample=int(g['n_sample']))
        ref = (g[f'edges_k{k}_X'], g[f'edges_k{k}_Y'], g[f'edges_k{k}_V'])
        assert X.dtype == np.int32 and Y.dtype == np.int32 and V.dtype == np.float64
        S = int(g['n_sample'])
        res = compare.assert_edges_parity(ref, (X, Y, V), bound_fn=None if bool(g['is_combined']) else
                                          (lambda x, y: oe.kl_weight_bound(depth, S, x, y)))
        assert res['n_only_ref'] <= 1 and res['n_only_got'] <= 1, res   # duplicated k-mer row (d == 0 tie class)


@pytest.mark.parametrize('S,comb', [(1, False), (3, False), (10, False), (10, True)])
def test_synth_handoff(golden, S, comb):
    from semibin_b200.cluster import build_edge_list
    g = golden('synth_edges')
    emb, feat, depth, L = synth.short_read_case(3000, n_sample=S, seed=200 + S)
    f = feat
    if comb:
        f = np.concatenate([feat, depth[:, 0::2].astype(np.float64) / 100.0], axis=1)
        n = f.shape[1] - 136
        if n >= 20 or (n >= 5 and np.mean(np.sum(f[:, 136:], axis=1)) > 2):
            f = f / np.sum(f, axis=0)
            f = f / np.abs(f).sum(axis=1, keepdims=True)
    for max_node in (1, 0.6):
        X, Y, V = build_edge_list(emb, f, depth, max_edges=50, max_node=max_node, is_combined=comb, n_sample=S)
        tag = f'S{S}_{"comb" if comb else "kl"}_mn{max_node}'
        ref = (g[tag + '_X'], g[tag + '_Y'], g[tag + '_V'])
        # combined mode has no evaluator-dependent operation: the weights must be bit-equal
        res = compare.assert_edges_parity(ref, (X, Y, V), rtol=0.0 if comb else 1e-5,
                                          bound_fn=None if comb else (lambda x, y: oe.kl_weight_bound(depth, S, x, y)))
        assert res['n_only_ref'] + res['n_only_got'] <= 2, res


def test_edges_20k_vs_oracle():
    """N=20000, k=200, S=3: kernels vs the sparse restatement fed with the GPU's own graphs."""
    import torch
    from semibin_b200.cluster import build_edge_list
    emb, feat, depth, L = synth.short_read_case(20000, n_sample=3, seed=77)
    (X, Y, V), gr = build_edge_list(emb, feat, depth, max_edges=200, max_node=0.9, is_combined=False, n_sample=3,
                                    return_graphs=True)
    ad, ai = [t.cpu().numpy() for t in gr['emb']]
    bd, bi = [t.cpu().numpy() for t in gr['kmer']]
    RX, RY, RV, thr = oe.edge_list_restated(ad.astype(np.float64), ai.astype(np.int64), bd, bi.astype(np.int64), depth,
                                            max_node=0.9, is_combined=False, n_sample=3)
    assert thr == gr['threshold']
    res = compare.assert_edges_parity((RX, RY, RV), (X, Y, V), bound_fn=lambda x, y: oe.kl_weight_bound(depth, 3, x, y))
    assert res['n_common'] > 10000
    # the restatement evaluated with the correctly rounded float32 log (what the kernel computes) is bit-equal
    QX, QY, QV, _ = oe.edge_list_restated(ad.astype(np.float64), ai.astype(np.int64), bd, bi.astype(np.int64), depth,
                                          max_node=0.9, is_combined=False, n_sample=3, log_mode='rounded')
    np.testing.assert_array_equal(QX, X); np.testing.assert_array_equal(QY, Y); np.testing.assert_array_equal(QV, V)
    # properties
    assert (Y > X).all()
    key = X.astype(np.int64) << 32 | Y
    assert (np.diff(key) > 0).all()
    assert (V > 1e-6).all() and (V <= 1).all()


def test_kl_term_matches_reference_golden(golden):
    """cal_kl known answers (float32) through the stage-2 kernel: a complete graph on 128 nodes."""
    import torch
    from semibin_b200.cluster import edges_stage2
    g = golden('cal_kl')
    for t in (0, 1):
        m, v, kl = g[f'm{t}'], g[f'v{t}'], g[f'kl{t}']
        n = 128
        idx = np.tile(np.arange(n, dtype=np.int32), (n, 1))
        keep = np.full((n, n // 32), -1, np.int32)            # every entry survives the intersection
        dist0 = np.zeros((n, n), np.float32)                  # w = 1 - min(0, 1) = 1: the edge weight IS the KL factor
        depth = np.stack([m, v], axis=1).astype(np.float32)
        X, Y, V = edges_stage2(torch.from_numpy(keep).cuda(), torch.from_numpy(dist0).cuda(), torch.from_numpy(idx).cuda(),
                               0, 0.0, torch.from_numpy(depth).cuda(), 1, False)
        X, Y, V = X.cpu().numpy(), Y.cpu().numpy(), V.cpu().numpy()
        fac = (np.float32(1) - kl).astype(np.float32)         # acc = -kl; += 1; /= 1
        exp = fac[X, Y].astype(np.float64)
        keep = ~(fac.astype(np.float64) <= 1e-6)
        iu = np.argwhere(np.triu(keep, 1))
        assert len(X) == len(iu)
        # numpy's SIMD float32 log is not correctly rounded (ours is): differences stay inside the derived bound
        bnd = oe.kl_weight_bound(depth, 1, X, Y)
        assert (np.abs(V - exp) <= 1e-5 * np.abs(exp) + bnd).all()
        assert (np.abs(V - exp) > 1e-7).mean() < 0.2
