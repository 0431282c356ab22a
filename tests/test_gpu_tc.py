"""GPU tier: the tcgen05 tile product and its per-pair error model (the premise of
the exactness certificate): the accumulator is a LOWER bound of the exact score
and L + slack an UPPER bound, for every pair."""
import numpy as np
import pytest

from semibin_b200 import synth, _lib

pytestmark = pytest.mark.gpu


def _scores(X, n_q):
    import torch
    lib = _lib.load()
    N, D = X.shape
    Xd = torch.from_numpy(np.ascontiguousarray(X)).cuda()
    dtype = _lib.SBK_F32 if X.dtype == np.float32 else _lib.SBK_F64
    npad = (N + 255) // 256 * 256
    out = torch.zeros((n_q, npad), dtype=torch.float32, device='cuda')
    norms4 = torch.zeros((N, 4), dtype=torch.float32, device='cuda')
    nrm = torch.zeros((N + 1,), dtype=torch.float64, device='cuda')
    info = torch.zeros((2,), dtype=torch.int32, device='cuda')
    wsb = lib.sbk_debug_tc_workspace_bytes(N, D)
    ws = torch.empty((wsb,), dtype=torch.uint8, device='cuda')
    rc = lib.sbk_debug_tc_scores(_lib.ptr(Xd), dtype, N, D, Xd.stride(0), n_q, _lib.ptr(out), _lib.ptr(norms4),
                                 _lib.ptr(nrm), _lib.ptr(info), _lib.ptr(ws), wsb, _lib.stream_ptr())
    _lib.check(rc)
    torch.cuda.synchronize()
    return out.cpu().numpy()[:, :N], norms4.cpu().numpy().astype(np.float64), nrm.cpu().numpy(), info.cpu().numpy()


def _upper_slack(ni, nj, c_acc):
    """tc_pair_upper_slack of semibin_b200/csrc/knn.cuh."""
    G = 2 * ni[:, None, 1] * nj[None, :, 0] + ni[:, None, 2] * (2 * nj[None, :, 1] + 2 * c_acc * nj[None, :, 2])
    J = c_acc * nj[None, :, 0] ** 2 + nj[None, :, 3]
    return 2.02 * G + 2.02 * J + ni[:, None, 3] + 2.4e-7 * (nj[None, :, 0] + nj[None, :, 1] + ni[:, None, 2] + ni[:, None, 1]) + 4e-7


def _case(name):
    rng = np.random.default_rng(17)
    if name == 'emb':
        return synth.embeddings(3000, seed=31)
    if name == 'kmer':
        return synth.kmer_features(3000, seed=31)
    if name == 'long':
        return synth.long_read_case(3000, n_sample=3, seed=31)[0]
    if name == 'two_far_clusters':            # large norms after centring, tiny intra-cluster distances
        c = rng.normal(0, 1, (2, 64)).astype(np.float32) * 5
        x = c[rng.integers(0, 2, 2500)] + rng.normal(0, 1e-3, (2500, 64)).astype(np.float32)
        return np.ascontiguousarray(x.astype(np.float32))
    if name == 'dominant_column':             # one column 300x the others + outlier rows
        x = rng.normal(0, 0.05, (2500, 40))
        x[:, 7] = rng.normal(0, 15, 2500)
        x[:5] *= 20
        return np.ascontiguousarray(x.astype(np.float64))
    if name == 'duplicates_and_zero':
        x = rng.normal(0, 1, (2000, 32)).astype(np.float32)
        x[100:400] = x[0]
        x[1000:] = 0
        return x
    raise KeyError(name)


@pytest.mark.parametrize('case', ['emb', 'kmer', 'long', 'two_far_clusters', 'dominant_column', 'duplicates_and_zero'])
def test_lower_and_upper_bounds_hold_for_every_pair(case):
    X = _case(case)
    n_q = 640
    L, n4, nrm, info = _scores(X, n_q)
    N = X.shape[0]
    S = nrm[-1]
    kpad, n_split = int(info[0]), int(info[1])
    c_acc = (kpad // 16) * 2.0 ** -20
    Xd = X.astype(np.float64)
    d2 = np.empty((n_q, N))
    for s in range(0, n_q, 64):
        diff = Xd[s:s + 64, None, :] - Xd[None, :, :]
        d2[s:s + 64] = np.einsum('ijk,ijk->ij', diff, diff)
    e = S * S * d2 - nrm[:n_q, None]                      # exact scores
    Ld = L.astype(np.float64)
    sigma = n4[:n_q, 3][:, None]
    # rigour: lower bound (up to the documented sigma_i) and upper bound, every pair
    tol = 1e-9 * (np.abs(e) + nrm[:n_q, None] + 1)
    viol_lo = (Ld - (e + sigma) > tol)
    assert not viol_lo.any(), (case, int(viol_lo.sum()), float((Ld - e - sigma).max()))
    U = Ld + _upper_slack(n4[:n_q], n4, c_acc)
    viol_hi = (e - U > tol)
    assert not viol_hi.any(), (case, int(viol_hi.sum()), float((e - U).max()))
    # tightness report: gap vs the spread of the 50 nearest exact scores
    es = np.sort(e, axis=1)
    spread = np.median(es[:, min(50, N - 1)] - es[:, 1])
    near = e <= es[:, [min(50, N - 1)]]
    gap = np.median((e - Ld)[near])
    used = np.max((e + sigma - Ld) / np.maximum(_upper_slack(n4[:n_q], n4, c_acc) / 2.02, 1e-300))
    print(f'{case}: S={S} kpad={kpad} split={n_split} median gap(e-L) near={gap:.3e} spread(50nn)={spread:.3e} '
          f'max (e-L)/bound={used:.3f}')
    if case in ('long', 'dominant_column'):
        assert n_split >= 1
    if case == 'emb':
        assert n_split == 0 and kpad == 112
