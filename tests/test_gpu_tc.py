"""GPU tier: the tcgen05 tile product and its error bound (the premise of the
exactness certificate)."""
import ctypes as C

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
    E = torch.zeros((n_q,), dtype=torch.float64, device='cuda')
    nrm = torch.zeros((n_q + 1,), dtype=torch.float64, device='cuda')
    wsb = lib.sbk_debug_tc_workspace_bytes(N, D)
    ws = torch.empty((wsb,), dtype=torch.uint8, device='cuda')
    rc = lib.sbk_debug_tc_scores(_lib.ptr(Xd), dtype, N, D, Xd.stride(0), n_q, _lib.ptr(out), _lib.ptr(E),
                                 _lib.ptr(nrm), _lib.ptr(ws), wsb, _lib.stream_ptr())
    _lib.check(rc)
    torch.cuda.synchronize()
    return out.cpu().numpy()[:, :N], E.cpu().numpy(), nrm.cpu().numpy()


@pytest.mark.parametrize('case', ['emb', 'kmer', 'long'])
def test_tile_scores_within_bound(case):
    if case == 'emb':
        X = synth.embeddings(3000, seed=31)
    elif case == 'kmer':
        X = synth.kmer_features(3000, seed=31)
    else:
        X, _ = synth.long_read_case(3000, n_sample=3, seed=31)
    n_q = 700
    s, E, nrm = _scores(X, n_q)
    S = nrm[-1]
    Xd = X.astype(np.float64)
    d2 = ((Xd[:n_q, None, :] - Xd[None, :, :]) ** 2).sum(-1) if X.shape[0] * n_q * X.shape[1] < 4e8 else None
    if d2 is None:
        nn = (Xd * Xd).sum(1)
        d2 = nn[:n_q, None] - 2 * Xd[:n_q] @ Xd.T + nn[None, :]
    exact = S * S * d2 - nrm[:n_q, None]
    err = np.abs(s.astype(np.float64) - exact)
    ratio = (err / E[:, None]).max()
    print(case, 'S', S, 'max err', err.max(), 'E range', E.min(), E.max(), 'max err/E', ratio,
          'typical score spread', np.median(np.sort(exact, axis=1)[:, 200] - np.sort(exact, axis=1)[:, 1]))
    assert ratio < 1.0, ratio          # the bound holds
    assert ratio > 1e-4                # and is not absurdly loose
