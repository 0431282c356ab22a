"""CPU tier: libsbk.so loads without a GPU and exports every symbol that
include/sbk.h declares; argument validation works without touching a device."""
import ctypes as C
import os
import re

import pytest

from semibin_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, 'include', 'sbk.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(sbk_[a-z0-9_]+)\s*\(', txt)))


def test_header_symbols_exported():
    lib = _lib.load()
    syms = _declared_symbols()
    assert 'sbk_knn' in syms and 'sbk_edges_stage2' in syms and 'sbk_dbscan_labels' in syms
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in sbk.h but not exported'
    assert set(syms) == set(_lib.SIGNATURES), (set(syms) ^ set(_lib.SIGNATURES))


def test_version_and_error_slot():
    lib = _lib.load()
    assert lib.sbk_version() == 100
    assert isinstance(_lib.last_error(), str)


def test_knn_argument_validation_matches_sklearn_message():
    lib = _lib.load()
    # k >= n -> workspace query reports the sklearn error text (neighbors/_base.py:840-851)
    assert lib.sbk_knn_workspace_bytes(_lib.SBK_F32, 10, 100, 10, 10, 0) == 0
    assert 'Expected n_neighbors < n_samples_fit' in _lib.last_error()
    assert lib.sbk_knn_workspace_bytes(_lib.SBK_F32, 100000, 100, 100000, 200, 0) > 0
    assert lib.sbk_knn_workspace_bytes(_lib.SBK_F32, 100000, 100, 100000, 2000, 0) == 0
    assert lib.sbk_knn_workspace_bytes(7, 100000, 100, 100000, 20, 0) == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    lib = _lib.load()
    assert lib.sbk_device_count() == 0
    buf = (C.c_char * 1024)()
    rc = lib.sbk_knn(C.cast(buf, C.c_void_p), 0, 100, 16, 16, 0, 100, 5, C.cast(buf, C.c_void_p),
                     C.cast(buf, C.c_void_p), C.cast(buf, C.c_void_p), 1 << 30, 1, None, None)
    assert rc == -3 and 'no CUDA device' in _lib.last_error()
    with pytest.raises(_lib.SbkError):
        _lib.require_cuda()
    from semibin_b200.neighbors import kneighbors_graph
    import numpy as np
    with pytest.raises(Exception):
        kneighbors_graph(np.zeros((50, 20), np.float32), 5)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'semibin_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), fn
