"""ctypes binding of libsbk.so (C-ABI declared in include/sbk.h).

The product path has NO CPU fallback: a missing library or a missing CUDA
device raises.  torch is used only to own device memory and streams.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SBK_LIB_VARIANT=probe selects the profiling build (libsbk_probe.so); nothing else does
PROBES = os.environ.get('SBK_LIB_VARIANT', '') == 'probe'
LIB_PATH = os.path.join(HERE, 'libsbk_probe.so' if PROBES else 'libsbk.so')

SBK_F32, SBK_F64 = 0, 1
SBK_OK, SBK_ERR_INVALID, SBK_ERR_WORKSPACE, SBK_ERR_CUDA, SBK_ERR_INTERNAL = 0, -1, -2, -3, -4
KNN_AUTO, KNN_EXACT, KNN_TENSOR, KNN_TENSOR_ONESIDED, KNN_TENSOR_SYMMETRIC, KNN_TENSOR_SYMMETRIC_STARVED = 0, 1, 2, 3, 4, 5
ALGOS = {'auto': KNN_AUTO, 'exact': KNN_EXACT, 'tensor': KNN_TENSOR, 'tensor_onesided': KNN_TENSOR_ONESIDED,
         'tensor_symmetric': KNN_TENSOR_SYMMETRIC, 'tensor_symmetric_starved': KNN_TENSOR_SYMMETRIC_STARVED}
MAX_THRESHOLDS = 32


class KnnStats(C.Structure):
    _fields_ = [('algo_used', C.c_int32), ('n_chunks', C.c_int32),
                ('n_fallback_rows', C.c_int64), ('n_candidates', C.c_int64),
                ('sample_size', C.c_int32), ('cand_capacity', C.c_int32),
                ('order_stat', C.c_int32), ('kpad', C.c_int32),
                ('n_launches', C.c_int32), ('n_filter_launches', C.c_int32),
                ('ms_prepare', C.c_float), ('ms_sample', C.c_float), ('ms_filter', C.c_float),
                ('ms_rerank', C.c_float), ('ms_fallback', C.c_float), ('ms_exact_filter', C.c_float),
                ('n_reranked', C.c_int64), ('n_fail_overflow', C.c_int64), ('n_fail_few', C.c_int64),
                ('n_fail_survivors', C.c_int64), ('n_fail_certificate', C.c_int64),
                ('n_split_cols', C.c_int32), ('symmetric', C.c_int32), ('n_retry_rows', C.c_int64)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class SbkError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); every symbol include/sbk.h declares
SIGNATURES = {
    'sbk_version': (C.c_int, []),
    'sbk_source_hash': (C.c_char_p, []),
    'sbk_probes_enabled': (C.c_int, []),
    'sbk_last_error': (C.c_char_p, []),
    'sbk_device_count': (C.c_int, []),
    'sbk_knn_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_int]),
    'sbk_knn': (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(KnnStats), C.c_void_p]),
    'sbk_knn_peers': (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                C.c_void_p, C.c_size_t, C.c_int, C.POINTER(KnnStats), C.c_void_p]),
    'sbk_knn_sym_shard_xchg_bytes': (C.c_size_t, [C.c_int, C.c_int64, C.c_int32, C.c_int32, C.c_int32]),
    'sbk_knn_sym_shard': (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                    C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                    C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_size_t,
                                    C.POINTER(KnnStats), C.c_void_p]),
    'sbk_knn_stream': (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, C.c_int,
                                 C.POINTER(KnnStats), C.c_void_p, C.c_void_p]),
    'sbk_knn_graph_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_int]),
    'sbk_knn_graph': (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, C.c_int,
                                C.POINTER(KnnStats), C.c_void_p, C.c_void_p]),
    'sbk_edges_stage1': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                   C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'sbk_edges_stage2_workspace_bytes': (C.c_size_t, [C.c_int64, C.c_int32, C.c_int64, C.c_int32]),
    'sbk_edges_stage2': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_double,
                                   C.c_void_p, C.c_int64, C.c_int32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                   C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'sbk_radius_sweep': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
                                   C.POINTER(C.c_double), C.c_int32, C.c_void_p, C.c_double,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    'sbk_dbscan_labels_workspace_bytes': (C.c_size_t, [C.c_int64, C.c_int32]),
    'sbk_dbscan_labels': (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                    C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'sbk_integrate_workspace_bytes': (C.c_size_t, [C.c_int64, C.c_int32]),
    'sbk_integrate': (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                C.c_int64, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_size_t, C.c_void_p]),
    'sbk_kmeans_batch_workspace_bytes': (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32, C.c_int64]),
    'sbk_kmeans_batch': (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                                   C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_size_t, C.c_void_p]),
    'sbk_debug_shard_mail_cap': (None, [C.c_int64]),
    'sbk_debug_sym_shard_schedule': (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
    'sbk_debug_tc_workspace_bytes': (C.c_size_t, [C.c_int64, C.c_int32]),
    'sbk_debug_tc_scores': (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.c_int64, C.c_int64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_void_p]),
}


def load():
    """Load libsbk.so (building is done by __graft_entry__.build / _build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SbkError(f'{LIB_PATH} not found: build it with `python -m semibin_b200._build` '
                       '(there is no CPU fallback)')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    # a binary that was not built from the sources in this tree is refused (stale prebuilt .so)
    from . import _build
    want = _build.source_hash(PROBES)
    got = lib.sbk_source_hash().decode('ascii', 'replace')
    if got != want:
        raise SbkError(f'{LIB_PATH} was built from other sources (embedded hash {got}, tree {want}): rebuild with '
                       '`python -m semibin_b200._build`')
    if bool(lib.sbk_probes_enabled()) != PROBES:
        raise SbkError(f'{LIB_PATH}: probe build flag does not match SBK_LIB_VARIANT')
    _lib = lib
    return lib


def last_error():
    return load().sbk_last_error().decode('utf-8', 'replace')


def check(rc):
    if rc == 0:
        return
    msg = last_error()
    if rc == -1:
        raise ValueError(msg)
    raise SbkError(f'libsbk error {rc}: {msg}')


def require_cuda():
    import torch
    if not torch.cuda.is_available() or load().sbk_device_count() < 1:
        raise SbkError('semibin_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    """Current torch stream of `device` (default: the current device) as a cudaStream_t."""
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
