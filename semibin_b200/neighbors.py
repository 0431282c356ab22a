"""kNN graph drop-in for sklearn.neighbors.kneighbors_graph on the path
SemiBin/cluster.py:94-105 and SemiBin/long_read_cluster.py:108-113.

`kneighbors_graph(X, n_neighbors, mode='distance', p=2, n_jobs=None)` returns
the same object the reference gets from scikit-learn: a scipy csr_matrix with
exactly k entries per row ordered by ascending distance, float64 data (values
float32-rounded for float32 input), int64 indices (sklearn/neighbors/_base.py:
1031-1041).  `knn()` is the array-level call (device tensors in / out) used by
the fused entry points and by the row-block sharding in dist.py.
"""
import ctypes as C
import numpy as np

from . import _lib


def _as_device_matrix(X):
    """Accept numpy (host) or torch (host/device) 2-D float32/float64; return a
    contiguous CUDA tensor.  Host inputs go through pinned memory."""
    import torch
    if isinstance(X, np.ndarray):
        if X.dtype not in (np.float32, np.float64):
            X = X.astype(np.float64)
        X = np.ascontiguousarray(X)
        t = torch.from_numpy(X)
        return t.pin_memory().cuda(non_blocking=True) if t.numel() else t.cuda()
    if not isinstance(X, torch.Tensor):
        raise TypeError('X must be a numpy array or a torch tensor')
    if X.dtype not in (torch.float32, torch.float64):
        X = X.double()
    return X.contiguous().cuda()


_ws_cache = {}


def _workspace(nbytes, device):
    """Reusable workspace tensor (grown on demand) so repeated calls do not
    re-allocate; owned by this module, passed to the C-ABI as a plain buffer."""
    import torch
    t = _ws_cache.get(device)
    if t is None or t.numel() < nbytes:
        _ws_cache.pop(device, None)
        t = None
        torch.cuda.empty_cache()
        t = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _ws_cache[device] = t
    return t


def release_workspace():
    _ws_cache.clear()


def knn(X, n_neighbors, *, q_begin=0, q_end=None, algo='auto', return_stats=False, out=None):
    """Exact Euclidean kNN of rows [q_begin, q_end) of X among all rows of X
    (self excluded by index).  X: CUDA tensor [N, D] float32/float64.
    Returns (dist, idx): dist float32 (float64 for float64 input) [n, k],
    idx int32 [n, k], both CUDA tensors, rows ordered by (distance, index).
    `out=(dist, idx)` reuses caller-owned CUDA tensors of those shapes/dtypes (no allocation in the call)."""
    import torch
    _lib.require_cuda()
    lib = _lib.load()
    if X.dim() != 2:
        raise ValueError('X must be 2-D')
    N, D = X.shape
    if q_end is None:
        q_end = N
    k = int(n_neighbors)
    if not (0 < k < N):
        # same condition/message family as sklearn/neighbors/_base.py:840-851
        raise ValueError(f'Expected n_neighbors < n_samples_fit, but n_neighbors = {k}, n_samples_fit = {N}, '
                         f'n_samples = {q_end - q_begin}')
    dtype = _lib.SBK_F32 if X.dtype == torch.float32 else _lib.SBK_F64
    n_q = q_end - q_begin
    a = _lib.ALGOS[algo] if isinstance(algo, str) else int(algo)
    ws_bytes = lib.sbk_knn_workspace_bytes(dtype, N, D, n_q, k, a)
    if ws_bytes == 0:
        raise ValueError(_lib.last_error())
    ws = _workspace(ws_bytes, X.device)
    if out is not None:
        out_d, out_i = out
        if (tuple(out_d.shape) != (n_q, k) or tuple(out_i.shape) != (n_q, k) or out_d.dtype != X.dtype
                or out_i.dtype != torch.int32 or not out_d.is_contiguous() or not out_i.is_contiguous()
                or out_d.device != X.device or out_i.device != X.device):
            raise ValueError('out must be contiguous (dist, idx) CUDA tensors of shape [n, k], dtypes (X.dtype, int32)')
    else:
        out_d = torch.empty((n_q, k), dtype=X.dtype, device=X.device)
        out_i = torch.empty((n_q, k), dtype=torch.int32, device=X.device)
    stats = _lib.KnnStats()
    rc = lib.sbk_knn(_lib.ptr(X), dtype, N, D, X.stride(0), q_begin, q_end, k,
                     _lib.ptr(out_d), _lib.ptr(out_i), _lib.ptr(ws), ws.numel(), a,
                     C.byref(stats), _lib.stream_ptr())
    _lib.check(rc)
    if return_stats:
        return out_d, out_i, stats.as_dict()
    return out_d, out_i


def plan_chunk_rows(n_q, n_blocks, wave_rows, min_block_rows=None):
    """Rows per chunk of the streamed host path.  Chunks are whole WAVES of the filter grid (one CTA pair per 256
    query rows, one pair per two SMs -> `wave_rows` rows per wave): a chunk that ends mid-wave idles most SMs for
    the tail of each of its filter launches, whole-wave chunks cost no extra waves at all.  The last chunk is the
    small remainder, so the download that nothing can hide is short.  Returns n_q when the rows are not split."""
    if min_block_rows is None:
        min_block_rows = 2 * wave_rows
    n_blocks = max(1, min(int(n_blocks), n_q // max(1, int(min_block_rows))))
    if n_blocks <= 1:
        return n_q
    waves = n_q / wave_rows
    return max(1, int(np.ceil(waves / (n_blocks - 0.7)))) * wave_rows


def knn_host(X_host, n_neighbors, *, q_begin=0, q_end=None, out_dist=None, out_idx=None, n_blocks=4, algo='auto',
             group=None, shard_upload=True, min_block_rows=None, return_stats=False):
    """Host-buffer form of knn(): X_host is a (preferably pinned) host tensor / array, the result lands in
    host tensors (pinned ones are allocated when none are passed).  One library call (`sbk_knn_stream`): the
    query rows are processed in `n_blocks` chunks and the device->host copy of a finished chunk runs on a
    second stream while the next chunk is computed, so only the upload and the last chunk's download are
    exposed; the reference operands are staged once.
    Under torch.distributed (one process per GPU, every rank holding the same host matrix) each rank uploads
    only its own row block of X and the blocks are all-gathered over NVLink (`shard_upload`), instead of every
    rank pulling the whole matrix through the host's PCIe links.
    Returns (dist, idx) host tensors [n, k]."""
    import torch
    _lib.require_cuda()
    lib = _lib.load()
    if isinstance(X_host, np.ndarray):
        X_host = torch.from_numpy(np.ascontiguousarray(X_host))
    if X_host.dtype not in (torch.float32, torch.float64):
        X_host = X_host.float()
    X_host = X_host.contiguous()
    N, D = X_host.shape
    if q_end is None:
        q_end = N
    n_q = q_end - q_begin
    k = int(n_neighbors)
    if k < 1 or k >= N:
        raise ValueError(f'Expected n_neighbors < n_samples_fit, but n_neighbors = {k}, n_samples_fit = {N}')
    if out_dist is None:
        out_dist = torch.empty((n_q, k), dtype=X_host.dtype).pin_memory()
    if out_idx is None:
        out_idx = torch.empty((n_q, k), dtype=torch.int32).pin_memory()
    if (tuple(out_dist.shape) != (n_q, k) or tuple(out_idx.shape) != (n_q, k) or out_dist.dtype != X_host.dtype
            or out_idx.dtype != torch.int32 or not out_dist.is_contiguous() or not out_idx.is_contiguous()):
        raise ValueError('out_dist / out_idx must be contiguous host tensors [n, k] of dtypes (X.dtype, int32)')
    from . import dist as sdist
    rank, world = sdist.world_info(group)
    if world > 1 and shard_upload:
        lo, hi = sdist.row_block(N, rank, world)
        Xd = sdist.all_gather_rows(X_host[lo:hi].cuda(non_blocking=True), N, group)
    else:
        Xd = X_host.cuda(non_blocking=True)
    copy = _copy_stream(Xd.device)
    wave_rows = (torch.cuda.get_device_properties(Xd.device).multi_processor_count // 2) * 256
    chunk = plan_chunk_rows(n_q, n_blocks, wave_rows, min_block_rows)
    dtype = _lib.SBK_F32 if Xd.dtype == torch.float32 else _lib.SBK_F64
    a = _lib.ALGOS[algo] if isinstance(algo, str) else int(algo)
    ws_bytes = lib.sbk_knn_workspace_bytes(dtype, N, D, n_q, k, a)
    if ws_bytes == 0:
        raise ValueError(_lib.last_error())
    ws = _workspace(ws_bytes, Xd.device)
    dev_d = torch.empty((n_q, k), dtype=Xd.dtype, device=Xd.device)
    dev_i = torch.empty((n_q, k), dtype=torch.int32, device=Xd.device)
    stats = _lib.KnnStats()
    rc = lib.sbk_knn_stream(_lib.ptr(Xd), dtype, N, D, Xd.stride(0), q_begin, q_end, k, _lib.ptr(dev_d), _lib.ptr(dev_i),
                            out_dist.data_ptr(), out_idx.data_ptr(), chunk, _lib.ptr(ws), ws.numel(), a,
                            C.byref(stats), _lib.stream_ptr(), copy.cuda_stream)
    _lib.check(rc)
    return (out_dist, out_idx, stats.as_dict()) if return_stats else (out_dist, out_idx)


_copy_streams = {}


def _copy_stream(device):
    import torch
    s = _copy_streams.get(device)
    if s is None:
        s = torch.cuda.Stream(device=device)
        _copy_streams[device] = s
    return s


_last_graph = {'ref': None, 'dev': None}      # the device copy of the most recent graph (for the dbscan drop-in)
_graph_listeners = []                         # callables run whenever a new graph is built (the dbscan cache resets)


def device_graph_of(graph):
    """(dist, idx) CUDA tensors of `graph` if it is the csr_matrix the last kneighbors_graph call returned and its
    device copy was kept (keep_device=True), else None."""
    ref = _last_graph['ref']
    if ref is not None and ref() is graph:
        return _last_graph['dev']
    return None


def _csr_from_arrays(data, indices, N, k):
    """csr_matrix over existing arrays without scipy's constructor pass (it would down-cast the int64 indices scikit-
    learn hands back, neighbors/_base.py:1031-1041, and copy 1.6 GB at 1M x 200)."""
    from scipy.sparse import csr_matrix
    g = csr_matrix((N, N), dtype=np.float64)
    g.data, g.indices = data, indices
    g.indptr = np.arange(0, N * k + 1, k, dtype=np.int64)
    return g


def kneighbors_graph(X, n_neighbors, *, mode='distance', metric='minkowski', p=2, metric_params=None,
                     include_self=False, n_jobs=None, algo='auto', devices=None, n_blocks=4, keep_device=False,
                     return_stats=False):
    """Same signature as sklearn.neighbors.kneighbors_graph for the arguments
    the reference uses (cluster.py:94-105, long_read_cluster.py:108-113): mode='distance', p=2, n_jobs
    ignored (it is ignored by sklearn's brute path as well).  Returns csr_matrix: float64 data (float32-rounded values
    for float32 input), int64 indices in ascending-distance order, indptr = arange * k.

    One library call per device (`sbk_knn_graph`): the device widens every finished row chunk to (float64, int64)
    and downloads it into the pinned arrays the csr_matrix is built over while the next chunk is computed; the
    host converts nothing.  `devices` (default: the SBK_DEVICES environment variable, else the current device)
    spreads the query rows over several GPUs of the box from this ONE process: one host thread and stream pair per
    device, the matrix replicated, each device writing its own row block of the same host arrays -- no collective.
    keep_device=True additionally keeps the (dist, idx) device tensors of a single-device build for
    long_read.dbscan_sweep (see device_graph_of)."""
    import torch
    if mode not in ('distance', 'connectivity'):
        raise ValueError("Unsupported mode, must be one of 'connectivity', or 'distance' but got '%s' instead" % mode)
    if metric not in ('minkowski', 'euclidean', 'l2') or (metric == 'minkowski' and p != 2) or metric_params:
        raise ValueError('semibin_b200.kneighbors_graph implements the Euclidean (p=2) metric only')
    if include_self not in (False, 'auto'):
        raise ValueError('include_self=True is not on the SemiBin path')
    _lib.require_cuda()
    for fn in _graph_listeners:
        fn()
    k = int(n_neighbors)
    if isinstance(X, np.ndarray):
        if X.dtype not in (np.float32, np.float64):
            X = X.astype(np.float64)
        Xt = torch.from_numpy(np.ascontiguousarray(X))
    elif isinstance(X, torch.Tensor):
        Xt = X if X.dtype in (torch.float32, torch.float64) else X.double()
        Xt = Xt.contiguous()
    else:
        raise TypeError('X must be a numpy array or a torch tensor')
    if Xt.dim() != 2:
        raise ValueError('X must be 2-D')
    N = Xt.shape[0]
    if not (0 < k < N):
        raise ValueError(f'Expected n_neighbors < n_samples_fit, but n_neighbors = {k}, n_samples_fit = {N}, '
                         f'n_samples = {N}')
    import time
    t_start = time.perf_counter()
    devs = resolve_devices(devices, Xt)
    # pinned result arrays: the csr_matrix is built over them (torch caches freed pinned blocks, so only the first
    # graph of a size pays for page-locking)
    indices = torch.empty((N * k,), dtype=torch.int64, pin_memory=True)
    data = torch.empty((N * k,), dtype=torch.float64, pin_memory=True) if mode == 'distance' else None
    t_alloc = time.perf_counter()
    stats = graph_rows_multi(Xt, k, devs, data, indices, algo=algo, n_blocks=n_blocks,
                             keep_device=keep_device and len(devs) == 1)
    data_np = data.numpy() if data is not None else np.ones(N * k, dtype=np.float64)
    g = _csr_from_arrays(data_np, indices.numpy(), N, k)
    import weakref
    dev_pair = stats.pop('device_graph', None) if isinstance(stats, dict) else None
    _last_graph['ref'], _last_graph['dev'] = (weakref.ref(g), dev_pair) if dev_pair is not None else (None, None)
    if isinstance(stats, dict):          # host wall-clock split of this call (ms)
        stats['host_ms'] = dict(alloc_result=(t_alloc - t_start) * 1e3, build=(time.perf_counter() - t_alloc) * 1e3)
    return (g, stats) if return_stats else g


def resolve_devices(devices, X=None):
    """Device list of a single-process multi-GPU call: explicit list / count, else SBK_DEVICES (a count or a comma
    separated list of ordinals), else the device X lives on, else the current device."""
    import os
    import torch
    if devices is None:
        env = os.environ.get('SBK_DEVICES', '').strip()
        if env:
            devices = [int(t) for t in env.split(',')] if ',' in env else int(env)
    if devices is None:
        if X is not None and isinstance(X, torch.Tensor) and X.is_cuda:
            return [X.device.index]
        return [torch.cuda.current_device()]
    if isinstance(devices, int):
        devices = list(range(devices))
    devs = [torch.device(d).index if not isinstance(d, int) else d for d in devices]
    n = torch.cuda.device_count()
    if not devs or any(d is None or d < 0 or d >= n for d in devs) or len(set(devs)) != len(devs):
        raise ValueError(f'devices={devices!r}: need distinct CUDA ordinals below {n}')
    return devs


def _upload_shared(Xt, device, share):
    """The host matrix on `device` when several devices of one process need it: every device uploads only its own row
    slice over its own PCIe link and fetches the other slices from its peers over NVLink (G full uploads of one host
    matrix are bound by host memory, not by the links: 8 x 400 MB took 30 ms).  `share` = (rank, slices, tensors,
    barrier) common to the device threads."""
    import torch
    rank, slices, tensors, barrier = share
    Xd = torch.empty(tuple(Xt.shape), dtype=Xt.dtype, device=device)
    lo, hi = slices[rank]
    Xd[lo:hi].copy_(Xt[lo:hi], non_blocking=True)
    torch.cuda.current_stream(device).synchronize()
    tensors[rank] = Xd
    barrier.wait()                                   # every slice is on its device
    for r, (rlo, rhi) in enumerate(slices):
        if r != rank and rhi > rlo:
            Xd[rlo:rhi].copy_(tensors[r][rlo:rhi], non_blocking=True)
    torch.cuda.current_stream(device).synchronize()
    barrier.wait()                                   # nobody's slice is read any more
    return Xd


def _graph_rows_one(Xt, k, dev, lo, hi, data, indices, algo, n_blocks, keep_device, out, share=None):
    """Rows [lo, hi) of the graph on device `dev`, written into rows [lo, hi) of the shared host arrays."""
    import torch
    try:
        lib = _lib.load()
        with torch.cuda.device(dev):
            device = torch.device('cuda', dev)
            import time
            t0 = time.perf_counter()
            if share is not None and not Xt.is_cuda:
                Xd = _upload_shared(Xt, device, share)
            else:
                Xd = Xt.to(device, non_blocking=True) if not (Xt.is_cuda and Xt.device == device) else Xt
                if not Xt.is_cuda:
                    torch.cuda.current_stream(device).synchronize()
            t_up = (time.perf_counter() - t0) * 1e3
            N, D = Xd.shape
            n_q = hi - lo
            dtype = _lib.SBK_F32 if Xd.dtype == torch.float32 else _lib.SBK_F64
            a = _lib.ALGOS[algo] if isinstance(algo, str) else int(algo)
            ws_bytes = lib.sbk_knn_graph_workspace_bytes(dtype, N, D, n_q, k, a)
            if ws_bytes == 0:
                raise ValueError(_lib.last_error())
            ws = _workspace(ws_bytes, device)
            copy = _copy_stream(device)
            wave_rows = (torch.cuda.get_device_properties(device).multi_processor_count // 2) * 256
            chunk = plan_chunk_rows(n_q, n_blocks, wave_rows)
            st = _lib.KnnStats()
            hd = None if data is None else C.c_void_p(data.data_ptr() + lo * k * 8)
            hi_ptr = C.c_void_p(indices.data_ptr() + lo * k * 8)
            t1 = time.perf_counter()
            rc = lib.sbk_knn_graph(_lib.ptr(Xd), dtype, N, D, Xd.stride(0), lo, hi, k, hd, hi_ptr, chunk,
                                   _lib.ptr(ws), ws.numel(), a, C.byref(st), _lib.stream_ptr(device), copy.cuda_stream)
            _lib.check(rc)
            res = st.as_dict()
            res['upload_ms'] = t_up
            res['setup_ms'] = (t1 - t0) * 1e3 - t_up
            res['call_ms'] = (time.perf_counter() - t1) * 1e3
            if keep_device:
                # the device outputs sit at the head of the workspace (graph_carve in api.cu): copy them out
                esz = 4 if dtype == _lib.SBK_F32 else 8
                nd = n_q * k
                dist = ws[:nd * esz].view(Xd.dtype).reshape(n_q, k).clone()
                off = (nd * esz + 255) // 256 * 256
                idx = ws[off:off + nd * 4].view(torch.int32).reshape(n_q, k).clone()
                res['device_graph'] = (dist, idx)
            out[dev] = res
    except BaseException as e:       # re-raised by the caller (a thread must not swallow it)
        if share is not None:
            share[3].abort()         # the other device threads must not wait for this one at the upload barrier
        out[dev] = e


def graph_rows_multi(Xt, k, devs, data, indices, *, algo='auto', n_blocks=4, keep_device=False):
    """Contiguous row blocks of the graph, one per device of `devs`, all written into the same host arrays.
    One device: a plain call.  Several: one host thread per device (ctypes releases the GIL for the library call)."""
    from .dist import row_block
    N = Xt.shape[0]
    out = {}
    if len(devs) == 1:
        _graph_rows_one(Xt, k, devs[0], 0, N, data, indices, algo, n_blocks, keep_device, out)
    else:
        import threading
        if not Xt.is_cuda and not Xt.is_pinned():
            Xt = Xt.pin_memory()              # G concurrent uploads of one host matrix: pin it once
        threads = []
        slices = [row_block(N, r, len(devs)) for r in range(len(devs))]
        tensors = [None] * len(devs)
        barrier = threading.Barrier(len(devs))
        for r, dev in enumerate(devs):
            lo, hi = slices[r]
            t = threading.Thread(target=_graph_rows_one, args=(Xt, k, dev, lo, hi, data, indices, algo, n_blocks, False, out,
                                                               (r, slices, tensors, barrier)))
            t.start()
            threads.append(t)
        for t in threads:
            t.join()
    for dev in devs:
        if isinstance(out.get(dev), BaseException):
            raise out[dev]
    if len(devs) == 1:
        return out[devs[0]]
    agg = {'per_device': {d: out[d] for d in devs}}
    for key in ('n_fallback_rows', 'n_candidates', 'n_reranked', 'n_launches'):
        agg[key] = sum(out[d][key] for d in devs)
    return agg
