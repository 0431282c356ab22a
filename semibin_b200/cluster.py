"""Short-read clustering front end: drop-in for SemiBin/cluster.py:64-168
(`run_embed_infomap`) with the graph build on the GPU.

    run_embed_infomap(logger, model, data, *, device, max_edges, max_node,
                      is_combined, n_sample, contig_dict, num_process, random_seed)
        -> (embedding float32 [N,100], contig_labels int [N])        cluster.py:64-66,168
    build_edge_list(embedding, features, depth, *, max_edges, max_node,
                    is_combined, n_sample) -> (X int32[E], Y int32[E], V float64[E])
        the exact (edges, edge_weights) hand-off of cluster.py:144-151, row-major
        sorted, Y > X, without materialising Python tuples.

The embedding MLP forward stays in the reference model (north_star); the two
kNN graphs, their intersection, the 1-min(d,1) transform, the max_node
threshold, the KL/abundance factor (cal_kl evaluated at edges only) and the
edge compaction run in libsbk (include/sbk.h).
"""
import ctypes as C
import numpy as np

from . import _lib
from .neighbors import knn, _as_device_matrix, _workspace


def threshold_sequence():
    """Values `threshold` takes in cluster.py:118-125 (repeated += 0.05 in
    float64 -- the drift is part of the reference semantics)."""
    out = []
    t = 0
    while t < 1:
        t += 0.05
        out.append(t)
    return out


def pick_threshold(counts, num_contigs, max_node):
    """cluster.py:118-126 given counts[m] = #(rowmax > threshold_m).  numpy
    scalar types are kept so that round() is numpy's, as in the reference."""
    threshold = 0
    m = 0
    while threshold < 1:
        threshold += 0.05
        n_above = np.int64(counts[m])
        m += 1
        if round(n_above / num_contigs, 2) < max_node:
            break
    threshold -= 0.05
    return threshold


def edges_stage1(emb_d, emb_i, kmer_d, kmer_i, thresholds):
    """Device tensors in; returns (keep_bits int32 [n, ceil(k/32)], rowmax [n] f64, counts int64[len(thresholds)]).
    keep_bits marks the entries of the embedding graph that survive the intersection (cluster.py:109-113)."""
    import torch
    lib = _lib.load()
    n, k = emb_i.shape
    dev = emb_i.device
    thr = torch.tensor(thresholds, dtype=torch.float64, device=dev)
    keep = torch.empty((n, (k + 31) // 32), dtype=torch.int32, device=dev)
    rowmax = torch.empty((n,), dtype=torch.float64, device=dev)
    counts = torch.zeros((_lib.MAX_THRESHOLDS,), dtype=torch.int64, device=dev)
    kd = kmer_d
    if kd is not None and kd.dtype != torch.float64:
        kd = kd.double()
    rc = lib.sbk_edges_stage1(_lib.ptr(emb_d), _lib.ptr(emb_i), _lib.ptr(kd), _lib.ptr(kmer_i), n, k,
                              _lib.ptr(thr), len(thresholds), _lib.ptr(keep), _lib.ptr(rowmax), _lib.ptr(counts),
                              _lib.stream_ptr(dev))
    _lib.check(rc)
    return keep, rowmax, counts[:len(thresholds)]


def edges_stage2(keep, emb_d, emb_i, row_begin, threshold, depth, n_sample, is_combined, capacity=None):
    """Returns device tensors (X int32, Y int32, V float64) for the local rows.  depth: [n_total, 2*n_sample]
    float32 for ALL contigs (columns of emb_i are global indices), unused when is_combined."""
    import torch
    lib = _lib.load()
    n, k = emb_i.shape
    dev = emb_i.device
    if capacity is None:
        capacity = n * k
    n_total = 0 if depth is None else int(depth.shape[0])
    ws_bytes = lib.sbk_edges_stage2_workspace_bytes(n, k, n_total, 0 if is_combined else int(n_sample))
    ws = _workspace(ws_bytes, dev)
    ox = torch.empty((capacity,), dtype=torch.int32, device=dev)
    oy = torch.empty((capacity,), dtype=torch.int32, device=dev)
    ov = torch.empty((capacity,), dtype=torch.float64, device=dev)
    ne = torch.zeros((1,), dtype=torch.int64, device=dev)
    rc = lib.sbk_edges_stage2(_lib.ptr(keep), _lib.ptr(emb_d), _lib.ptr(emb_i), row_begin, n, k, float(threshold),
                              _lib.ptr(depth), n_total, int(n_sample), int(bool(is_combined)),
                              _lib.ptr(ox), _lib.ptr(oy), _lib.ptr(ov), capacity, _lib.ptr(ne),
                              _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
    _lib.check(rc)
    e = int(ne.item())
    if e > capacity:
        raise _lib.SbkError(f'edge list capacity {capacity} < {e}')
    return ox[:e], oy[:e], ov[:e]


_pinned = {}


def to_host_view(t, tag):
    """Device tensor -> numpy VIEW of a cached pinned staging buffer (valid until the next call with the same
    tag).  Pageable copies of a 0.5 GB edge list run at a fraction of the PCIe rate."""
    import torch
    n = t.numel()
    buf = _pinned.get((tag, t.dtype))
    if buf is None or buf.numel() < n:
        buf = torch.empty((max(n, 1),), dtype=t.dtype).pin_memory()
        _pinned[(tag, t.dtype)] = buf
    buf[:n].copy_(t.reshape(-1), non_blocking=True)
    return buf[:n]


def edges_to_host(X, Y, V, mode):
    """mode 'copy': fresh arrays; 'view': views into cached pinned buffers (valid until the next call)."""
    import torch
    if mode == 'view':
        bufs = [to_host_view(t, tag) for t, tag in ((X, 'x'), (Y, 'y'), (V, 'v'))]
        torch.cuda.current_stream().synchronize()
        return tuple(b.numpy() for b in bufs)
    return X.cpu().numpy(), Y.cpu().numpy(), V.cpu().numpy()


def _edge_rows(dev, lo, hi, emb, feat, depth, *, k, max_node, is_combined, n_sample, algo, reduce_counts,
               want_graphs):
    """Rows [lo, hi) of the edge list on CUDA device `dev` (inputs are host or device tensors; they are brought to
    `dev`).  `reduce_counts(counts_host int64[20]) -> global counts` is the one exchange of the path
    (cluster.py:122: the max_node threshold counts all rows).  Returns device tensors (X, Y, V) and extras."""
    import torch
    device = torch.device('cuda', dev)
    with torch.cuda.device(device):
        emb_d = emb.to(device, non_blocking=True)
        feat_d = feat.to(device, non_blocking=True)
        n = emb_d.shape[0]
        a_d, a_i = knn(emb_d, k, q_begin=lo, q_end=hi, algo=algo)                 # cluster.py:94-99
        b_d, b_i = knn(feat_d, k, q_begin=lo, q_end=hi, algo=algo)                # :100-105
        ths = threshold_sequence()
        keep, rowmax, counts = edges_stage1(a_d, a_i, b_d, b_i, ths)             # :109-122
        counts = reduce_counts(counts.cpu().numpy())
        threshold = pick_threshold(counts, n, max_node)                          # :118-126
        dep = None
        if not is_combined:
            dep = depth.to(device, non_blocking=True)
        X, Y, V = edges_stage2(keep, a_d, a_i, lo, threshold, dep, n_sample, is_combined)   # :115-116, :127-151
        extra = dict(threshold=threshold)
        if want_graphs:
            extra.update(emb=(a_d, a_i), kmer=(b_d, b_i), rowmax=rowmax)
        return X, Y, V, extra


def build_edge_list(embedding, features, depth, *, max_edges, max_node, is_combined, n_sample,
                    algo='auto', return_graphs=False, to_host='copy', devices=None):
    """Fused GPU path for cluster.py:94-151.

    embedding : [N,100] float32      (model.embedding output, cluster.py:93)
    features  : [N,136] float64, or [N,136+S] when is_combined (already normalised
                as in cluster.py:81-86)
    depth     : [N,2S] float32 interleaved [mean_s,var_s] (cluster.py:88); unused
                when is_combined
    devices   : CUDA ordinals to spread the query rows over from this ONE process (default: SBK_DEVICES, else the
                device of a CUDA `embedding`, else the current device).  Every device holds the whole matrices and
                builds the edges of its own row block (one host thread per device); the only exchange is the sum of
                the 20 threshold counts; the per-device edge blocks concatenate in row order.
    Returns host numpy arrays (X int32[E], Y int32[E], V float64[E]); to_host='view' hands out views into
    cached pinned staging buffers instead of fresh arrays (valid until the next call), False the CUDA tensors
    (single device only)."""
    import torch
    from .neighbors import resolve_devices
    from .dist import row_block
    _lib.require_cuda()

    def as_tensor(x, want=None):
        if isinstance(x, np.ndarray):
            if x.dtype not in (np.float32, np.float64):
                x = x.astype(np.float64)
            x = torch.from_numpy(np.ascontiguousarray(x))
        if not isinstance(x, torch.Tensor):
            raise TypeError('numpy array or torch tensor expected')
        if want is not None and x.dtype != want:
            x = x.to(want)
        elif x.dtype not in (torch.float32, torch.float64):
            x = x.double()
        return x.contiguous()
    emb = as_tensor(embedding, torch.float32)
    feat = as_tensor(features)
    N = emb.shape[0]
    k = min(int(max_edges), N - 1)                       # cluster.py:96
    dep = None
    if not is_combined:
        dep = as_tensor(depth, torch.float32)
        if tuple(dep.shape) != (N, 2 * n_sample):
            raise ValueError(f'depth must be [N, 2*n_sample] = {(N, 2 * n_sample)}, got {tuple(dep.shape)}')
    devs = resolve_devices(devices, emb)
    kw = dict(k=k, max_node=max_node, is_combined=is_combined, n_sample=n_sample, algo=algo)
    if len(devs) == 1:
        X, Y, V, extra = _edge_rows(devs[0], 0, N, emb, feat, dep, reduce_counts=lambda c: c,
                                    want_graphs=return_graphs, **kw)
        with torch.cuda.device(devs[0]):
            out = (X, Y, V) if to_host is False else edges_to_host(X, Y, V, to_host)
        return (out, extra) if return_graphs else out
    if to_host is False:
        raise ValueError('to_host=False needs a single device')
    import threading
    G = len(devs)
    barrier = threading.Barrier(G)
    partial = [None] * G
    results = [None] * G

    def reduce_counts_for(r):
        def reduce(c):
            partial[r] = c
            barrier.wait()
            tot = np.sum(np.stack(partial), axis=0)
            barrier.wait()                       # nobody overwrites `partial` before everyone has summed
            return tot
        return reduce

    def work(r):
        try:
            lo, hi = row_block(N, r, G)
            X, Y, V, extra = _edge_rows(devs[r], lo, hi, emb, feat, dep, reduce_counts=reduce_counts_for(r),
                                        want_graphs=False, **kw)
            with torch.cuda.device(devs[r]):
                results[r] = tuple(t.cpu().numpy() for t in (X, Y, V)) + (extra,)
        except BaseException as e:
            results[r] = e
            barrier.abort()
    # G concurrent uploads of each host matrix: pin them once
    emb, feat = (t if t.is_cuda or t.is_pinned() else t.pin_memory() for t in (emb, feat))
    if dep is not None and not dep.is_cuda and not dep.is_pinned():
        dep = dep.pin_memory()
    threads = [threading.Thread(target=work, args=(r,)) for r in range(G)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for r in results:
        if isinstance(r, BaseException) and not isinstance(r, threading.BrokenBarrierError):
            raise r
    for r in results:
        if isinstance(r, BaseException):
            raise r
    out = tuple(np.concatenate([results[r][c] for r in range(G)]) for c in range(3))   # rank order = row-major order
    if return_graphs:
        return out, dict(threshold=results[0][3]['threshold'])
    return out


def add_edges_contiguous(graph, X, Y):
    """Hand-off #1 to igraph without the reference's per-edge Python tuples (cluster.py:150 builds up to 1e8
    of them): the endpoints go in as ONE contiguous int64 [E, 2] buffer (python-igraph reads a 2-D memoryview of
    its integer type directly); an igraph build without that path gets the same pairs as nested lists, built by
    numpy in C.  Edge order and endpoints are unchanged."""
    pairs = np.empty((len(X), 2), dtype=np.int64)
    pairs[:, 0] = X
    pairs[:, 1] = Y
    try:
        graph.add_edges(memoryview(pairs))
    except (TypeError, ValueError):
        graph.add_edges(pairs.tolist())
    return pairs


def run_embed_infomap(logger, model, data, *, device, max_edges, max_node, is_combined, n_sample,
                      contig_dict, num_process: int, random_seed):
    """Drop-in for SemiBin.cluster.run_embed_infomap (same arguments, same
    returns).  Infomap itself stays in igraph (consumer of the hand-off)."""
    import torch
    from igraph import Graph                                     # cluster.py:72 (consumer, out of scope)
    NR_INFOMAP_TRIALS = 10                                       # cluster.py:11

    train_data_input = data.values[:, 0:136] if not is_combined else data.values   # :79
    if is_combined:
        n = train_data_input.shape[1] - 136                      # utils.py:630-636 norm_abundance
        if n >= 20 or (n >= 5 and np.mean(np.sum(train_data_input[:, 136:], axis=1)) > 2):
            norm = np.sum(train_data_input, axis=0)              # cluster.py:83-86
            train_data_input = train_data_input / norm
            train_data_input = train_data_input / np.abs(train_data_input).sum(axis=1, keepdims=True)
    depth = data.values[:, 136:len(data.values[0])].astype(np.float32)              # :88
    num_contigs = train_data_input.shape[0]
    with torch.no_grad():
        model.eval()
        x = torch.from_numpy(np.ascontiguousarray(train_data_input)).to(device)
        emb_t = model.embedding(x.float()).detach()              # :90-93 (reference model)
        embedding = emb_t.cpu().numpy()
    # a CUDA embedding stays on ITS device (the `device` argument may name a non-current GPU); SBK_DEVICES spreads
    # the rows over several GPUs from this one process
    X, Y, V = build_edge_list(emb_t if emb_t.is_cuda else embedding, train_data_input, depth,
                              max_edges=max_edges, max_node=max_node, is_combined=is_combined,
                              n_sample=n_sample, to_host='view')      # consumed by igraph within this call
    logger.debug(f'Number of edges in clustering graph: {len(X)}')
    g = Graph()
    g.add_vertices(np.arange(num_contigs))
    add_edges_contiguous(g, X, Y)                                # cluster.py:150-157 without the tuple list
    length_weight = np.array([len(contig_dict[name]) for name in data.index])
    logger.debug(f'Running infomap with {num_process} processes...')
    result = g.community_infomap(edge_weights=V, vertex_weights=length_weight, trials=NR_INFOMAP_TRIALS)
    contig_labels = np.zeros(shape=num_contigs, dtype=int)
    for i, r in enumerate(result):
        contig_labels[r] = i
    return embedding, contig_labels
