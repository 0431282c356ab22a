"""Row-block sharding of the kNN graph build across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink; gloo in the CPU
tests).  The reference matrix is replicated; rank r computes query rows
[r*N/G, (r+1)*N/G) and the fixed-size (idx int32, dist) blocks are all-gathered.
Rank-local row-major order + ascending rank order = global row-major order, so
nothing is re-sorted.  The only other exchange on the
path is the 20-bin all-reduce of the max_node threshold counts
(SemiBin/cluster.py:122).
"""
import numpy as np


def row_block(n, rank, world):
    """Contiguous, balanced split: the first n % world ranks get one extra row."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist


def world_info(group=None):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def all_gather_rows(local, n_total, group=None, out=None):
    """Gather row blocks of possibly different heights (differ by at most one
    row) into the full [n_total, ...] tensor on every rank.  Equal blocks go
    straight into the result (`out`, if given, is reused) with one collective and
    no staging copy."""
    import torch
    dist = _dist()
    rank, world = world_info(group)
    if world == 1:
        return local
    sizes = [row_block(n_total, r, world) for r in range(world)]
    max_rows = max(hi - lo for lo, hi in sizes)
    if all(hi - lo == max_rows for lo, hi in sizes):
        if out is None:
            out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = [out[r * max_rows: r * max_rows + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)


def knn_sharded(X, n_neighbors, *, group=None, algo='auto', gather=True, local_knn=None):
    """Every rank holds the full X; returns the full (dist, idx) on every rank
    (gather=True) or only this rank's row block (gather=False).
    `local_knn(X, k, q_begin, q_end)` defaults to the libsbk kernel path."""
    rank, world = world_info(group)
    n = X.shape[0]
    lo, hi = row_block(n, rank, world)
    if local_knn is None:
        from .neighbors import knn
        d, i = knn(X, n_neighbors, q_begin=lo, q_end=hi, algo=algo)
    else:
        d, i = local_knn(X, n_neighbors, lo, hi)
    if not gather or world == 1:
        return d, i
    return all_gather_rows(d, n, group), all_gather_rows(i, n, group)


class PeerResult:
    """Full [n, k] (dist, idx) result buffers of every rank, each mapped into every process of the group (torch
    symmetric memory: CUDA virtual-memory handles exchanged once at rendezvous).  `sbk_knn_peers` makes the rerank
    kernels store every finished row into all of them with P2P stores over NVLink, so the row-block gather rides on
    the compute kernel; the only thing left after the call is a barrier."""

    def __init__(self, n, k, dtype, device, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = world_info(group)
        self.n, self.k = n, k
        self.dist = symm.empty((n, k), dtype=dtype, device=device)
        self.idx = symm.empty((n, k), dtype=torch.int32, device=device)
        hd = symm.rendezvous(self.dist, self.group)
        hi = symm.rendezvous(self.idx, self.group)
        self._handles = (hd, hi)
        peers = [r for r in range(self.world) if r != self.rank]
        self.n_peers = len(peers)
        self.peer_dist = (C.c_void_p * max(1, len(peers)))(*[int(hd.buffer_ptrs[r]) for r in peers])
        self.peer_idx = (C.c_void_p * max(1, len(peers)))(*[int(hi.buffer_ptrs[r]) for r in peers])


def knn_sharded_p2p(X, n_neighbors, bufs, *, algo='auto', return_stats=False):
    """Row-block sharded kNN graph with the gather fused into the rerank kernels: every rank computes its row block
    [lo, hi) straight into rows [lo, hi) of its own full buffer AND of every peer's (bufs: PeerResult); after the
    barrier every rank holds the full (dist, idx).  Returns (bufs.dist, bufs.idx)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from . import _lib
    from .neighbors import _workspace
    lib = _lib.load()
    n, d = X.shape
    k = int(n_neighbors)
    lo, hi = row_block(n, bufs.rank, bufs.world)
    dtype = _lib.SBK_F32 if X.dtype == torch.float32 else _lib.SBK_F64
    a = _lib.ALGOS[algo] if isinstance(algo, str) else int(algo)
    ws_bytes = lib.sbk_knn_workspace_bytes(dtype, n, d, hi - lo, k, a)
    if ws_bytes == 0:
        raise ValueError(_lib.last_error())
    ws = _workspace(ws_bytes, X.device)
    st = _lib.KnnStats()
    out_d, out_i = bufs.dist[lo:hi], bufs.idx[lo:hi]
    rc = lib.sbk_knn_peers(_lib.ptr(X), dtype, n, d, X.stride(0), lo, hi, k, _lib.ptr(out_d), _lib.ptr(out_i),
                           bufs.n_peers, bufs.peer_dist, bufs.peer_idx, _lib.ptr(ws), ws.numel(), a, C.byref(st),
                           _lib.stream_ptr(X.device))
    _lib.check(rc)
    # the call returns with this rank's stream drained (all its P2P stores issued and complete); once every rank
    # has got here, every full buffer is complete
    dist.barrier(group=bufs.group)
    if return_stats:
        return bufs.dist, bufs.idx, st.as_dict()
    return bufs.dist, bufs.idx


class SymShard:
    """Exchange buffers of the sharded symmetric build (`sbk_knn_sym_shard`): per rank the thresholds of all positions,
    the record counts and one mailbox region per source rank, allocated as torch symmetric memory so that every rank
    can store into every other rank's buffer over NVLink.  Raises ValueError when the symmetric filter does not take a
    build of this shape (the caller then shards by row blocks)."""

    def __init__(self, n, d, k, dtype, device, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        lib = _lib.load()
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = world_info(group)
        self.shape = (n, d, k, dtype)
        code = _lib.SBK_F32 if dtype == torch.float32 else _lib.SBK_F64
        nbytes = lib.sbk_knn_sym_shard_xchg_bytes(code, n, d, k, self.world)
        if nbytes == 0:
            raise ValueError('the symmetric filter does not take this build on %d GPUs' % self.world)
        self.buf = symm.empty((int(nbytes),), dtype=torch.uint8, device=device)
        self._handle = symm.rendezvous(self.buf, self.group)
        self.xchg = (C.c_void_p * self.world)(*[int(self._handle.buffer_ptrs[r]) for r in range(self.world)])


def knn_sharded_sym(X, n_neighbors, bufs, shard, *, return_stats=False):
    """The whole kNN graph on all GPUs of the group with every tile pair multiplied ONCE across them (symmetric filter;
    `knn_sharded_p2p` multiplies every pair twice: once for either row's owner).  Every rank owns a block of staged
    positions, sends the candidates it finds for other ranks' rows into their mailboxes (shard: SymShard), reranks its
    own rows and stores them into every rank's full result (bufs: PeerResult).  Returns (bufs.dist, bufs.idx), complete
    on every rank."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from . import _lib
    from .neighbors import _workspace
    lib = _lib.load()
    n, d = X.shape
    k = int(n_neighbors)
    if (n, d, k, X.dtype) != shard.shape:
        raise ValueError('SymShard was made for another problem shape')
    dtype = _lib.SBK_F32 if X.dtype == torch.float32 else _lib.SBK_F64
    ws_bytes = lib.sbk_knn_workspace_bytes(dtype, n, d, n, k, _lib.KNN_TENSOR_SYMMETRIC)
    if ws_bytes == 0:
        raise ValueError(_lib.last_error())
    ws = _workspace(ws_bytes, X.device)
    st = _lib.KnnStats()
    group = bufs.group

    flag = torch.zeros(1, dtype=torch.int32, device=X.device)

    def vote(failed):
        # barrier + "did anybody fail": one small MAX all-reduce, read back
        flag.fill_(1 if failed else 0)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        return int(flag.item())

    @C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int)
    def barrier(_arg, failed):
        return vote(failed)

    rc = lib.sbk_knn_sym_shard(_lib.ptr(X), dtype, n, d, X.stride(0), k, shard.rank, shard.world, shard.xchg,
                               C.cast(barrier, C.c_void_p), None, _lib.ptr(bufs.dist), _lib.ptr(bufs.idx),
                               bufs.n_peers, bufs.peer_dist, bufs.peer_idx, _lib.ptr(ws), ws.numel(), C.byref(st),
                               _lib.stream_ptr(X.device))
    # the call returns with this rank's stream drained; once every rank has got here every full buffer is complete and
    # nobody reads its mailboxes any more.  The same vote tells every rank whether the build held everywhere (a pool or
    # mailbox that was too small fails it on all ranks together): if not, all ranks shard by row blocks instead.
    if rc not in (_lib.SBK_OK, _lib.SBK_ERR_INTERNAL):
        _lib.check(rc)
    why = _lib.last_error() if rc else ''
    if vote(rc != _lib.SBK_OK):
        out = knn_sharded_p2p(X, k, bufs, algo='tensor_onesided', return_stats=True)
        out[2]['sym_shard_fallback'] = why or 'another rank failed'
        return out if return_stats else out[:2]
    if return_stats:
        return bufs.dist, bufs.idx, st.as_dict()
    return bufs.dist, bufs.idx


def allreduce_counts(counts, group=None):
    """Sum of the per-threshold row counts over ranks (cluster.py:122)."""
    dist = _dist()
    rank, world = world_info(group)
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def gather_edges(X, Y, V, group=None):
    """Concatenate per-rank edge blocks in rank order (= global row-major order)."""
    import torch
    dist = _dist()
    rank, world = world_info(group)
    if world == 1:
        return X, Y, V
    n_local = torch.tensor([X.shape[0]], dtype=torch.int64, device=X.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts) if counts else 0
    outs = []
    for t in (X, Y, V):
        pad = torch.zeros((m,), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        buf = torch.empty((world * m,), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, pad, group=group)
        outs.append(torch.cat([buf[r * m: r * m + counts[r]] for r in range(world)]))
    return tuple(outs)


def build_edge_list_sharded(embedding, features, depth, *, max_edges, max_node, is_combined, n_sample,
                            group=None, algo='auto', to_host='all'):
    """Multi-GPU version of cluster.build_edge_list: kNN graphs, intersection
    and edge weights on this rank's row block; one all-reduce (threshold
    counts) and one gather (edge blocks).  `depth` may be a CUDA tensor.
    to_host='all': fresh host arrays (X, Y, V) on every rank; 'rank0': on rank 0 only (others get None);
    'view': on every rank, as views into cached pinned staging buffers (several times faster than pageable
    copies, but only valid until the next call); False: the gathered CUDA tensors."""
    import torch
    from . import cluster
    from .neighbors import knn, _as_device_matrix
    rank, world = world_info(group)
    emb = _as_device_matrix(embedding).float()
    feat = _as_device_matrix(features)
    n = emb.shape[0]
    k = min(int(max_edges), n - 1)
    lo, hi = row_block(n, rank, world)
    a_d, a_i = knn(emb, k, q_begin=lo, q_end=hi, algo=algo)
    b_d, b_i = knn(feat, k, q_begin=lo, q_end=hi, algo=algo)
    keep, rowmax, counts = cluster.edges_stage1(a_d, a_i, b_d, b_i, cluster.threshold_sequence())
    counts = allreduce_counts(counts.contiguous(), group)
    threshold = cluster.pick_threshold(counts.cpu().numpy(), n, max_node)
    dep = None
    if not is_combined:
        dep = depth.float().contiguous().cuda() if isinstance(depth, torch.Tensor) else \
            torch.as_tensor(np.ascontiguousarray(depth, dtype=np.float32)).cuda()
    X, Y, V = cluster.edges_stage2(keep, a_d, a_i, lo, threshold, dep, n_sample, is_combined)
    X, Y, V = gather_edges(X, Y, V, group)
    if to_host is False:
        return X, Y, V
    if to_host == 'rank0' and rank != 0:
        return None
    return cluster.edges_to_host(X, Y, V, 'view' if to_host == 'view' else 'copy')
