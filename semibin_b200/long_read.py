"""Long-read clustering front end: drop-ins for the hot part of
SemiBin/long_read_cluster.py:51-161.

    dbscan(X, eps, *, min_samples, metric='precomputed', sample_weight)
        -> (core_sample_indices, labels)      same as sklearn.cluster.dbscan on the
        precomputed kNN graph (long_read_cluster.py:117-119)
    dbscan_sweep(graph, eps_list, *, min_samples, sample_weight) -> [labels ...]
        all eps values in one pass (the 12-value loop of :115-120)
    long_read_labels(embedding, depth, lengths, ...) -> [12 label vectors]
        features concat (:75-81) + kNN graph (:108-113) + sweep (:115-120)
    integrate_labels(labels, lengths, markers, minfasta) -> (contig_labels, n_bins)
        the greedy ensemble integration (get_best_bin + extraction loop, :12-49, :122-150)
    integrate_bins(dbscan_results, contig2marker, contig_list, contig_dict, minfasta) -> extracted
        same, on the reference's own Python objects (lists / dicts of names)
    cluster_long_read(...)   reference signature: the body of :51-161 with the kNN graph, the
        DBSCAN sweep and the integration loop on the GPU; marker-gene search and bin
        writing are the reference package's (out of scope).
"""
import ctypes as C
import numpy as np

from . import _lib
from .neighbors import knn, _as_device_matrix, _workspace

EPS_GRID = [0.01, 0.05, 0.1, 0.15, 0.2, 0.25, 0.3, 0.35, 0.4, 0.45, 0.5, 0.55]   # long_read_cluster.py:116
_MAX_EPS = 16


def _graph_arrays(graph):
    """csr_matrix with k entries per row (kneighbors_graph output) or a
    (dist, idx) pair -> CUDA tensors dist float32 [N,k], idx int32 [N,k]."""
    import torch
    if isinstance(graph, (tuple, list)):
        dist, idx = graph
        if isinstance(dist, np.ndarray):
            d32 = np.ascontiguousarray(dist, dtype=np.float32)
            if dist.dtype != np.float32 and not np.array_equal(d32.astype(dist.dtype), dist):
                raise NotImplementedError('dbscan_sweep: distances must be float32-representable')
            dist = torch.from_numpy(d32)
        if isinstance(idx, np.ndarray):
            idx = torch.from_numpy(np.ascontiguousarray(idx).astype(np.int32))
        if dist.dtype == torch.float64 and not torch.equal(dist.float().double(), dist):
            raise NotImplementedError('dbscan_sweep: distances must be float32-representable')
        return dist.float().contiguous().cuda(), idx.int().contiguous().cuda()
    from . import neighbors
    kept = neighbors.device_graph_of(graph)
    if kept is not None:                                   # the graph our kneighbors_graph drop-in just built: still on the device
        return kept[0].float(), kept[1]
    indptr = np.asarray(graph.indptr)
    n = graph.shape[0]
    k = int(indptr[1] - indptr[0]) if n else 0
    if not np.array_equal(indptr, np.arange(0, n * k + 1, k)):
        raise NotImplementedError('dbscan(metric="precomputed") is implemented for kneighbors_graph '
                                  'output (the same number of neighbours in every row)')
    data = np.asarray(graph.data)
    d32 = data.astype(np.float32)
    if data.dtype != np.float32 and not np.array_equal(d32.astype(data.dtype), data):
        # the sweep compares float32-valued distances (what kneighbors_graph yields for the float32 features of
        # long_read_cluster.py:75-81); a float64-valued graph would be rounded -> refuse instead of mislabelling
        raise NotImplementedError('dbscan(metric="precomputed"): graph distances must be float32-representable '
                                  '(build the graph from float32 features, as SemiBin does)')
    dist = torch.from_numpy(d32.reshape(n, k)).cuda()
    idx = torch.from_numpy(np.asarray(graph.indices).astype(np.int32).reshape(n, k)).cuda()
    return dist, idx


def radius_sweep(dist, idx, eps_list, *, min_samples, sample_weight=None, row_begin=0):
    """Device tensors (local rows) -> (prefix_len int32 [E,n], core uint8 [E,n])."""
    import torch
    lib = _lib.load()
    n, k = idx.shape
    E = len(eps_list)
    if not (1 <= E <= _MAX_EPS):
        raise ValueError(f'between 1 and {_MAX_EPS} eps values per sweep')
    dev = idx.device
    plen = torch.empty((E, n), dtype=torch.int32, device=dev)
    core = torch.empty((E, n), dtype=torch.uint8, device=dev)
    eps_arr = (C.c_double * E)(*[float(e) for e in eps_list])
    rc = lib.sbk_radius_sweep(_lib.ptr(dist), _lib.ptr(idx), row_begin, n, k, eps_arr, E,
                              _lib.ptr(sample_weight), float(min_samples), _lib.ptr(plen), _lib.ptr(core),
                              _lib.stream_ptr())
    _lib.check(rc)
    return plen, core


def dbscan_labels(idx, plen, core):
    import torch
    lib = _lib.load()
    n, k = idx.shape
    E = plen.shape[0]
    labels = torch.empty((E, n), dtype=torch.int64, device=idx.device)
    ws_bytes = lib.sbk_dbscan_labels_workspace_bytes(n, E)
    ws = _workspace(ws_bytes, idx.device)
    rc = lib.sbk_dbscan_labels(_lib.ptr(idx), n, k, _lib.ptr(plen), _lib.ptr(core), E, _lib.ptr(labels),
                               _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.check(rc)
    return labels


def dbscan_sweep(graph, eps_list, *, min_samples=5, sample_weight=None, return_core=False):
    """All eps values at once.  Returns a list of label arrays (np.intp, -1 =
    noise), bit-identical to calling sklearn.cluster.dbscan per eps."""
    import torch
    _lib.require_cuda()
    dist, idx = _graph_arrays(graph)
    sw = None
    if sample_weight is not None:
        sw = torch.from_numpy(np.ascontiguousarray(sample_weight, dtype=np.float64)).cuda()
    out, cores = [], []
    for s in range(0, len(eps_list), _MAX_EPS):
        chunk = list(eps_list[s:s + _MAX_EPS])
        plen, core = radius_sweep(dist, idx, chunk, min_samples=min_samples, sample_weight=sw)
        labels = dbscan_labels(idx, plen, core).cpu().numpy()
        core_h = core.cpu().numpy()
        for e in range(len(chunk)):
            out.append(labels[e].astype(np.intp))
            cores.append(np.where(core_h[e])[0])
    return (out, cores) if return_core else out


def dbscan(X, eps=0.5, *, min_samples=5, metric='precomputed', metric_params=None, algorithm='auto',
           leaf_size=30, p=2, sample_weight=None, n_jobs=None):
    """sklearn.cluster.dbscan for the call the reference makes
    (long_read_cluster.py:117-119): precomputed sparse kNN graph."""
    if metric != 'precomputed':
        raise ValueError('semibin_b200.dbscan implements metric="precomputed" (the SemiBin path) only')
    if not eps > 0.0:
        raise ValueError('eps must be positive.')
    labels, cores = dbscan_sweep(X, [eps], min_samples=min_samples, sample_weight=sample_weight, return_core=True)
    return cores[0], labels[0]


def long_read_features(embedding, depth, n_sample, is_combined):
    """long_read_cluster.py:72-81.  A CUDA embedding stays on its device (N3: the reference copies it to the host and
    back, long_read_cluster.py:70): only the S log-depth columns are computed on the host -- with numpy's float32 log,
    the reference's own evaluator, so the features are bit-identical -- and uploaded for the concat."""
    import torch
    on_device = isinstance(embedding, torch.Tensor) and embedding.is_cuda
    if is_combined:
        return embedding.float().contiguous() if on_device else np.ascontiguousarray(embedding, dtype=np.float32)
    depth = np.asarray(depth, dtype=np.float32)
    mean_index = [2 * temp for temp in range(n_sample)]
    logd = np.log(np.clip(depth[:, mean_index], 1e-6, None))
    if on_device:
        return torch.cat((embedding.float(), torch.from_numpy(logd).to(embedding.device)), dim=1).contiguous()
    if isinstance(embedding, torch.Tensor):
        embedding = embedding.numpy()
    return np.concatenate((embedding, logd), axis=1)


def long_read_labels(embedding, depth, lengths, *, n_sample, is_combined, eps_list=None, min_samples=5,
                     n_neighbors=200, algo='auto', return_graph=False):
    """Hand-off #2 of the long-read path: the 12 DBSCAN label vectors."""
    import torch
    _lib.require_cuda()
    x = long_read_features(embedding, depth, n_sample, is_combined)
    xd = x if (isinstance(x, torch.Tensor) and x.is_cuda) else _as_device_matrix(x)
    k = min(n_neighbors, xd.shape[0] - 1)                 # long_read_cluster.py:110
    with torch.cuda.device(xd.device):
        dist, idx = knn(xd, k, algo=algo)
        res = dbscan_sweep((dist, idx), EPS_GRID if eps_list is None else eps_list, min_samples=min_samples,
                           sample_weight=lengths)
    return (res, (dist, idx)) if return_graph else res


class _SweepCache:
    """Serves the reference's 12 sequential dbscan() calls (long_read_cluster.py:
    115-120) from ONE GPU sweep: the first call on a graph computes every eps of
    EPS_GRID, later calls on the same graph object are lookups.
    The cache holds STRONG references to the graph and the weights it was computed for and compares them by
    identity, so a later graph that happens to land on a freed address can never be served stale labels; it is
    dropped as soon as a different graph (the next sample) comes in, or by `reset()` (the kneighbors_graph drop-in
    calls it whenever it builds a new graph)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.graph = None
        self.weight = None
        self.min_samples = None
        self.res = {}

    def _same(self, X, min_samples, sample_weight):
        return self.graph is X and self.weight is sample_weight and self.min_samples == min_samples

    def __call__(self, X, eps=0.5, *, min_samples=5, metric='precomputed', sample_weight=None, n_jobs=None,
                 **kw):
        if metric != 'precomputed':
            raise ValueError('semibin_b200.dbscan implements metric="precomputed" only')
        same = self._same(X, min_samples, sample_weight)
        if not same or eps not in self.res:
            grid = EPS_GRID if eps in EPS_GRID else [eps]
            labels, cores = dbscan_sweep(X, grid, min_samples=min_samples, sample_weight=sample_weight,
                                         return_core=True)
            if not same:
                self.reset()
                self.graph, self.weight, self.min_samples = X, sample_weight, min_samples
            for e, lab, c in zip(grid, labels, cores):
                self.res[e] = (c, lab)
        return self.res[eps]


N_MARKERS = 107            # recall denominator, long_read_cluster.py:35


def marker_masks(markers, n=None):
    """list (per contig) of marker ids -> uint64 [N, words] bit masks (ids < 256)."""
    n = len(markers) if n is None else n
    top = max((int(np.max(m)) for m in markers if len(m)), default=0)
    words = top // 64 + 1
    if words > 4:
        raise ValueError(f'{top + 1} distinct markers: at most 256 are supported')
    masks = np.zeros((n, words), dtype=np.uint64)
    rows = np.repeat(np.arange(n), [len(m) for m in markers])
    if len(rows):
        ids = np.concatenate([np.asarray(m, dtype=np.int64) for m in markers if len(m)])
        np.bitwise_or.at(masks, (rows, ids // 64), np.uint64(1) << (ids % 64).astype(np.uint64))
    return masks


def integrate_labels(labels, lengths, markers, minfasta, n_markers=N_MARKERS):
    """Greedy ensemble integration on the GPU.  labels: [R, N] int64 (CUDA tensor or array; -1 noise),
    lengths: [N] ints, markers: per-contig marker-id lists or a uint64 [N, words] mask array.
    Returns (contig_labels np.int64 [N] with -1 = unbinned, number of bins), identical to
    long_read_cluster.py:122-150."""
    import torch
    _lib.require_cuda()
    lib = _lib.load()
    if not isinstance(labels, torch.Tensor):
        labels = torch.from_numpy(np.ascontiguousarray(np.asarray(labels, dtype=np.int64)))
    labels = labels.to(device='cuda', dtype=torch.int64).contiguous()
    R, N = labels.shape
    masks = markers if isinstance(markers, np.ndarray) and markers.dtype == np.uint64 else marker_masks(markers, N)
    if masks.shape[0] != N:
        raise ValueError('one marker list per contig expected')
    md = torch.from_numpy(np.ascontiguousarray(masks).view(np.int64)).cuda()
    ld = torch.from_numpy(np.ascontiguousarray(np.asarray(lengths, dtype=np.int64))).cuda()
    out = torch.empty((N,), dtype=torch.int64, device='cuda')
    ws_bytes = lib.sbk_integrate_workspace_bytes(N, R)
    if ws_bytes == 0:
        raise ValueError(f'integrate: unsupported shape R={R}, N={N}')
    ws = _workspace(ws_bytes, labels.device)
    n_ext = C.c_int32(0)
    rc = lib.sbk_integrate(_lib.ptr(labels), R, N, _lib.ptr(ld), _lib.ptr(md), masks.shape[1], int(n_markers),
                           int(minfasta), _lib.ptr(out), C.byref(n_ext), _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.check(rc)
    return out.cpu().numpy(), int(n_ext.value)


def integrate_bins(dbscan_results, contig2marker, contig_list, contig_dict, minfasta):
    """The reference's objects in, `extracted` (list of lists of contig names, in extraction order, names in
    contig_list order) out: long_read_cluster.py:122-143."""
    names = list(contig_list)
    marker_id = {}
    per_contig = []
    for name in names:
        ms = contig2marker[name] if name in contig2marker else []
        per_contig.append(np.array(sorted({marker_id.setdefault(m, len(marker_id)) for m in ms}), dtype=np.int64))
    lengths = np.array([len(contig_dict[name]) for name in names], dtype=np.int64)
    labels = np.asarray([np.asarray(r, dtype=np.int64) for r in dbscan_results])
    lab, n_ext = integrate_labels(labels, lengths, per_contig, minfasta)
    extracted = [[] for _ in range(n_ext)]
    for name, l in zip(names, lab):
        if l >= 0:
            extracted[l].append(name)
    return extracted


def _network_input(table, is_combined):
    """What the embedding network is fed (long_read_cluster.py:57-65): the 136 k-mer columns, or in combined
    mode all columns, column- then row-normalised when the reference's norm_abundance() asks for it."""
    if not is_combined:
        return table[:, :136]
    from SemiBin.utils import norm_abundance
    if not norm_abundance(table):
        return table
    scaled = table / table.sum(axis=0)
    return scaled / np.abs(scaled).sum(axis=1, keepdims=True)        # == sklearn normalize(axis=1, norm='l1')


def _markers_per_contig(names, contig_dict, out, binned_length, args):
    """Marker genes of every contig through the reference's own tooling (ORF finder + hmmsearch,
    long_read_cluster.py:83-99); returns its contig -> [marker names] mapping."""
    import os
    import tempfile
    from SemiBin.markers import estimate_seeds, get_marker
    with tempfile.TemporaryDirectory() as scratch:
        fasta = os.path.join(scratch, 'concatenated.fna')
        with open(fasta, 'wt') as fh:
            fh.writelines(f'>{name}\n{contig_dict[name]}\n' for name in names)
        estimate_seeds(fasta, binned_length, args.num_process, output=out, orf_finder=args.orf_finder,
                       prodigal_output_faa=args.prodigal_output_faa)
        found = get_marker(os.path.join(out, 'markers.hmmout'), orf_finder=args.orf_finder,
                           min_contig_len=binned_length, fasta_path=fasta, contig_to_marker=True)
    return found if len(found) else {}


def cluster_long_read(logger, model, data, device, is_combined, n_sample, out, contig_dict, *,
                      binned_length, args, minfasta):
    """Drop-in for SemiBin.long_read_cluster.cluster_long_read (:51-161): same arguments, same files written
    (`output_bins/`, `bins_info.tsv`, `contig_bins.tsv`).  The kNN graph (:108-113), the 12 DBSCAN runs
    (:115-120) and the bin extraction loop (:122-150) run on the GPU; the embedding forward is the reference
    model's, marker search and write_bins are the reference package's own functions."""
    import os
    import shutil
    import pandas as pd
    import torch
    from SemiBin.utils import write_bins
    names = list(data.index)
    table = data.values
    with torch.no_grad():                                            # :67-70
        model.eval()
        net_in = torch.from_numpy(np.ascontiguousarray(_network_input(table, is_combined))).to(device)
        embedding = model.embedding(net_in.float()).detach()         # stays on `device` (one batch, as the reference: a
        #                                                              chunked forward could change GEMM summation order)
    lengths = np.fromiter((len(contig_dict[name]) for name in names), dtype=np.int64, count=len(names))
    depth = None if is_combined else table[:, 136:].astype(np.float32)
    contig2marker = _markers_per_contig(names, contig_dict, out, binned_length, args)
    bin_dir = os.path.join(out, 'output_bins')                       # :101-105
    if os.path.exists(bin_dir):
        logger.warning(f'Previous output directory `{bin_dir}` found. Over-writing it.')
        shutil.rmtree(bin_dir)
    os.makedirs(bin_dir, exist_ok=True)
    logger.debug('Running DBSCAN.')
    label_stack = long_read_labels(embedding, depth, lengths, n_sample=n_sample, is_combined=is_combined)
    logger.debug('Integrating results.')
    extracted = integrate_bins(label_stack, contig2marker, names, contig_dict, minfasta)
    bin_of = {name: b for b, members in enumerate(extracted) for name in members}     # :145-150
    contig_labels = [bin_of.get(name, -1) for name in names]
    written = write_bins(names, contig_labels, bin_dir, contig_dict, minfasta=minfasta, output_tag=args.output_tag,
                         output_compression=args.output_compression)
    logger.info(f'Number of bins: {len(written)}')
    written.to_csv(os.path.join(out, 'bins_info.tsv'), index=False, sep='\t')
    pd.DataFrame({'contig': names, 'bin': contig_labels}).to_csv(os.path.join(out, 'contig_bins.tsv'), index=False,
                                                                 sep='\t')
