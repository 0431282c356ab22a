"""Re-clustering drop-in: SemiBin/cluster.py:171-255 (`recluster_bins`), row N4 of SURVEY.md section 8.

The reference walks its bins on the host and runs one scikit-learn KMeans per bin that has more than one seed
(`KMeans(n_clusters=num_bin, init=seeds_embedding, n_init=1, random_state=seed).fit(re_bin_features,
sample_weight=length_weight)`, :236-241).  Here every such bin becomes one problem of ONE batched launch
(`sbk_kmeans_batch`: one CTA per bin runs the whole seeded, weighted Lloyd loop on the device); the label bookkeeping
around it (:221-253) is the reference's.

    kmeans_batch(X, weights, point_offsets, seed_rows, cluster_offsets)   the array-level call
    recluster_labels(embedding_new, contig_labels, seeds_by_bin, lengths, total_size, minfasta)
    recluster_bins(logger, data, *, n_sample, embedding, is_combined, contig_labels, minfasta, contig_dict,
                   binned_length, orf_finder, num_process, random_seed)   same signature as the reference

Seed estimation (`estimate_seeds`: ORF finder + hmmsearch on the concatenated bins, :203-214) is marker tooling and
stays the reference's: `recluster_bins` imports it from the installed SemiBin (or takes `estimate_seeds=`).
There is no CPU fallback: without the library or a GPU the call raises.
"""
import ctypes as C
import math
import os
import tempfile
from collections import defaultdict

import numpy as np

from . import _lib
from .neighbors import _workspace

KMEANS_MAX_ITER = 300      # sklearn.cluster.KMeans defaults, which the reference leaves untouched
KMEANS_TOL = 1e-4


def kmeans_batch(X, weights, point_offsets, centers_init, cluster_offsets, *, max_iter=KMEANS_MAX_ITER, tol=KMEANS_TOL,
                 return_centers=False):
    """Seeded, weighted Lloyd k-means of many independent problems in one launch.
    X: CUDA float32 [n_total, d], the rows of problem p contiguous at [point_offsets[p], point_offsets[p+1]);
    weights: CUDA float32 [n_total]; centers_init: CUDA float32 [cluster_offsets[-1], d] (the seed rows);
    point_offsets (int64) / cluster_offsets (int32): host arrays of length n_problems + 1.
    Returns (labels int32 CUDA [n_total] -- cluster index within the problem, n_iter int32 CUDA [n_problems])
    and, with return_centers, the final centres (float64 CUDA)."""
    import torch
    _lib.require_cuda()
    lib = _lib.load()
    if X.dim() != 2 or X.dtype != torch.float32 or not X.is_cuda:
        raise ValueError('X must be a 2-D CUDA float32 tensor (the reference re-clusters float32 features)')
    X = X.contiguous()
    n_total, d = X.shape
    po = np.ascontiguousarray(point_offsets, dtype=np.int64)
    co = np.ascontiguousarray(cluster_offsets, dtype=np.int32)
    n_prob = len(po) - 1
    if len(co) != n_prob + 1 or n_prob < 0:
        raise ValueError('point_offsets and cluster_offsets must both have n_problems + 1 entries')
    weights = weights.to(device=X.device, dtype=torch.float32).contiguous()
    centers_init = centers_init.to(device=X.device, dtype=torch.float32).contiguous()
    k_total = int(co[-1]) if n_prob else 0
    if weights.numel() != n_total or tuple(centers_init.shape) != (k_total, d):
        raise ValueError('weights / centers_init do not match X and the offsets')
    labels = torch.full((n_total,), -1, dtype=torch.int32, device=X.device)
    n_iter = torch.zeros((max(n_prob, 1),), dtype=torch.int32, device=X.device)
    centers = torch.empty((max(k_total, 1), d), dtype=torch.float64, device=X.device) if return_centers else None
    if n_prob:
        with torch.cuda.device(X.device):
            ws = _workspace(lib.sbk_kmeans_batch_workspace_bytes(n_total, d, n_prob, k_total), X.device)
            rc = lib.sbk_kmeans_batch(_lib.ptr(X), X.stride(0), d, _lib.ptr(weights), n_total,
                                      po.ctypes.data_as(C.c_void_p), co.ctypes.data_as(C.c_void_p), n_prob,
                                      _lib.ptr(centers_init), int(max_iter), float(tol), _lib.ptr(labels), _lib.ptr(n_iter),
                                      _lib.ptr(centers) if centers is not None else None,
                                      _lib.ptr(ws), ws.numel(), _lib.stream_ptr(X.device))
            _lib.check(rc)                     # SBK_ERR_INVALID -> ValueError with scikit-learn's message
    n_iter = n_iter[:n_prob]
    return (labels, n_iter, centers[:k_total]) if return_centers else (labels, n_iter)


def _kmeans_problems(embedding_new, rows, point_offsets, seed_rows, cluster_offsets, lengths, device=None):
    """The device part of recluster_labels: gather the member and seed rows of every problem on the GPU, one batched
    launch, labels back as numpy int32 [len(rows)]."""
    import torch
    _lib.require_cuda()
    E = embedding_new if isinstance(embedding_new, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(embedding_new))
    if E.dtype != torch.float32:
        E = E.float()
    dev = device if device is not None else (E.device if E.is_cuda else torch.device('cuda', torch.cuda.current_device()))
    E = E.to(dev)
    Xg = E[torch.from_numpy(rows).to(dev)]
    Cg = E[torch.from_numpy(seed_rows).to(dev)]
    w = torch.from_numpy(np.asarray(lengths)[rows].astype(np.float32)).to(dev)
    lab, _ = kmeans_batch(Xg, w, point_offsets, Cg, cluster_offsets)
    return lab.cpu().numpy()


def recluster_labels(embedding_new, contig_labels, seeds_by_bin, lengths, total_size, minfasta, device=None):
    """cluster.py:221-253 with all per-bin KMeans runs in one launch.
    embedding_new: [N, d] float32 (numpy or torch, host or device); contig_labels: int array [N];
    seeds_by_bin: {bin: [row indices of the seed contigs]}; lengths: int array [N] (len(contig_dict[name]));
    total_size: {bin: bases}.  Returns the re-clustered labels (numpy, dtype of contig_labels)."""
    contig_labels = np.asarray(contig_labels)
    n_bins = int(contig_labels.max()) + 1 if len(contig_labels) else 0
    order = np.argsort(contig_labels, kind='stable')              # members of a bin in ascending contig order (:231)
    bounds = np.searchsorted(contig_labels[order], np.arange(n_bins + 1))
    split = [b for b in range(n_bins) if len(seeds_by_bin.get(b, [])) > 1 and total_size.get(b, 0) >= minfasta]
    out = np.full(len(contig_labels), -1, dtype=contig_labels.dtype)
    labels_k = {}
    if split:
        rows = np.concatenate([order[bounds[b]:bounds[b + 1]] for b in split])
        po = np.concatenate([[0], np.cumsum([bounds[b + 1] - bounds[b] for b in split])]).astype(np.int64)
        seed_rows = np.concatenate([np.asarray(seeds_by_bin[b], dtype=np.int64) for b in split])
        co = np.concatenate([[0], np.cumsum([len(seeds_by_bin[b]) for b in split])]).astype(np.int32)
        lab = _kmeans_problems(embedding_new, rows, po, seed_rows, co, lengths, device)
        for t, b in enumerate(split):
            labels_k[b] = lab[po[t]:po[t + 1]]
    next_label = 0
    for b in range(n_bins):
        members = order[bounds[b]:bounds[b + 1]]
        if b in labels_k:
            out[members] = next_label + labels_k[b]
            next_label += len(seeds_by_bin[b])
        else:
            out[members] = next_label
            next_label += 1
    return out


def _recluster_features(data, embedding, n_sample, is_combined):
    """cluster.py:184-195: in single-sample / multi-sample mode the embedding is extended by the per-sample mean depths,
    each divided by its sample's scale (mean depth rounded up to the next hundred) and multiplied by one common integer
    weight that brings them to the magnitude of the embedding; combined mode re-clusters the embedding alone."""
    if is_combined:
        return embedding
    depth = data.values[:, 136:].astype(np.float32)
    mean_depth = depth[:, 0:2 * n_sample:2]                      # (mean, variance) pairs per sample: the means
    mean_depth = mean_depth / (np.ceil(mean_depth.mean(axis=0) / 100) * 100)
    ratio = np.mean(np.abs(embedding)) / np.mean(mean_depth)
    weight = 20 * math.ceil(ratio / 10)
    return np.concatenate((embedding, mean_depth * weight), axis=1)


def recluster_bins(logger, data, *, n_sample, embedding, is_combined: bool, contig_labels, minfasta, contig_dict,
                   binned_length, orf_finder, num_process, random_seed, estimate_seeds=None):
    """Same signature and results as SemiBin.cluster.recluster_bins (cluster.py:171-255); `random_seed` is accepted and,
    as in the reference (array init, n_init=1), changes nothing."""
    if estimate_seeds is None:
        from SemiBin.markers import estimate_seeds
    features = _recluster_features(data, np.asarray(embedding), n_sample, is_combined)
    contig_labels = np.asarray(contig_labels)
    names = list(data.index)
    lengths = np.fromiter((len(contig_dict[h]) for h in names), dtype=np.int64, count=len(names))
    bases = np.bincount(contig_labels, weights=lengths, minlength=int(contig_labels.max()) + 1).astype(np.int64)
    total_size = defaultdict(int, {b: int(v) for b, v in enumerate(bases)})
    big_enough = bases[contig_labels] >= minfasta               # :197-205: only bins of at least minfasta bases are searched
    with tempfile.TemporaryDirectory() as tdir:
        cfasta = os.path.join(tdir, 'concatenated.fna')
        with open(cfasta, 'wt') as fh:
            fh.writelines(f'>bin{contig_labels[ix]:06}.{h}\n{contig_dict[h]}\n' for ix, h in enumerate(names) if big_enough[ix])
        # the ORF finder cannot be bypassed: the contigs were renamed (:213)
        seeds = estimate_seeds(cfasta, binned_length, num_process, multi_mode=True, orf_finder=orf_finder)
    if not seeds:
        logger.warning('No bins found in the concatenated fasta file.')
        return contig_labels
    name2ix = {name: ix for ix, name in enumerate(data.index)}
    seeds_by_bin = {}
    for b in range(int(contig_labels.max()) + 1):
        seed = seeds.get(f'bin{b:06}', [])
        if len(seed) > 1:
            seeds_by_bin[b] = [name2ix[s] for s in seed]
    out = recluster_labels(features, contig_labels, seeds_by_bin, lengths, total_size, minfasta)
    assert out.min() >= 0
    return out
