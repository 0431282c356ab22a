"""TEST INFRASTRUCTURE (see oracle/__init__.py) -- CPU restatement of the seeded, weighted KMeans the reference
runs per bin when it re-clusters (SemiBin/cluster.py:171-255, row N4 of SURVEY.md section 8):

    KMeans(n_clusters=num_bin, init=seeds_embedding, n_init=1, random_state=seed).fit(X_bin, sample_weight=length)

The algorithm lives in scikit-learn 1.9.0 (absent from /root/reference; pinned by pixi.lock), `algorithm='lloyd'`:
  sklearn/cluster/_kmeans.py:1436-1560     fit: X -= X.mean(axis=0) (and the init array), tol = mean(var(X)) * 1e-4
  sklearn/cluster/_kmeans.py:_kmeans_single_lloyd   the loop, strict convergence (labels unchanged) / centre shift <= tol,
                                            one more E-step when it stopped on the shift
  sklearn/cluster/_k_means_lloyd.pyx        E-step: argmin_j ||c_j||^2 - 2 <x, c_j> (first minimum), M-step: weighted sums
  sklearn/cluster/_k_means_common.pyx:167-211  _relocate_empty_clusters_dense
  sklearn/cluster/_k_means_common.pyx:274-295  _average_centers (an empty cluster is put on the heaviest one)
With an array `init` and n_init=1 nothing is random: `random_state` is unused.

Pinned against live scikit-learn: tests/golden/kmeans_cases.npz (oracle/make_golden.py::kmeans_cases) holds sklearn's
labels, centres and iteration counts; tests/test_oracle.py compares.  scikit-learn accumulates in float32 in an order
that depends on its chunking and thread count, this restatement (and the CUDA kernel) in float64 in index order, so the
contract is: identical labels except for points whose two nearest final centres are closer than `TIE_RTOL` in squared
distance (compare_labels below)."""
import numpy as np

TIE_RTOL = 1e-4


def lloyd_seeded(X, sample_weight, init, max_iter=300, tol=1e-4):
    """Returns (labels int32 [n], centers float64 [K, d] in the caller's frame, n_iter)."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    init = np.ascontiguousarray(init, dtype=np.float32)
    w = np.asarray(sample_weight, dtype=np.float32).astype(np.float64)
    n, d = X.shape
    K = init.shape[0]
    # fit(): centre the data and the seeds with the float32 column mean (_kmeans.py: X_mean = X.mean(axis=0); X -= X_mean)
    mean = X.astype(np.float64).mean(axis=0).astype(np.float32)
    Xc = (X - mean).astype(np.float64)
    cen = (init - mean).astype(np.float64)
    tol_abs = float(np.mean(np.var(Xc, axis=0)) * tol)
    labels = np.full(n, -1, np.int32)
    labels_old = labels.copy()
    strict = False
    n_iter = 0
    for it in range(max_iter):
        n_iter = it + 1
        # E-step (first minimum wins ties, as the strict '<' scan of _k_means_lloyd.pyx)
        d2 = ((Xc[:, None, :] - cen[None, :, :]) ** 2).sum(axis=2)
        labels = d2.argmin(axis=1).astype(np.int32)
        # M-step
        new = np.zeros_like(cen)
        wsum = np.zeros(K)
        np.add.at(new, labels, Xc * w[:, None])
        np.add.at(wsum, labels, w)
        _relocate_empty(Xc, w, cen, new, wsum, labels)
        _average(new, wsum)
        shift_tot = float(((new - cen) ** 2).sum())
        cen = new
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if shift_tot <= tol_abs:
            break
        labels_old = labels.copy()
    if not strict:
        d2 = ((Xc[:, None, :] - cen[None, :, :]) ** 2).sum(axis=2)
        labels = d2.argmin(axis=1).astype(np.int32)
    return labels, cen + mean.astype(np.float64), n_iter


def _relocate_empty(Xc, w, cen_old, new, wsum, labels):
    """_k_means_common.pyx:167-211."""
    empty = np.where(wsum == 0)[0]
    if len(empty) == 0:
        return
    dist = ((Xc - cen_old[labels]) ** 2).sum(axis=1)
    if dist.max() == 0:
        return
    # the n_empty farthest points, farthest first (argpartition + reverse in the reference; ties by lower index here)
    far = np.lexsort((np.arange(len(dist)), -dist))[:len(empty)]
    for e, f in zip(empty, far):
        old = labels[f]
        new[old] -= Xc[f] * w[f]
        new[e] = Xc[f] * w[f]
        wsum[e] = w[f]
        wsum[old] -= w[f]


def _average(new, wsum):
    """_k_means_common.pyx:274-295 (sequential: an empty cluster copies the heaviest one as it is at that moment)."""
    amax = int(np.argmax(wsum))
    for j in range(len(wsum)):
        if wsum[j] > 0:
            new[j] *= 1.0 / wsum[j]
        else:
            new[j] = new[amax]


def compare_labels(X, centers, ref_labels, got_labels, rtol=TIE_RTOL):
    """Labels must agree except where the point's two nearest centres (of `centers`, the reference's final centres) are
    within rtol of each other in squared distance.  Returns dict(n, n_differ, n_unexplained)."""
    X = np.asarray(X, np.float64); C = np.asarray(centers, np.float64)
    ref = np.asarray(ref_labels); got = np.asarray(got_labels)
    diff = np.where(ref != got)[0]
    bad = 0
    for i in diff:
        d2 = ((X[i] - C) ** 2).sum(axis=1)
        a, b = d2[ref[i]], d2[got[i]]
        if abs(a - b) > rtol * max(a, b, 1e-300):
            bad += 1
    return dict(n=len(ref), n_differ=int(len(diff)), n_unexplained=int(bad))


def recluster_labels(embedding_new, contig_labels, seeds_by_bin, lengths, total_size, minfasta, kmeans=lloyd_seeded):
    """The label bookkeeping of SemiBin/cluster.py:221-253 around the per-bin KMeans: bins with more than one seed and at
    least minfasta bases are split into len(seed) bins, every other bin keeps one label; labels are handed out in bin
    order.  seeds_by_bin[bin] = list of row indices of the seed contigs."""
    contig_labels = np.asarray(contig_labels)
    out = np.full(len(contig_labels), -1, dtype=contig_labels.dtype)
    next_label = 0
    for b in range(int(contig_labels.max()) + 1):
        seed = seeds_by_bin.get(b, [])
        if len(seed) > 1 and total_size.get(b, 0) >= minfasta:
            idx = np.where(contig_labels == b)[0]
            lab, _, _ = kmeans(embedding_new[idx], lengths[idx], embedding_new[np.asarray(seed)])
            out[idx] = next_label + lab
            next_label += len(seed)
        else:
            out[contig_labels == b] = next_label
            next_label += 1
    return out
