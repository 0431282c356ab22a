"""TEST ORACLE -- the long-read ensemble integration step.

Reference: SemiBin/long_read_cluster.py
  :12-49    get_best_bin(results, contig_to_marker, namelist, contig_dict, minfasta)
  :122-143  the extraction loop of cluster_long_read (inline, not a function)
  :145-150  contig2ix / contig_labels

`reference_loop` runs the reference's OWN get_best_bin inside a transcription of
the inline loop (build container only: imports /root/reference); it generates the
golden vectors.  `integrate_restated` is the numpy restatement the GPU path is
checked against at sizes the reference loop cannot reach (it is O(iterations * R * N)
interpreter work).
"""
import numpy as np

CONTAMINATION_STEPS = (0.1, 0.2, 0.3, 0.4, 0.5, 1)        # long_read_cluster.py:16


def reference_loop(dbscan_results, contig2marker, contig_list, contig_dict, minfasta):
    """long_read_cluster.py:122-150 with the reference's get_best_bin.  Returns
    (extracted: list of lists of names, contig_labels: list of int)."""
    from oracle.reference_calls import _import_reference
    _import_reference()
    from SemiBin.long_read_cluster import get_best_bin
    namelist = list(contig_list)
    contig_list = list(contig_list)
    dbscan_results = [list(r) for r in dbscan_results]
    extracted = []
    while sum(len(contig_dict[contig]) for contig in contig_list) >= minfasta:      # :123
        if len(contig_list) == 1:                                                   # :124-126
            extracted.append(contig_list)
            break
        max_bin = get_best_bin(dbscan_results, contig2marker, contig_list, contig_dict, minfasta)   # :128-132
        if not max_bin:                                                             # :133-134
            break
        extracted.append(max_bin)                                                   # :136
        for clustered in max_bin:                                                   # :137-141
            ix = contig_list.index(clustered)
            contig_list.pop(ix)
            for r in dbscan_results:
                r.pop(ix)
    contig2ix = {}                                                                  # :145-148
    for i, cs in enumerate(extracted):
        for c in cs:
            contig2ix[c] = i
    return extracted, [contig2ix.get(c, -1) for c in namelist]                      # :150


def _f1(distinct, total, n_markers):
    """long_read_cluster.py:35-40, Python float arithmetic in the reference's order."""
    recall = distinct / n_markers
    contamination = (total - distinct) / total
    f1 = 2 * recall * (1 - contamination) / (recall + (1 - contamination))
    return f1, contamination


def integrate_restated(labels, lengths, markers, minfasta, n_markers=107):
    """labels: int array [R, N] (-1 = noise); lengths: int [N]; markers: list of N
    int arrays (marker ids, unique per contig).  Returns (contig_labels int64 [N],
    n_extracted).  Same result as reference_loop, per-iteration work vectorised."""
    labels = np.asarray(labels)
    R, N = labels.shape
    lengths = np.asarray(lengths, dtype=np.int64)
    nm = max(int(max((m.max() for m in markers if len(m)), default=-1)) + 1, 1)
    M = np.zeros((N, nm), dtype=np.int32)
    for i, m in enumerate(markers):
        M[i, np.asarray(m, dtype=np.int64)] = 1
    alive = np.ones(N, dtype=bool)
    out = np.full(N, -1, dtype=np.int64)
    n_ext = 0
    idx = np.arange(N)
    while lengths[alive].sum() >= minfasta:
        if alive.sum() == 1:
            out[alive] = n_ext
            n_ext += 1
            break
        best = None          # (class, f1, -weight, r, first) maximised with class minimised
        for r in range(R):
            lab = labels[r]
            ok = alive & (lab >= 0)
            if not ok.any():
                continue
            nl = int(lab[ok].max()) + 1
            w = np.bincount(lab[ok], weights=lengths[ok].astype(np.float64), minlength=nl).astype(np.int64)
            cnt = np.zeros((nl, nm), dtype=np.int64)
            np.add.at(cnt, lab[ok], M[ok])
            total = cnt.sum(axis=1)
            distinct = (cnt > 0).sum(axis=1)
            first = np.full(nl, N, dtype=np.int64)
            np.minimum.at(first, lab[ok], idx[ok])
            for b in np.flatnonzero((w >= minfasta) & (total > 0)):
                f1, cont = _f1(int(distinct[b]), int(total[b]), n_markers)
                cls = next(i for i, t in enumerate(CONTAMINATION_STEPS) if cont <= t)
                key = (-cls, f1, -int(w[b]), r, int(first[b]))
                if best is None or key > best[0]:
                    best = (key, r, int(b))
        if best is None:
            break
        _, r, b = best
        sel = alive & (labels[r] == b)
        out[sel] = n_ext
        n_ext += 1
        alive &= ~sel
    return out, n_ext


def random_case(seed, n=120, R=4, nm=12):
    """Adversarial small input: many equal lengths and marker sets, one result
    duplicated (ties across results), noise labels.  Used by make_golden and the
    tests (regenerated from the seed; only the reference's OUTPUT is stored)."""
    rng = np.random.default_rng(seed)
    labels = np.full((R, n), -1, dtype=np.int64)
    for r in range(R):
        lab = rng.integers(0, rng.integers(2, n // 3), n)
        lab[rng.random(n) < 0.15] = -1
        u, inv = np.unique(lab[lab >= 0], return_inverse=True)
        lab[lab >= 0] = inv
        labels[r] = lab
    if R > 2:
        labels[2] = labels[0]
    lengths = rng.integers(1, 6, n) * 1000
    markers = [np.unique(rng.integers(0, nm, rng.integers(0, 3))).astype(np.int32) for _ in range(n)]
    return labels, lengths, markers, (1000, 3000, 8000)[seed % 3]


def synth_case(n, seed, k=100):
    """DBSCAN label stack of a synthetic long-read set + markers (CPU: oracle kNN)."""
    from semibin_b200 import synth
    from oracle import dbscan as od, knn as ok
    X, L = synth.long_read_case(n, n_sample=1, seed=seed)
    _, genome = synth.embeddings(n, seed, return_genome=True)
    markers = synth.marker_table(genome, L, seed=seed)
    rd, ri = ok.knn_restated(X, min(k, n - 1))
    labels = np.stack([od.dbscan_minprop(rd, ri, e, 5, L)[1] for e in od.EPS_GRID])
    return labels, L, markers
