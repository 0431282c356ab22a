"""Seeded synthetic inputs shaped like SemiBin's.

Used by bench.py, the smoke test and the parity tests so that the CUDA path,
the oracle and the CPU baseline all see the same arrays.  numpy only.
"""
import numpy as np

D_EMB = 100      # SemiBin/semi_supervised_model.py:27,75
D_KMER = 136     # SemiBin/generate_kmer.py:7-18


def _genome_assignment(rng, n, mean_size=200):
    n_genomes = max(2, n // mean_size)
    sizes = rng.lognormal(mean=np.log(mean_size), sigma=1.0, size=n_genomes)
    sizes = np.maximum(1, np.floor(sizes / sizes.sum() * n)).astype(np.int64)
    sizes[-1] += n - sizes.sum()
    if sizes[-1] < 1:                      # keep every genome non-empty
        deficit = 1 - sizes[-1]
        sizes[-1] = 1
        sizes[np.argmax(sizes)] -= deficit
    genome = np.repeat(np.arange(n_genomes), sizes)
    rng.shuffle(genome)
    return genome, n_genomes


def embeddings(n, seed=20240, d=D_EMB, sigma=0.02, dtype=np.float32, return_genome=False, mean_size=200):
    """Gaussian mixture mimicking the contrastive embedding: genome centres
    ~ N(0, 0.16^2 I), within-genome noise sigma per dimension (intra-genome
    distance ~0.28 < 1 <= inter-genome ~2.3)."""
    rng = np.random.default_rng(seed)
    genome, g = _genome_assignment(rng, n, mean_size)
    centres = rng.normal(0.0, 0.16, size=(g, d)).astype(np.float32)
    x = centres[genome]
    x += rng.standard_normal(size=(n, d), dtype=np.float32) * np.float32(sigma)
    x = np.ascontiguousarray(x.astype(dtype))
    return (x, genome) if return_genome else x


def contig_lengths(n, seed=20240):
    rng = np.random.default_rng(seed + 7)
    L = rng.lognormal(mean=np.log(5000.0), sigma=1.0, size=n)
    return np.clip(L, 1000, 500000).astype(np.int64)


def kmer_features(n, seed=20240, genome=None, lengths=None):
    """Per-genome Dirichlet(50) 4-mer profile, per-contig counts ~ Poisson(L*p)
    (multinomial up to the total), +1e-5 and row-normalised as in
    SemiBin/generate_kmer.py:46-47.  float64, rows sum to 1."""
    rng = np.random.default_rng(seed + 1)
    if genome is None:
        genome, g = _genome_assignment(np.random.default_rng(seed), n)
    g = int(genome.max()) + 1
    if lengths is None:
        lengths = contig_lengths(n, seed)
    prof = rng.dirichlet(np.full(D_KMER, 50.0), size=g)
    lam = prof[genome] * (lengths[:, None] - 3)
    cnt = rng.poisson(lam).astype(np.float64)
    cnt += 1e-5
    cnt /= cnt.sum(axis=1, keepdims=True)
    return np.ascontiguousarray(cnt)


def depth_table(n, n_sample, seed=20240, genome=None):
    """float32 (n, 2*n_sample) interleaved [mean_s, var_s]
    (SemiBin/generate_coverage.py:76-79 layout)."""
    rng = np.random.default_rng(seed + 2)
    if genome is None:
        genome, g = _genome_assignment(np.random.default_rng(seed), n)
    g = int(genome.max()) + 1
    a = rng.lognormal(np.log(10.0), 1.5, size=(g, n_sample))
    mean = a[genome] * (1 + 0.1 * rng.standard_normal((n, n_sample))) + 1e-5
    mean = np.abs(mean)
    var = mean * (1 + np.abs(rng.standard_normal((n, n_sample)))) * 3
    out = np.empty((n, 2 * n_sample), np.float32)
    out[:, 0::2] = mean
    out[:, 1::2] = var
    return out


def short_read_case(n, n_sample=1, seed=20240):
    """(embedding f32 N x 100, features f64 N x 136, depth f32 N x 2S, lengths)"""
    emb, genome = embeddings(n, seed, return_genome=True)
    L = contig_lengths(n, seed)
    return emb, kmer_features(n, seed, genome, L), depth_table(n, n_sample, seed, genome), L


def long_read_case(n, n_sample=1, seed=20240):
    """(embedding_new f32 N x (100+S), lengths) as long_read_cluster.py:75-81."""
    emb, genome = embeddings(n, seed, return_genome=True)
    dep = depth_table(n, n_sample, seed, genome)[:, 0::2]
    x = np.concatenate((emb, np.log(np.clip(dep, 1e-6, None))), axis=1).astype(np.float32)
    return np.ascontiguousarray(x), contig_lengths(n, seed)


def tie_grid(n, d=24, seed=3, dtype=np.float32, n_dup=0):
    """Integer-grid points: many EXACT distance ties (all arithmetic exact in
    float32/float64), optionally with n_dup exact duplicate rows appended."""
    rng = np.random.default_rng(seed)
    x = rng.integers(-1, 2, size=(n, d)).astype(dtype)
    if n_dup:
        x[-n_dup:] = x[0]
    return np.ascontiguousarray(x)


def marker_table(genome, lengths, seed=20240, n_markers=107, completeness=(0.4, 1.0), dup_rate=0.03):
    """Single-copy marker genes per contig for the long-read ensemble step
    (SemiBin/long_read_cluster.py:12-49 consumes contig -> [marker names]).  Every
    genome carries a random subset of the `n_markers` markers (its completeness),
    each placed on one of its contigs with probability proportional to length; a
    few markers are duplicated on a second contig (contamination).  Returns a list
    of sorted int arrays, one per contig (each marker at most once per contig, as
    SemiBin/markers.py:63 guarantees)."""
    rng = np.random.default_rng(seed + 11)
    n = len(genome)
    per_contig = [[] for _ in range(n)]
    order = np.argsort(genome, kind='stable')
    bounds = np.flatnonzero(np.diff(genome[order])) + 1
    for members in np.split(order, bounds):
        frac = rng.uniform(*completeness)
        present = np.flatnonzero(rng.random(n_markers) < frac)
        w = lengths[members].astype(np.float64)
        w /= w.sum()
        home = rng.choice(members, size=len(present), p=w)
        for m, c in zip(present, home):
            per_contig[c].append(int(m))
        dup = present[rng.random(len(present)) < dup_rate]
        for m in dup:
            c = int(rng.choice(members, p=w))
            if int(m) not in per_contig[c]:
                per_contig[c].append(int(m))
    return [np.array(sorted(v), dtype=np.int32) for v in per_contig]


def recluster_problem(seed, n, d, K, spread, sep, dup_seeds=0):
    """One re-clustering problem in the shape of SemiBin/cluster.py:229-243: float32 features (embedding + scaled depth
    columns), K seed contigs picked from the data, integer contig lengths as weights."""
    rng = np.random.default_rng(seed)
    centres = rng.normal(size=(K, d)) * sep
    sizes = rng.multinomial(n - K, rng.dirichlet(np.ones(K) * 2.0)) + 1
    lab = np.repeat(np.arange(K), sizes)
    X = (centres[lab] + rng.normal(size=(n, d)) * spread).astype(np.float32)
    perm = rng.permutation(n)
    X, lab = X[perm], lab[perm]
    seeds = np.array([np.where(lab == j)[0][0] for j in range(K)])
    if dup_seeds:                              # several seeds from the same true cluster (over-estimated bin count)
        extra = np.where(lab == 0)[0][1:1 + dup_seeds]
        seeds = np.concatenate([seeds, extra])
    w = np.exp(rng.normal(8.5, 1.0, size=n)).astype(np.int64) + 1000
    return X, w, seeds
