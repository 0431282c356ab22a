#!/usr/bin/env python
"""bench_path.py -- stage-by-stage timings of the remaining SURVEY section 8 rows
(not the driver's benchmark; that is bench.py).  Prints one JSON line per config:

  C2  short-read, N=200k, S=1  : both kNN graphs + edge list (KL)
  C3  multi-sample, N=1M, S=10 : (ii) KL-weighted edges; (i) combined D=146 with --combined
  C4  long-read, N=500k, D=100+S: kNN graph + 12-eps sweep + DBSCAN labels

Achieved GB/s figures use the ALGORITHMIC bytes of SURVEY section 8(d) / DESIGN.md section 4.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def timed(fn, reps=1):
    import torch
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / reps


def short_read(n, S, combined, k=200, max_node=1, cpu_check=False, sklearn_graphs=False):
    import torch
    from semibin_b200 import synth, cluster
    from semibin_b200.neighbors import knn
    emb, feat, depth, L = synth.short_read_case(n, n_sample=S, seed=20243)
    if combined:
        f = np.concatenate([feat, depth[:, 0::2].astype(np.float64) / 100.0], axis=1)
        f = f / np.sum(f, axis=0)
        f = f / np.abs(f).sum(axis=1, keepdims=True)
        feat = f
    E = torch.from_numpy(emb).cuda()
    F = torch.from_numpy(np.ascontiguousarray(feat)).cuda()
    Dp = torch.from_numpy(depth).cuda()
    res = {}
    knn(E, k)  # warm
    (a_d, a_i, st_a), res['knn_emb_ms'] = timed(lambda: knn(E, k, return_stats=True))
    knn(F, k)
    (b_d, b_i, st_b), res['knn_feat_ms'] = timed(lambda: knn(F, k, return_stats=True))
    ths = cluster.threshold_sequence()
    cluster.edges_stage1(a_d, a_i, b_d, b_i, ths)
    (keep, rowmax, counts), res['stage1_ms'] = timed(lambda: cluster.edges_stage1(a_d, a_i, b_d, b_i, ths))
    thr = cluster.pick_threshold(counts.cpu().numpy(), n, max_node)
    cluster.edges_stage2(keep, a_d, a_i, 0, thr, Dp, S, combined)
    (X, Y, V), res['stage2_ms'] = timed(lambda: cluster.edges_stage2(keep, a_d, a_i, 0, thr, Dp, S, combined))
    n_edges = int(X.shape[0])
    kept_dir = n_edges * 2                       # directed edges surviving ~ 2x upper triangle
    res.update(n=n, S=S, combined=combined, k=k, d_feat=int(F.shape[1]), threshold=thr, n_edges=n_edges,
               stats_emb={k2: st_a[k2] for k2 in ('n_fallback_rows', 'n_candidates', 'n_reranked', 'ms_filter', 'ms_rerank', 'ms_sample', 'ms_prepare', 'ms_fallback', 'n_fail_overflow', 'n_fail_few', 'n_fail_survivors', 'n_fail_certificate', 'cand_capacity', 'order_stat', 'sample_size', 'n_split_cols', 'n_retry_rows')},
               stats_feat={k2: st_b[k2] for k2 in ('n_fallback_rows', 'n_candidates', 'n_reranked', 'ms_filter', 'ms_rerank', 'ms_sample', 'ms_prepare', 'ms_fallback', 'kpad', 'n_fail_overflow', 'n_fail_few', 'n_fail_survivors', 'n_fail_certificate', 'cand_capacity', 'order_stat', 'sample_size', 'n_split_cols', 'n_retry_rows')})
    b1 = n * k * (4 + 4 + 4 + 8) + n * ((k + 31) // 32) * 4
    res['stage1_GBps'] = b1 / res['stage1_ms'] / 1e6
    b2 = n * k * 8 + n * ((k + 31) // 32) * 4 + kept_dir * (12 * S if not combined else 0) + n_edges * 16 * 2
    res['stage2_GBps'] = b2 / res['stage2_ms'] / 1e6
    res['edge_list_total_ms'] = res['knn_emb_ms'] + res['knn_feat_ms'] + res['stage1_ms'] + res['stage2_ms']
    res['tflops_feat_filter'] = 2.0 * n * n * F.shape[1] / (st_b['ms_filter'] * 1e-3) / 1e12 if st_b['ms_filter'] else None
    if cpu_check:
        from oracle import edges as oe, compare, knn as ok
        if sklearn_graphs:
            # the reference's OWN graphs: live scikit-learn kneighbors_graph on the host cores for both matrices
            # (cluster.py:94-105), then the sparse restatement of :109-151 -- nothing of the GPU path feeds the checker
            t0 = time.perf_counter()
            ra_d, ra_i = ok.knn_sklearn(emb, k)
            res['cpu_sklearn_emb_graph_s'] = time.perf_counter() - t0
            t0 = time.perf_counter()
            rb_d, rb_i = ok.knn_sklearn(np.ascontiguousarray(feat), k)
            res['cpu_sklearn_feat_graph_s'] = time.perf_counter() - t0
            res['graphs'] = 'live scikit-learn kneighbors_graph (both), %d host cores' % os.cpu_count()
            res['knn_parity_emb'] = {kk: vv for kk, vv in compare.compare_knn(ra_d, ra_i, a_d.cpu().numpy().astype(np.float64), a_i.cpu().numpy().astype(np.int64), rtol=1e-6).items() if kk != 'bad_rows'}
            res['knn_parity_feat'] = {kk: vv for kk, vv in compare.compare_knn(rb_d, rb_i, b_d.cpu().numpy(), b_i.cpu().numpy().astype(np.int64), rtol=1e-9).items() if kk != 'bad_rows'}
        else:
            ra_d, ra_i = a_d.cpu().numpy().astype(np.float64), a_i.cpu().numpy().astype(np.int64)
            rb_d, rb_i = b_d.cpu().numpy(), b_i.cpu().numpy().astype(np.int64)
            res['graphs'] = "the GPU's own kNN graphs (edge stages only are checked)"
        t0 = time.perf_counter()
        RX, RY, RV, t = oe.edge_list_restated(ra_d, ra_i, rb_d, rb_i, depth, max_node=max_node, is_combined=combined, n_sample=S)
        res['cpu_sparse_restatement_s'] = time.perf_counter() - t0
        res['threshold_ref'] = t
        res['parity'] = compare.compare_edges((RX, RY, RV), (X.cpu().numpy(), Y.cpu().numpy(), V.cpu().numpy()), explain=8,
                                              rtol=0.0 if combined else 1e-5,
                                              bound_fn=None if combined else (lambda x, y: oe.kl_weight_bound(depth, S, x, y)))
    return res


def long_read(n, S, k=200, cpu_check=False):
    import torch
    from semibin_b200 import synth, long_read as lr
    from semibin_b200.neighbors import knn
    X, L = synth.long_read_case(n, n_sample=S, seed=20244)
    Xd = torch.from_numpy(X).cuda()
    Ld = torch.from_numpy(L.astype(np.float64)).cuda()
    res = dict(n=n, S=S, k=k, d=int(X.shape[1]))
    knn(Xd, k)
    (dist, idx, st), res['knn_ms'] = timed(lambda: knn(Xd, k, return_stats=True))
    lr.radius_sweep(dist, idx, lr.EPS_GRID, min_samples=5, sample_weight=Ld)
    (plen, core), res['sweep_ms'] = timed(lambda: lr.radius_sweep(dist, idx, lr.EPS_GRID, min_samples=5, sample_weight=Ld))
    lr.dbscan_labels(idx, plen, core)
    labels, res['labels_ms'] = timed(lambda: lr.dbscan_labels(idx, plen, core))
    res['sweep_GBps'] = (n * k * 8 + n * 12 * 5) / res['sweep_ms'] / 1e6
    res['n_clusters'] = [int(labels[e].max().item()) + 1 for e in range(12)]
    res['stats'] = {k2: st[k2] for k2 in ('n_fallback_rows', 'n_candidates', 'n_reranked', 'ms_filter', 'ms_rerank', 'ms_sample', 'ms_fallback', 'kpad', 'n_fail_overflow', 'n_fail_few', 'n_fail_survivors', 'n_fail_certificate', 'cand_capacity', 'order_stat', 'sample_size', 'n_split_cols')}
    res['total_ms'] = res['knn_ms'] + res['sweep_ms'] + res['labels_ms']
    # N2: ensemble integration (long_read_cluster.py:12-49,122-150) on the 12 label vectors
    _, genome = synth.embeddings(n, 20244, return_genome=True)
    markers = lr.marker_masks(synth.marker_table(genome, L, seed=20244), n)
    minfasta = 200000
    lr.integrate_labels(labels, L, markers, minfasta)
    t0 = time.perf_counter()
    lab, n_bins = lr.integrate_labels(labels, L, markers, minfasta)
    res['integrate_ms'] = (time.perf_counter() - t0) * 1e3
    res['integrate_bins'] = n_bins
    res['integrate_binned_contigs'] = int((lab >= 0).sum())
    if cpu_check:
        from oracle import dbscan as od
        nb = min(n, 50000)
        d = dist[:nb].cpu().numpy().astype(np.float64)
        # CPU baseline: the reference's dbscan call on a 50k-row sub-graph is not meaningful (indices
        # point outside); instead time sklearn on a fresh 50k problem
        from semibin_b200 import synth as sy
        Xs, Ls = sy.long_read_case(nb, n_sample=S, seed=20245)
        from sklearn.neighbors import kneighbors_graph
        from sklearn.cluster import dbscan
        import warnings
        t0 = time.perf_counter()
        G = kneighbors_graph(Xs, k, mode='distance', p=2)
        res['cpu_knn_50k_s'] = time.perf_counter() - t0
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            for eps in (0.05, 0.3, 0.55):
                dbscan(G, eps=eps, min_samples=5, metric='precomputed', sample_weight=Ls)
        res['cpu_dbscan_50k_s_per_eps'] = (time.perf_counter() - t0) / 3
        # the reference's integration loop is O(bins * R * N) interpreter work: time the restatement (faster than
        # the reference's own loop) on a 20k-contig problem and check the GPU result against it
        from oracle import integrate as oi
        Xs, Ls = sy.long_read_case(20000, n_sample=S, seed=20246)
        _, gs = sy.embeddings(20000, 20246, return_genome=True)
        ms = sy.marker_table(gs, Ls, seed=20246)
        labs = np.stack(lr.long_read_labels(Xs, None, Ls, n_sample=S, is_combined=True))
        t0 = time.perf_counter()
        ref, nref = oi.integrate_restated(labs, Ls, ms, 200000)
        res['cpu_integrate_20k_s'] = time.perf_counter() - t0
        t0 = time.perf_counter()
        got, ngot = lr.integrate_labels(labs, Ls, ms, 200000)
        res['gpu_integrate_20k_ms'] = (time.perf_counter() - t0) * 1e3
        res['integrate_parity_20k'] = bool(nref == ngot and (ref == got).all())
    return res


def scaling_case(n, k=200, check_rows=512):
    """C5: kNN graph of N=4M embeddings (several query chunks per call), sampled-row parity against
    live scikit-learn on the host."""
    import torch
    from semibin_b200 import synth
    from semibin_b200.neighbors import knn
    X = synth.embeddings(n, seed=20245)
    Xd = torch.from_numpy(X).cuda()
    res = dict(n=n, k=k, d=int(X.shape[1]))
    (dist, idx, st), res['knn_first_ms'] = timed(lambda: knn(Xd, k, return_stats=True))
    (dist, idx, st), res['knn_ms'] = timed(lambda: knn(Xd, k, return_stats=True))
    res['contigs_per_s'] = n / (res['knn_ms'] * 1e-3)
    res['filter_tflops_true_d'] = 2.0 * n * n * X.shape[1] / (st['ms_filter'] * 1e-3) / 1e12
    res['stats'] = {k2: st[k2] for k2 in ('n_chunks', 'n_filter_launches', 'n_fallback_rows', 'n_candidates', 'n_reranked', 'ms_prepare',
                                          'ms_filter', 'ms_rerank', 'ms_sample', 'ms_fallback', 'kpad', 'cand_capacity',
                                          'order_stat', 'sample_size')}
    if check_rows:
        from oracle import knn as ok, compare
        rng = np.random.default_rng(5)
        lo = int(rng.integers(0, n - check_rows))
        t0 = time.perf_counter()
        rd, ri = ok.knn_sklearn(X, k, q_begin=lo, q_end=lo + check_rows)
        res['cpu_rows_per_s'] = check_rows / (time.perf_counter() - t0)
        r = compare.compare_knn(rd, ri, dist[lo:lo + check_rows].cpu().numpy().astype(np.float64),
                                idx[lo:lo + check_rows].cpu().numpy().astype(np.int64), rtol=1e-5)
        res['parity'] = {kk: r[kk] for kk in ('n_rows_bad', 'n_rows_tied')}
    return res


def short_read_sharded(n, S, combined, k=200, max_node=1):
    """C3 across the GPUs of one box (launch under torchrun): row-block kNN graphs, intersection and weights per
    rank, one all-reduce (threshold counts) and one gather (edge blocks).  Rank 0 prints the line; the edge list
    is checksummed so that runs at different GPU counts can be compared."""
    import hashlib
    import torch
    import torch.distributed as dist
    from semibin_b200 import synth, dist as sdist
    rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0))))
    emb, feat, depth, L = synth.short_read_case(n, n_sample=S, seed=20243)
    if combined:
        f = np.concatenate([feat, depth[:, 0::2].astype(np.float64) / 100.0], axis=1)
        f = f / np.sum(f, axis=0)
        feat = f / np.abs(f).sum(axis=1, keepdims=True)
    E = torch.from_numpy(emb).cuda()
    F = torch.from_numpy(np.ascontiguousarray(feat)).cuda()
    Dp = torch.from_numpy(depth).cuda()
    kw = dict(max_edges=k, max_node=max_node, is_combined=combined, n_sample=S)
    sdist.build_edge_list_sharded(E, F, Dp, **kw)              # warm-up
    times, times_dev = [], []
    for mode, acc in ((False, times_dev), ('view', times)):
        for _ in range(3):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = sdist.build_edge_list_sharded(E, F, Dp, to_host=mode, **kw)
            torch.cuda.synchronize()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            acc.append(float(t.item()) * 1e3)
    X, Y, V = out
    h = hashlib.sha256()
    for a in (X, Y, V):
        h.update(np.ascontiguousarray(a).tobytes())
    res = dict(n=n, S=S, combined=combined, k=k, n_gpus=world, edge_list_ms=min(times), edge_list_ms_all=times,
               edge_list_on_device_ms=min(times_dev), edge_list_on_device_ms_all=times_dev,
               n_edges=int(len(X)), sha256=h.hexdigest()[:16],
               what='embedding, feature and depth matrices resident on every GPU; timed: both kNN graphs (row block per rank), '
                    'intersect + max_node threshold (all-reduce) + KL weights, NCCL gather of the edge blocks, '
                    '(edge_list_ms only) download of the edge list to every rank through pinned buffers')
    if world > 1:
        dist.destroy_process_group()
    return res if rank == 0 else None


def recluster_case(n_total, cpu_bins=60):
    """Row N4 (SemiBin/cluster.py:229-246): the per-bin seeded / weighted KMeans of the re-clustering step for a sample
    of ~n_total binned contigs (bins of 50..6000 contigs, D = 103, 2..8 seeds each) as ONE batched launch; beside it live
    scikit-learn called bin after bin the way the reference calls it, on the first `cpu_bins` bins, with the label check."""
    import torch
    from semibin_b200.recluster import kmeans_batch
    from semibin_b200 import synth
    from oracle import kmeans as okm
    rng = np.random.default_rng(123)
    probs, tot = [], 0
    while tot < n_total:
        n = int(min(6000, max(50, rng.lognormal(6.0, 0.9)))); K = int(rng.integers(2, 9))
        probs.append(synth.recluster_problem(9000 + len(probs), n, 103, K, float(rng.uniform(0.1, 0.4)), float(rng.uniform(0.04, 0.3))))
        tot += n
    X = torch.from_numpy(np.concatenate([p[0] for p in probs])).cuda()
    w = torch.from_numpy(np.concatenate([p[1] for p in probs]).astype(np.float32)).cuda()
    C0 = torch.from_numpy(np.concatenate([p[0][p[2]] for p in probs])).cuda()
    po = np.concatenate([[0], np.cumsum([len(p[0]) for p in probs])]).astype(np.int64)
    co = np.concatenate([[0], np.cumsum([len(p[2]) for p in probs])]).astype(np.int32)
    kmeans_batch(X, w, po, C0, co)
    (lab, nit), ms = timed(lambda: kmeans_batch(X, w, po, C0, co), reps=3)
    lab = lab.cpu().numpy(); nit = nit.cpu().numpy()
    res = dict(n_contigs=int(tot), n_bins=len(probs), d=103, kmeans_batch_ms=ms, lloyd_iterations_total=int(nit.sum()),
               largest_bin=int(max(len(p[0]) for p in probs)), max_iterations=int(nit.max()))
    try:
        from sklearn.cluster import KMeans
        t0 = time.perf_counter(); n_cpu = 0; n_diff = 0; n_bad = 0
        for t, (Xb, wb, sb) in enumerate(probs[:cpu_bins]):
            km = KMeans(n_clusters=len(sb), init=Xb[sb], n_init=1, random_state=0).fit(Xb, sample_weight=wb)
            c = okm.compare_labels(Xb, km.cluster_centers_, km.labels_, lab[po[t]:po[t + 1]])
            n_diff += c['n_differ']; n_bad += c['n_unexplained']; n_cpu += len(Xb)
        dt = time.perf_counter() - t0
        res.update(cpu_sklearn_bins=min(cpu_bins, len(probs)), cpu_sklearn_contigs=n_cpu, cpu_sklearn_s=dt,
                   cpu_sklearn_extrapolated_s=dt * tot / max(n_cpu, 1), labels_differ=n_diff, labels_unexplained=n_bad)
    except ImportError:
        pass
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='C2')
    ap.add_argument('--n', type=int, default=0)
    ap.add_argument('--combined', action='store_true')
    ap.add_argument('--cpu-check', action='store_true')
    ap.add_argument('--sklearn-graphs', action='store_true', help='cpu-check against live scikit-learn graphs (minutes at 200k)')
    a = ap.parse_args()
    if a.config == 'C2':
        r = short_read(a.n or 200000, 1, False, cpu_check=a.cpu_check, sklearn_graphs=a.sklearn_graphs)
    elif a.config == 'C3':
        r = short_read(a.n or 1000000, 10, a.combined, cpu_check=a.cpu_check, sklearn_graphs=a.sklearn_graphs)
    elif a.config == 'C4':
        r = long_read(a.n or 500000, 1, cpu_check=a.cpu_check)
    elif a.config == 'C5':
        r = scaling_case(a.n or 4000000)
    elif a.config == 'N4':
        r = recluster_case(a.n or 1000000)
    elif a.config == 'C3dist':
        r = short_read_sharded(a.n or 1000000, 10, a.combined)
        if r is None:
            return
    else:
        raise SystemExit('unknown config')
    r['config'] = a.config
    print(json.dumps(r), flush=True)


if __name__ == '__main__':
    main()
