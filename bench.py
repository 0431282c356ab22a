#!/usr/bin/env python
"""bench.py -- kNN-graph contigs/sec on the headline workload of BASELINE.json:
N = 1,000,000 synthetic contigs, embedding width 100 (float32), k = 200
(SemiBin/cluster.py:94-99, max_edges default 200), exact neighbour sets.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--contigs N] [--neighbors K]

A "step" is one complete kNN-graph build over all N contigs: FP16 staging,
tcgen05 threshold + filter passes over all N x N pairs, FP64 rerank with the
exactness certificate, and (N > 1 GPUs) the NCCL all-gather of the row blocks.
`value` is timed with the embedding already resident in HBM; `e2e` goes through
the public host-buffer call (pinned host embedding in, pinned host (dist, idx)
out, copies inside the timed region).
N > 1: one process per GPU (torchrun), query rows sharded, references
replicated, total work fixed -> "scaling": "strong".
--impl reference times the reference's own path (scikit-learn brute-force
kneighbors, the code SemiBin calls) on the host cores on a bounded row sample.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# torchrun exports OMP_NUM_THREADS=1; the CPU legs (reference arm, cpu_baseline) must use every host
# core, as SemiBin does through --threads (SemiBin/utils.py:71-83).  Set before numpy/sklearn load.
if ('reference' in sys.argv) or os.environ.get('RANK', '0') == '0':
    os.environ['OMP_NUM_THREADS'] = str(os.cpu_count())     # rank 0 runs the scikit-learn legs (baseline, parity spot check)

import numpy as np  # noqa: E402

N_DEFAULT = 1_000_000
K_DEFAULT = 200
D_EMB = 100
SEED = 20242


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def filter_traffic():
    """DRAM bytes per filter launch from the newest committed ncu --set full capture (profiles/)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', '*_filter_traffic.json')))
    if not files:
        return None
    try:
        j = json.load(open(files[-1]))
        return int(j['dram_bytes_read_per_launch']) + int(j['dram_bytes_write_per_launch'])
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return dict(tflops=float(j.get('bf16_tflops_sustained', j['bf16_tflops'])), hbm=float(j['hbm_gbs']),
                        source='MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)')
        except Exception:
            pass
    return dict(tflops=1400.0, hbm=6650.0, source='fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)')


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs: the `nvidia-smi -lms 200`
    line of B200_PROFILING.md as a child process (an in-process NVML poller can hold driver locks the CUDA
    launches of this process need; a 50 ms pynvml thread cost an occasional 100+ ms stall in one step)."""
    FIELDS = ('timestamp,clocks.sm,clocks.max.sm,utilization.gpu,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.stop_flag = False
        self.err = None
        self.lines = []
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def start(self):
        import subprocess
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def join(self, timeout=None):
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=5)
            self.lines = [ln for ln in out.splitlines() if ln.strip()]
        except Exception as e:  # pragma: no cover
            self.err = repr(e)
            try:
                self.proc.kill()
            except Exception:
                pass

    def summary(self):
        import datetime
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        rows, max_mhz = [], None
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                rows.append((ts, float(f[1]), float(f[3]), [nm for nm, v in zip(names, f[4:8]) if v.lower().startswith('active')]))
                max_mhz = float(f[2])
            except ValueError:
                continue
        # the child is started before the warm-up steps (its start-up takes longer than a short timed region); the samples
        # of the timed steps count, and when the timed region is shorter than the sampling period, those of the warm-up
        # steps right before it (same kernels, same load)
        window = 'timed steps'
        inside = [r for r in rows if self.t_begin is None or (self.t_begin <= r[0] <= (self.t_end or r[0]))]
        if not inside and rows:
            inside = [r for r in rows if r[2] >= 50] or rows[-2:]
            window = 'warm-up + timed steps (no sample fell inside the timed region)'
        samples = [(m, u) for _, m, u, _ in inside]
        reasons = {nm for r in inside for nm in r[3]}
        loaded = [m for m, u in samples if u >= 50] or [m for m, _ in samples]
        med = float(np.median(loaded)) if loaded else None
        out = {'sm_mhz': med, 'sm_max_mhz': max_mhz, 'reasons': sorted(reasons), 'samples': len(samples), 'window': window,
               'how': 'nvidia-smi -lms 200 child process started before the warm-up'}
        if self.err:
            out['error'] = self.err
        return out


_DATA = {'file': None, 'genome_size': 200}


def synth_embeddings(n):
    """The workload matrix: seeded synthetic mixture (default: genomes of ~200 contigs), a heavier-tailed mix
    (--genome-size), or a matrix from a file (--embedding-file: e.g. synthetic k-mer features pushed through a shipped
    SemiBin model by scripts/make_real_embeddings.py)."""
    from semibin_b200 import synth
    if _DATA['file']:
        X = np.ascontiguousarray(np.load(_DATA['file'])[:n], dtype=np.float32)
        assert X.shape[0] == n, f'{_DATA["file"]} holds {X.shape[0]} rows, --contigs {n} asked'
        return X
    return synth.embeddings(n, seed=SEED, mean_size=_DATA['genome_size'])


def reference_knn_block(X, k, rows):
    """The reference's own code path for a block of query rows: sklearn brute
    force (ArgKmin32), the call SemiBin/cluster.py:94-99 makes."""
    from oracle import knn as ok
    return ok.knn_sklearn(X, k, q_begin=rows[0], q_end=rows[1])


def run_reference(args):
    rank = env_int('RANK', 0)
    if rank != 0:
        return 0
    n, k = args.n, args.k
    X = synth_embeddings(n)
    cores = os.cpu_count()
    total_steps = args.steps + args.warmup
    # bounded sample, but never fewer than 4 * 256 * cores query rows: scikit-learn then parallelises over X chunks
    # (parallel_on_X, _argkmin.pyx.tp) exactly as it does for the full graph; below that it switches to parallel_on_Y
    block = max(256, min(n, int(8.2e9 / n)))
    if total_steps > 8:
        block = max(256, block * 8 // total_steps)
    block = min(n, max(block, 4 * 256 * cores))
    lo = (n // 2 // block) * block
    for _ in range(args.warmup):
        reference_knn_block(X, k, (lo, lo + block))
    t0 = time.perf_counter()
    for s in range(args.steps):
        reference_knn_block(X, k, (lo, lo + block))
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = block / dt
    sample = (f'{block} query rows x all {n} references per step (sklearn NearestNeighbors brute, k+1={k + 1}), contigs/s = rows/s: '
              f'EXTRAPOLATED to the full graph from this row block (>= 4*256*cores rows, so the chunking/threading strategy is the full graph\'s)')
    line = {
        'impl': 'reference', 'metric': 'kNN-graph contigs/sec', 'value': value, 'unit': 'contigs/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': f'kNN graph, N={n} contigs x D={D_EMB} float32 embedding, k={k} (BASELINE configs[2] shape / metric quote)',
                   'n': n, 'd': D_EMB, 'k': k},
        'extrapolated': True,
        'cpu_baseline': {'value': value, 'unit': 'contigs/s', 'cores': cores, 'kind': 'reference', 'sample': sample, 'extrapolated': True},
        'e2e': {'value': value, 'unit': 'contigs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from semibin_b200 import _lib
    from semibin_b200.neighbors import knn, knn_host, kneighbors_graph
    from semibin_b200 import dist as sdist

    rank = env_int('RANK', 0)
    world = env_int('WORLD_SIZE', 1)
    local_rank = env_int('LOCAL_RANK', 0)
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    # host-side barrier for the phases in which ONE process drives every GPU: an NCCL barrier would park a spinning
    # kernel of the waiting ranks on the very GPUs rank 0 is about to use from its own context
    cpu_group = dist.new_group(backend='gloo') if world > 1 else None
    _lib.require_cuda()
    n, k = args.n, args.k
    X_host = torch.from_numpy(synth_embeddings(n)).pin_memory()
    X = X_host.cuda()
    lo, hi = sdist.row_block(n, rank, world)
    n_loc = hi - lo
    out_d_host = torch.empty((n_loc, k), dtype=torch.float32).pin_memory()
    out_i_host = torch.empty((n_loc, k), dtype=torch.int32).pin_memory()

    # caller-owned device outputs, reused by every step (a fresh 1.6 GB allocation inside a timed step costs a
    # cudaMalloc that showed up as a 80-140 ms outlier in the second step)
    dev_d = torch.empty((n_loc, k), dtype=torch.float32, device='cuda')
    dev_i = torch.empty((n_loc, k), dtype=torch.int32, device='cuda')

    full_d = torch.empty((n, k), dtype=torch.float32, device='cuda') if (world > 1 and n % world == 0) else None
    full_i = torch.empty((n, k), dtype=torch.int32, device='cuda') if (world > 1 and n % world == 0) else None
    # N > 1: the rerank kernels store every finished row into every rank's full result buffer (P2P over NVLink,
    # symmetric memory), the gather rides on the compute kernel; --gather nccl (or a box without peer mapping)
    # falls back to two all-gathers after the kernels
    peer_bufs, gather_mode, sym_shard = None, 'none', None
    sharding = f'row-block x{world}, references replicated'
    if world > 1:
        gather_mode = 'nccl all_gather_into_tensor x2'
        if args.gather in ('p2p', 'sym'):
            try:
                peer_bufs = sdist.PeerResult(n, k, torch.float32, X.device)
                gather_mode = 'p2p stores from the rerank kernels into symmetric-memory result buffers + barrier'
            except Exception as e:      # noqa: BLE001  (reported in the JSON line)
                gather_mode = f'nccl all_gather_into_tensor x2 (symmetric memory unavailable: {type(e).__name__}: {e})'[:300]
        # --gather sym (default): the tile pairs are split over the ranks so that each is multiplied once in the whole
        # job (sbk_knn_sym_shard); candidates for other ranks' rows travel through NVLink mailboxes.  A shape the
        # symmetric filter does not take falls back to row blocks (every pair multiplied twice, no exchange).
        if peer_bufs is not None and args.gather == 'sym' and args.algo in ('auto', 'tensor', 'tensor_symmetric'):
            try:
                sym_shard = sdist.SymShard(n, X.shape[1], k, torch.float32, X.device)
                sharding = f'symmetric tile pairs x{world} (each pair multiplied once), candidate mailboxes over NVLink'
            except ValueError as e:
                sharding += f' (symmetric sharding not taken: {e})'

    def step(collect=None):
        if sym_shard is not None:
            d, i, st = sdist.knn_sharded_sym(X, k, peer_bufs, sym_shard, return_stats=True)
        elif peer_bufs is not None:
            d, i, st = sdist.knn_sharded_p2p(X, k, peer_bufs, algo=args.algo, return_stats=True)
        else:
            d, i, st = knn(X, k, q_begin=lo, q_end=hi, algo=args.algo, return_stats=True, out=(dev_d, dev_i))
            if world > 1:
                d = sdist.all_gather_rows(d, n, out=full_d)
                i = sdist.all_gather_rows(i, n, out=full_i)
        if collect is not None:
            collect.append(st)
        return d, i

    def step_e2e():
        # H2D of this step's input, the build, D2H of this step's result: all inside the public host-buffer call
        knn_host(X_host, k, q_begin=lo, q_end=hi, out_dist=out_d_host, out_idx=out_i_host, algo=args.algo)
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    sampler.mark_begin()
    stats = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev0.record()
    step_evs[0].record()
    for s_i in range(args.steps):
        d, i = step(stats)
        step_evs[s_i + 1].record()
    ev1.record()
    barrier()
    sampler.mark_end()
    sampler.join()
    ms_total = ev0.elapsed_time(ev1)
    ms_each = [step_evs[j].elapsed_time(step_evs[j + 1]) for j in range(args.steps)]
    t = torch.tensor([ms_total], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps

    # end-to-end through the reference-facing call: sklearn-signature kneighbors_graph on a plain numpy matrix ->
    # scipy csr_matrix (cluster.py:94-99).  ONE process drives all N GPUs of the run (rank 0; the other ranks wait at
    # the barrier), exactly what an unmodified single-process SemiBin gets from patch.install() + SBK_DEVICES=N.
    X_np = X_host.numpy()
    e2e_steps = max(1, min(args.steps, 3))
    e2e_ms, g = None, None
    barrier()
    host_barrier()
    if rank == 0:
        g = kneighbors_graph(X_np, k, mode='distance', p=2, devices=list(range(world)))     # warm-up (pins the result arrays once)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            g = None        # the caller drops the previous graph: its pinned arrays go back to torch's host allocator
            g, g_stats = kneighbors_graph(X_np, k, mode='distance', p=2, devices=list(range(world)), return_stats=True)
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    host_barrier()
    barrier()
    # second figure: the SPMD host-buffer call (one process per GPU, each rank its own row block)
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    spmd_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    te = torch.tensor([spmd_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    spmd_ms = float(te.item())

    # dominant kernel: tcgen05 filter pass (per launch: one launch per step per rank)
    peaks = measured_peaks()
    filt_ms = float(np.mean([s['ms_filter'] for s in stats])) if stats else 0.0
    n_filter_launch = max(1, stats[0].get('n_filter_launches') or stats[0]['n_chunks']) if stats else 1
    flops_onesided = 2.0 * n_loc * n * D_EMB                      # algorithmic, one-sided: true D, no padding credit
    symmetric = bool(stats and stats[0].get('symmetric'))
    if symmetric:
        # the symmetric self-join filter multiplies every unordered pair of 256-row tiles once: T * (T // 2 + 1) tile
        # products instead of T * T; `achieved` counts the flops the kernel executes (true D), never the one-sided 2 N^2 D
        T = (n + 255) // 256
        flops = 2.0 * 256 * 256 * D_EMB * T * (T // 2 + 1) / world     # N > 1: the tile pairs are split over the ranks
    else:
        flops = flops_onesided
    roofline = None
    if filt_ms > 0:
        ach = flops / n_filter_launch / (filt_ms / n_filter_launch * 1e-3) / 1e12
        roofline = {'bound': 'tensor', 'kernel': 'tc_kernel<FILTER_SYM>' if symmetric else 'tc_kernel<FILTER>',
                    'achieved': ach, 'peak': peaks['tflops'],
                    'unit': 'TFLOP/s', 'frac': ach / peaks['tflops'],
                    'traffic': filter_traffic() if (n == N_DEFAULT and world == 1) else None,
                    'peak_source': peaks['source'], 'ms_per_launch': filt_ms / n_filter_launch,
                    'algorithmic_flops_per_launch': flops / n_filter_launch,
                    'symmetric': symmetric,
                    'note': ('executed tensor flops of the symmetric filter (half the one-sided product) over the whole filter stage, '
                             'record transposes and threshold refinements included; the same time against the one-sided 2*N^2*D '
                             'would read %.0f TFLOP/s' % (flops_onesided / (filt_ms * 1e-3) / 1e12)) if symmetric else None}
    stage_ms = {key: float(np.mean([s[key] for s in stats])) for key in
                ('ms_prepare', 'ms_sample', 'ms_filter', 'ms_rerank', 'ms_fallback', 'ms_exact_filter')} if stats else {}

    if rank == 0:
        line = {
            'metric': 'kNN-graph contigs/sec', 'value': n / (ms_step * 1e-3), 'unit': 'contigs/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'ms_each_step_rank0': ms_each,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f16 tensor-core filter + f64 rerank (exact result)', 'data': 'synthetic',
            'config': {'workload': f'kNN graph, N={n} contigs x D={D_EMB} float32 embedding, k={k} (BASELINE configs[2] shape / metric quote)',
                       'n': n, 'd': D_EMB, 'k': k, 'algo': args.algo,
                       'embedding': (f'file {os.path.basename(args.embedding_file)}' if args.embedding_file else
                                     f'synthetic Gaussian mixture, lognormal genome sizes (mean {args.genome_size} contigs)'), 'parallelism': sharding, 'gather': gather_mode,
                       'l2': 'inputs (400 MB embedding + 224 MB staged operands + candidate lists) exceed the 126 MB L2'},
            'clocks': sampler.summary(),
            'e2e': {'value': n / (e2e_ms * 1e-3), 'unit': 'contigs/s', 'ms_per_step': e2e_ms,
                    # whole-job bytes: the numpy matrix is uploaded to every device that takes part, every result row
                    # is downloaded once as (float64 distance, int64 index), the dtypes scipy / scikit-learn hand back
                    'h2d_bytes_per_step': int(X_host.numel() * 4) * world,
                    'd2h_bytes_per_step': int(n * k * 16),
                    'api': "semibin_b200.neighbors.kneighbors_graph(X_numpy, 200, mode='distance', p=2) -> scipy csr_matrix "
                           '(the call at SemiBin/cluster.py:94-99), one process driving all N GPUs (devices=N / SBK_DEVICES): numpy '
                           'matrix over pinned host memory in; one GPU: the symmetric build finishes its rows block row by block row and a '
                           'zero-copy scatter kernel writes every finished block as (float64, int64) into the pinned arrays of the '
                           'csr_matrix while later block rows are multiplied; several GPUs: device-widened row chunks downloaded '
                           'while the next chunk is computed',
                    'host_split_ms': {'alloc_result': g_stats['host_ms']['alloc_result'], 'build': g_stats['host_ms']['build'],
                                      'upload': g_stats.get('upload_ms'),
                                      'per_device': {str(dv): {q: round(v, 2) for q, v in ds.items() if isinstance(v, float)}
                                                     for dv, ds in g_stats.get('per_device', {}).items()}} if rank == 0 else None,
                    'spmd_host_buffers': {'value': n / (spmd_ms * 1e-3), 'ms_per_step': spmd_ms,
                                          'api': 'semibin_b200.neighbors.knn_host under torchrun (one process per GPU, (dist f32, idx i32) row '
                                                 'blocks into pinned host memory)', 'd2h_bytes_per_step': int(n * k * 8)}},
            'gpu_launches': int(sum(s['n_launches'] for s in stats)),
            'roofline': roofline,
            'stages_ms': stage_ms,
            'knn_stats': {kk: stats[-1][kk] for kk in ('algo_used', 'symmetric', 'n_chunks', 'n_filter_launches', 'n_fallback_rows', 'n_candidates',
                                                       'n_reranked', 'n_retry_rows', 'n_fail_overflow', 'n_fail_few', 'n_fail_survivors', 'n_fail_certificate',
                                                       'sample_size', 'cand_capacity', 'order_stat', 'kpad')} if stats else None,
        }
        if not args.no_cpu_baseline:
            # live scikit-learn on a row block: the reported CPU baseline (N = 1) and, at EVERY N, a parity spot check
            # of the result the reference-facing call just returned (the csr_matrix gathered from all N devices)
            Xn = X_host.numpy()
            block = max(256, min(n, int(8.2e9 / n)))
            if world > 1:
                block = max(256, block // 4)
            lo_b = (n // 2 // block) * block
            t0 = time.perf_counter()
            rd, ri = reference_knn_block(Xn, k, (lo_b, lo_b + block))
            dt = time.perf_counter() - t0
            from oracle import compare
            gd = g.data.reshape(n, k)[lo_b:lo_b + block]
            gi = g.indices.reshape(n, k)[lo_b:lo_b + block]
            res = compare.compare_knn(rd, ri, gd, gi, rtol=1e-5)
            base = {'value': block / dt, 'unit': 'contigs/s', 'cores': os.cpu_count(), 'kind': 'reference',
                    'sample': f'{block} query rows x all {n} references, live scikit-learn brute kneighbors (the call SemiBin makes), one run',
                    'parity_rows_bad': res['n_rows_bad'], 'parity_rows_tied': res['n_rows_tied'],
                    'parity_of': f'kneighbors_graph csr_matrix built on {world} device(s)'}
            # the device-resident (gathered) result of the timed steps against the same block
            res2 = compare.compare_knn(rd, ri, d[lo_b:lo_b + block].cpu().numpy().astype(np.float64),
                                       i[lo_b:lo_b + block].cpu().numpy().astype(np.int64), rtol=1e-5)
            base['parity_rows_bad_device_result'] = res2['n_rows_bad']
            if world == 1:
                line['cpu_baseline'] = base
            else:
                line['parity'] = base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--contigs', dest='n', type=int, default=N_DEFAULT)
    ap.add_argument('--neighbors', dest='k', type=int, default=K_DEFAULT)
    ap.add_argument('--algo', default='auto')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--gather', default='sym', choices=['sym', 'p2p', 'nccl'])
    ap.add_argument('--embedding-file', default=None, help='.npy float32 [N, 100] matrix instead of the synthetic mixture (realism runs)')
    ap.add_argument('--genome-size', type=int, default=200, help='mean contigs per synthetic genome (lognormal, sigma 1)')
    args = ap.parse_args()
    _DATA['file'], _DATA['genome_size'] = args.embedding_file, args.genome_size
    if args.impl == 'reference':
        return run_reference(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
