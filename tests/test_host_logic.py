"""CPU tier: host-side logic of the drop-ins (threshold search, row-block
sharding, gather) incl. a world_size=2 gloo run with the oracle standing in for
the per-rank kernel call."""
import os
import sys

import numpy as np
import pytest

from semibin_b200 import cluster as sc, dist as sd, synth
from oracle import edges as oe, knn as ok


def test_threshold_sequence_matches_reference_loop():
    assert sc.threshold_sequence() == oe.threshold_sequence()


@pytest.mark.parametrize('max_node', [1, 0.95, 0.6, 0.3, 0.0])
def test_pick_threshold(max_node):
    rng = np.random.default_rng(3)
    for _ in range(20):
        rowmax = rng.random(997) ** rng.uniform(0.3, 3)
        counts = [int(np.sum(rowmax > t)) for t in sc.threshold_sequence()]
        assert sc.pick_threshold(counts, 997, max_node) == oe.max_node_threshold(rowmax, 997, max_node)


def test_row_block_partition():
    for n in (1, 7, 128, 1000003):
        for w in (1, 2, 3, 8):
            blocks = [sd.row_block(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    X = synth.embeddings(1001, seed=11)

    def local_knn(Xt, k, lo, hi):
        d, i = ok.knn_restated(Xt.numpy(), k, q_begin=lo, q_end=hi)
        return torch.from_numpy(d.astype(np.float32)), torch.from_numpy(i.astype(np.int32))

    d, i = sd.knn_sharded(torch.from_numpy(X), 17, local_knn=local_knn)
    c = torch.tensor([rank + 1, 10 * (rank + 1)], dtype=torch.int64)
    sd.allreduce_counts(c)
    lo, hi = sd.row_block(1001, rank, world)
    ex = torch.arange(lo, lo + 3 + rank, dtype=torch.int32)
    gx, gy, gv = sd.gather_edges(ex, ex + 1, ex.double() * 0.5)
    np.savez(os.path.join(tmp, f'r{rank}.npz'), d=d.numpy(), i=i.numpy(), c=c.numpy(), gx=gx.numpy(), gv=gv.numpy())
    dist.destroy_process_group()


def test_sharded_knn_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    X = synth.embeddings(1001, seed=11)
    d, i = ok.knn_restated(X, 17)
    for r in range(2):
        z = np.load(tmp_path / f'r{r}.npz')
        np.testing.assert_array_equal(z['i'], i.astype(np.int32))          # 1-rank result == gathered result
        np.testing.assert_array_equal(z['d'], d.astype(np.float32))
        np.testing.assert_array_equal(z['c'], [3, 30])
        np.testing.assert_array_equal(z['gx'], [0, 1, 2, 501, 502, 503, 504])
        np.testing.assert_array_equal(z['gv'], np.array([0, 1, 2, 501, 502, 503, 504]) * 0.5)


def test_igraph_handoff_contiguous_buffer():
    """cluster.py:150-157: the edge list reaches igraph as one [E,2] buffer (or, for an igraph without the
    memoryview path, as the same pairs in nested lists); order and endpoints unchanged."""
    from semibin_b200.cluster import add_edges_contiguous

    class NewGraph:
        def add_edges(self, es):
            assert isinstance(es, memoryview) and es.ndim == 2 and es.shape[1] == 2 and es.format in ('l', 'q')
            self.got = np.asarray(es).copy()

    class OldGraph:
        def add_edges(self, es):
            if isinstance(es, memoryview):
                raise TypeError('only sequences of pairs')
            self.got = np.asarray(es)

    X = np.array([0, 0, 3], dtype=np.int32)
    Y = np.array([1, 5, 4], dtype=np.int32)
    for G in (NewGraph, OldGraph):
        g = G()
        add_edges_contiguous(g, X, Y)
        np.testing.assert_array_equal(g.got, np.stack([X, Y], axis=1))


def test_plan_chunk_rows_whole_waves():
    """Streamed host path: chunks are whole waves of the filter grid, the last chunk is the short remainder, and the
    total number of waves never grows."""
    from semibin_b200.neighbors import plan_chunk_rows
    wave = 74 * 256
    for n_q in (1000000, 500000, 125000, 250000, 4000000):
        c = plan_chunk_rows(n_q, 4, wave)
        assert c % wave == 0 and c % 256 == 0
        n_chunks = -(-n_q // c)
        assert 2 <= n_chunks <= 4
        last = n_q - (n_chunks - 1) * c
        assert 0 < last <= c
        waves_split = (n_chunks - 1) * (c // wave) + -(-last // wave)
        assert waves_split == -(-n_q // wave)
    assert plan_chunk_rows(30000, 4, wave) == 30000                  # under two waves: not split
    assert plan_chunk_rows(140000, 3, wave, min_block_rows=1) == 4 * wave


def test_marker_masks_and_features():
    from semibin_b200.long_read import marker_masks, long_read_features
    m = marker_masks([np.array([0, 5]), np.array([], dtype=int), np.array([64, 70, 106])])
    assert m.dtype == np.uint64 and m.shape == (3, 2)
    assert m[0, 0] == (1 << 0) | (1 << 5) and m[0, 1] == 0 and not m[1].any()
    assert m[2, 0] == 0 and m[2, 1] == (1 << 0) | (1 << 6) | (1 << 42)
    with pytest.raises(ValueError):
        marker_masks([np.array([300])])
    emb = np.ones((4, 100), np.float32)
    depth = np.array([[2.0, 9.0, 0.0, 1.0]] * 4, np.float32)         # [mean_0, var_0, mean_1, var_1]
    x = long_read_features(emb, depth, 2, False)                     # long_read_cluster.py:75-81
    assert x.shape == (4, 102) and x.dtype == np.float32
    np.testing.assert_allclose(x[0, 100:], np.log(np.clip(np.array([2.0, 0.0], np.float32), 1e-6, None)))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's own scikit-learn path on the host cores): one JSON line with the
    keys the driver reads."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--contigs', '3000',
                          '--neighbors', '20', '--steps', '1', '--warmup', '1'], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['metric'] == 'kNN-graph contigs/sec' and line['unit'] == 'contigs/s'
    assert line['higher_is_better'] is True and line['value'] > 0 and line['n_gpus'] == 1
    assert line['cpu_baseline']['kind'] == 'reference' and line['cpu_baseline']['cores'] >= 1
    assert line['cpu_baseline']['value'] == line['value'] == line['e2e']['value']
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in line['config'] and 'model' not in line['config']
    # ranks other than 0 print nothing and exit 0 under torchrun
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--contigs', '3000',
                          '--neighbors', '20', '--steps', '1', '--warmup', '1'], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''


@pytest.mark.parametrize('world', [1, 2, 3, 4, 5, 6, 7, 8])
def test_sym_shard_schedule_tests_every_ordered_tile_pair_once(world):
    """The tile products `sbk_knn_sym_shard` launches (sbk_debug_sym_shard_schedule returns the very list the launcher
    walks; host only): over all ranks every ORDERED pair of tiles must be tested exactly once -- a product (I, J) tests
    the rows of I against J and, unless the circulant own block holds both orders of the pair itself (diagonal, opposite
    tiles of an even block), the rows of J against I.  Tile counts: fewer tiles than ranks allow blocks for, ragged last
    blocks, odd and even block lengths, the BASELINE size (3907 tiles)."""
    import ctypes as C
    from semibin_b200 import _lib
    lib = _lib.load()
    for T in (8, 12, 13, 37, 161, 1000, 3907):
        seen = np.zeros((T, T), dtype=np.uint8)
        n_products = 0
        for rank in range(world):
            buf = np.zeros(5 * 16, dtype=np.int64)
            ns = lib.sbk_debug_sym_shard_schedule(T, world, rank, buf.ctypes.data_as(C.c_void_p), 16)
            assert 0 <= ns <= 16
            for kind, q_lo, q_len, j_lo, j_len in buf[:5 * ns].reshape(ns, 5):
                if kind == 0:
                    assert (j_lo, j_len) == (q_lo, q_len)
                    I = np.arange(q_lo, q_lo + q_len)
                    for s in range(q_len // 2 + 1):
                        J = q_lo + (I - q_lo + s) % q_len
                        seen[I, J] += 1
                        if not (s == 0 or 2 * s == q_len):
                            seen[J, I] += 1
                        n_products += q_len
                else:
                    assert j_lo + j_len <= q_lo or q_lo + q_len <= j_lo      # never its own block
                    seen[q_lo:q_lo + q_len, j_lo:j_lo + j_len] += 1
                    seen[j_lo:j_lo + j_len, q_lo:q_lo + q_len] += 1
                    n_products += q_len * j_len
        assert seen.min() == 1 and seen.max() == 1, (T, world, int(seen.min()), int(seen.max()))
        assert n_products <= T * (T // 2 + 1) + world * T               # half the one-sided T * T products
    assert lib.sbk_debug_sym_shard_schedule(8, 9, 0, None, 0) < 0


def test_bench_clock_sampler_window():
    """bench.py's clock summary: samples inside the timed region count; a timed region shorter than the sampling period
    (8 GPUs: 5 steps of 27 ms) falls back to the loaded samples of the warm-up steps and says so."""
    import datetime
    import importlib.util
    import time
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('bench_for_test', os.path.join(root, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ['bench.py']
    try:
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv
    now = time.time()
    fmt = lambda t: datetime.datetime.fromtimestamp(t).strftime('%Y/%m/%d %H:%M:%S.%f')[:-3]
    s = bench.ClockSampler(0)
    s.lines = [f'{fmt(now - 1.0)}, 1800, 1965, 100, Not Active, Not Active, Not Active, Active',
               f'{fmt(now - 0.5)}, 1900, 1965, 10, Not Active, Not Active, Not Active, Not Active',
               f'{fmt(now + 0.1)}, 1950, 1965, 99, Not Active, Not Active, Not Active, Not Active',
               'garbage line']
    s.t_begin, s.t_end = now, now + 0.2
    out = s.summary()
    assert out['samples'] == 1 and out['sm_mhz'] == 1950.0 and out['reasons'] == [] and out['window'] == 'timed steps'
    s.t_begin, s.t_end = now + 0.12, now + 0.2
    out = s.summary()
    assert out['samples'] == 2 and out['sm_mhz'] == 1875.0 and out['reasons'] == ['sw_power_cap']
    assert out['window'].startswith('warm-up') and out['sm_max_mhz'] == 1965.0


class _FakeShardLib:
    """Stands in for libsbk in the world-2 gloo test of `knn_sharded_sym`'s control flow: calls the barrier callback three
    times like sbk_knn_sym_shard (failing rank: reports the failure at `fail_at`, everybody leaves at the first vote that
    says so), or fails only at the end (a mailbox overflow is seen by the receiver alone)."""

    def __init__(self, rank, fail_rank, fail_at):
        self.rank, self.fail_rank, self.fail_at = rank, fail_rank, fail_at
        self.votes = 0

    def sbk_knn_workspace_bytes(self, *a):
        return 64

    def sbk_last_error(self):
        return b'fake: mailbox overflowed'

    def sbk_knn_sym_shard(self, X, dtype, n, d, ld, k, rank, world, xchg, barrier, barrier_arg, *rest):
        import ctypes as C
        vote = C.cast(barrier, C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int))
        for phase in range(3):
            mine = int(self.rank == self.fail_rank and self.fail_at == phase)
            self.votes += 1
            if vote(None, mine):
                return -4
        return -4 if (self.rank == self.fail_rank and self.fail_at == 3) else 0


def _shard_vote_worker(rank, world, port, tmp):
    import types
    import torch
    import torch.distributed as dist
    from semibin_b200 import _lib, neighbors
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    X = torch.zeros((64, 8), dtype=torch.float32)
    bufs = types.SimpleNamespace(group=dist.group.WORLD, dist=torch.zeros(1), idx=torch.zeros(1, dtype=torch.int32), n_peers=0,
                                 peer_dist=None, peer_idx=None)
    shard = types.SimpleNamespace(shape=(64, 8, 5, torch.float32), rank=rank, world=world, xchg=None)
    fell_back = []
    sd.knn_sharded_p2p = lambda X, k, bufs, algo='auto', return_stats=False: (fell_back.append(algo) or (bufs.dist, bufs.idx, {}))
    _lib.stream_ptr = lambda dev: None
    neighbors._workspace = lambda nbytes, dev: torch.zeros(int(nbytes), dtype=torch.uint8)
    rows = []
    for fail_rank, fail_at in ((-1, -1), (1, 0), (0, 1), (1, 2), (0, 3)):
        fake = _FakeShardLib(rank, fail_rank, fail_at)
        _lib.load = lambda fake=fake: fake
        fell_back.clear()
        _, _, st = sd.knn_sharded_sym(X, 5, bufs, shard, return_stats=True)
        rows.append((fake.votes, list(fell_back), st.get('sym_shard_fallback')))
    import json
    with open(os.path.join(tmp, f'vote{rank}.json'), 'w') as fh:
        json.dump(rows, fh)
    dist.destroy_process_group()


def test_sharded_symmetric_vote_gloo_world2(tmp_path):
    """Host side of the sharded symmetric build on two gloo ranks with a fake library: whichever rank fails and wherever,
    BOTH ranks leave the build at the same vote and both redo it by row blocks; a build that held everywhere is kept."""
    import json
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_shard_vote_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = [json.load(open(tmp_path / f'vote{r}.json')) for r in range(2)]
    expect_votes = [3, 1, 2, 3, 3]            # votes called until everybody has left
    for r in range(2):
        for case, (votes, fell_back, why) in enumerate(got[r]):
            assert votes == expect_votes[case], (r, case, votes)
            if case == 0:
                assert fell_back == [] and why is None
            else:
                assert fell_back == ['tensor_onesided'] and why, (r, case, fell_back, why)
