"""Sharded symmetric build (`sbk_knn_sym_shard`, semibin_b200.dist.knn_sharded_sym): one process per GPU under torchrun,
NCCL + symmetric memory.  Needs at least two GPUs (skipped on a one-GPU box; scripts/gpu_two_gpus.sh runs it)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(world, *script_args, port=29533):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(ROOT, 'scripts', 'shard_sym_check.py'), *script_args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    return [json.loads(l) for l in r.stdout.splitlines() if l.startswith('{')]


def _worlds():
    """World sizes to run: an even one (the opposite block pair is shared by two ranks) and, with three or more GPUs, an
    odd one.  SBK_TEST_WORLD8=1 on an eight-GPU box adds eight (ranks without any tile at n = 3000; the 1M build on eight
    GPUs is checked by scripts/shard_sym_check.py, profiles/r02shard_trace_8gpu.log)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs two GPUs')
    worlds = [w for w in (2, 3, 4) if w <= n and (w != 2 or n < 4)]
    if n >= 8 and os.environ.get('SBK_TEST_WORLD8') == '1':
        worlds.append(8)
    return worlds


def _world():
    return _worlds()[-1]


@pytest.mark.parametrize('slot', [0, 1, 2])
def test_sharded_symmetric_build_equals_exact_and_row_blocks(slot):
    """3000 x k=50 (12 tiles: blocks of one or two tiles, ragged last block, ranks without rows at world 8) and
    41k x k=50 (every rank: circulant own block, full and split block pairs, mailboxes): bitwise the exact path's and the
    row-block build's result on every rank; 450k x k=200 against the row-block build on the largest world."""
    worlds = _worlds()
    if slot >= len(worlds):
        pytest.skip('no further world size on this box')
    world = worlds[slot]
    sizes = ['3000', '41000'] + (['450000'] if world == worlds[-1] else [])
    rows = _torchrun(world, '--iters', '1', *sizes, port=29535 + slot)
    assert len(rows) == len(sizes) * world
    for r in rows:
        assert r['sym_equals_rowblock'], r
        assert r['sym']['symmetric'] == 1 and 'sym_shard_fallback' not in r['sym'], r
        if r['n'] <= 100000:
            assert r['sym_equals_exact'], r


def test_sharded_symmetric_build_mailbox_overflow_votes_for_row_blocks():
    """Mailboxes of 1024 records: the receiving ranks see the overflow, the vote sends all ranks to the row-block build
    and the result is still the exact one."""
    world = _world()
    rows = _torchrun(world, '--starve', '--iters', '1', '41000', port=29534)
    assert len(rows) == world
    for r in rows:
        assert 'sym_shard_fallback' in r['sym'], r
        assert r['sym_equals_rowblock'] and r['sym_equals_exact'], r
