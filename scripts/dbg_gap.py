"""Where does the time between the stage timers go?  (debug helper, not part of the product)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from semibin_b200 import synth
from semibin_b200.neighbors import knn
import bench

n, k = 1_000_000, 200
X = torch.from_numpy(synth.embeddings(n, seed=20242)).cuda()
for _ in range(2):
    knn(X, k)
torch.cuda.synchronize()
def run(tag, hold=False, sampler=False):
    s = None
    if sampler:
        s = bench.ClockSampler(0); s.start()
    keep = None
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d, i, st = knn(X, k, return_stats=True)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        ssum = sum(st[x] for x in ('ms_prepare', 'ms_sample', 'ms_filter', 'ms_rerank', 'ms_fallback'))
        print(tag, it, 'ev_ms=%.1f host_call_ms=%.1f host_sync_ms=%.1f stage_sum=%.1f' % (e0.elapsed_time(e1), (t1 - t0) * 1e3, (t2 - t0) * 1e3, ssum),
              {x: round(st[x], 1) for x in ('ms_prepare', 'ms_sample', 'ms_filter', 'ms_rerank', 'ms_fallback')}, flush=True)
        if hold:
            keep = (d, i)
        else:
            del d, i
    if s:
        s.stop_flag = True; s.join()
        print(tag, s.summary())
run('plain')
run('hold', hold=True)
run('sampler', sampler=True)
run('hold+sampler', hold=True, sampler=True)
