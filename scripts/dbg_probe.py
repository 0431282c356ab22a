"""Filter-kernel probes: SBK_TC_PROBE=3 (no accumulator read) / 4 (read only).  Debug helper."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semibin_b200 import synth, _lib
from semibin_b200.neighbors import knn
import ctypes as C
n, k = int(os.environ.get("N", 1000000)), 200
X = torch.from_numpy(synth.embeddings(n, seed=20242)).cuda()
lib = _lib.load()
N, D = X.shape
ws_bytes = lib.sbk_knn_workspace_bytes(_lib.SBK_F32, N, D, N, k, 2)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device='cuda')
od = torch.empty((N, k), dtype=torch.float32, device='cuda'); oi = torch.empty((N, k), dtype=torch.int32, device='cuda')
import bench
sm = bench.ClockSampler(0); sm.start()
for it in range(3):
    st = _lib.KnnStats()
    rc = lib.sbk_knn(_lib.ptr(X), _lib.SBK_F32, N, D, X.stride(0), 0, N, k, _lib.ptr(od), _lib.ptr(oi), _lib.ptr(ws), ws.numel(), 2, C.byref(st), _lib.stream_ptr())
    print('probe', os.environ.get('SBK_TC_PROBE'), 'rc', rc, 'ms_filter %.2f ms_sample %.2f' % (st.ms_filter, st.ms_sample), flush=True)
    if rc != 0:
        print(_lib.last_error()[:200])

sm.stop_flag = True; sm.join(); print('clocks', sm.summary())
