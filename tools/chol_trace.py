"""Per-task timeline of the tile-task Cholesky kernel (development aid).
    python tools/chol_trace.py [n]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleekit_b200 import _lib, ops  # noqa: E402
from sleekit_b200 import workloads as wl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 768
_, H, _ = wl.synthetic_layer(4, n, 0, samples=4096)
Hd = torch.from_numpy(H).cuda()
damp = ops.damp_value(Hd, 0.01)
order = ops.argsort(ops.order_keys(Hd, damp, None))
T = (n + 63) // 64
nt = T * (T + 1) // 2
for _ in range(3):
    ops.chol_factor(Hd, order, damp)
buf = torch.zeros(nt * 16, dtype=torch.int64, device="cuda")
_lib.call("slk_debug_chol_trace", ctypes.c_void_p(buf.data_ptr()))
ops.chol_factor(Hd, order, damp)
torch.cuda.synchronize()
_lib.call("slk_debug_chol_trace", None)
trf = buf.cpu().numpy().reshape(nt, 16)
tr = trf[:, :8]
t0 = tr[:, 2].min()
clk = 1.0  # clock64 ticks -> reported raw (SM clock)
print(f"n={n} T={T} tasks={nt} kernel span {(tr[:, 3].max() - t0) / 1e3:.1f} us")
print(" i  j   start_us   end_us   kloop_clk   math_clk   publish_clk")
for row in tr:
    i, j, gs, ge, c0, c1, c2, c3 = row
    if i == j or i == j + 1:
        print(f"{i:3d}{j:3d} {(gs - t0) / 1e3:9.2f} {(ge - t0) / 1e3:9.2f} {c1 - c0:10d} {c2 - c1:10d} {c3 - c2:10d}")
d = tr[tr[:, 0] == tr[:, 1]]
o = tr[tr[:, 0] != tr[:, 1]]
print("diag  mean clk: kloop %.0f math %.0f publish %.0f" % tuple((d[:, k + 1] - d[:, k]).mean() for k in (4, 5, 6)))
if len(o):
    print("offd  mean clk: kloop %.0f math %.0f publish %.0f" % tuple((o[:, k + 1] - o[:, k]).mean() for k in (4, 5, 6)))
dd = trf[trf[:, 0] == trf[:, 1]]
print("diag factor clk: panels %.0f trailing %.0f diag-inv %.0f doubling %.0f" % (dd[:, 8].mean(), dd[:, 9].mean(), (dd[:, 11] - dd[:, 10]).mean(), (dd[:, 12] - dd[:, 11]).mean()))
