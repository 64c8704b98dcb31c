"""Phase clocks of the macro-block sweep kernel, CTA 0 (development aid).  python tools/sweep_trace.py r n"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleekit_b200 import _lib, codebook, ops  # noqa: E402
from sleekit_b200 import workloads as wl  # noqa: E402

r, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (768, 3072)
cb = codebook.UniformCodebook(8, -1, 1)
W, H, _ = wl.synthetic_layer(r, n, 0, samples=2048)
Wd, Hd = torch.from_numpy(W).cuda(), torch.from_numpy(H).cuda()
damp = ops.damp_value(Hd, 0.01)
order = ops.argsort(ops.order_keys(Hd, damp, None))
r32, rt32, ud32, info = ops.chol_factor(Hd, order, damp)
sc = ops.scale_search(Wd, cb, torch.linspace(0.05, 1, 100, device="cuda"), Hd.diagonal().contiguous())[0]
Ws = ops.scale_rows(Wd, sc, 0)
for _ in range(2):
    ops.gptq_sweep_r(Ws.clone(), r32, rt32, ud32, cb)
buf = torch.zeros(128, dtype=torch.int64, device="cuda")
_lib.call("slk_debug_sweep_trace", ctypes.c_void_p(buf.data_ptr()))
ops.gptq_sweep_r(Ws.clone(), r32, rt32, ud32, cb)
torch.cuda.synchronize()
_lib.call("slk_debug_sweep_trace", None)
t = buf.cpu().numpy().reshape(8, 16)
print("last macro block, CTA 0: per 32-column block clocks  wait | product+reduce | Ud multiply | leaf | copy")
for b in range(8):
    x = t[b]
    print(b, x[1] - x[0], x[2] - x[1], x[3] - x[2], x[4] - x[3], x[5] - x[4], " total", (t[b + 1][0] - x[0]) if b < 7 else "",
          " leaf: owner walks", x[6], "shuffle+update", x[7], "| phase 0: chain", x[8], "+stores", x[9], "shuffles", x[10],
          "leaf call", x[11])
