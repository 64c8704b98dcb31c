"""Per-task timeline of the tile-task Cholesky kernel, single GPU or distributed (development aid).
    python tools/chol_trace_dist.py n                (one GPU)
    torchrun --nproc-per-node P tools/chol_trace_dist.py n      (tile rows cyclic over P GPUs; rank 0 reports)"""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleekit_b200 import _lib, ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", torch.cuda.current_device())
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device=dev)
g.manual_seed(5)
X = torch.randn((2048, n), generator=g, device=dev)
H = torch.zeros((n, n), device=dev)
m = torch.zeros(n, device=dev)
ops.hessian_accum(X, H, m, 0.0, 2048)
del X
damp = ops.damp_value(H, 0.01)
order = ops.argsort(ops.order_keys(H, damp, None))
T = (n + 63) // 64
nt = T * (T + 1) // 2
token = torch.zeros(1, device=dev)


def barrier():
    if world > 1:
        dist.all_reduce(token)


pws = ops.PeerWorkspace(ops.chol_dist_ws_bytes(n)) if world > 1 else None


def factor():
    if world > 1:
        return ops.chol_factor_dist(H, order, damp, pws, barrier, want_rt=False)
    return ops.chol_factor(H, order, damp, want_rt=False)


for _ in range(2):
    factor()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
barrier()
e0.record()
factor()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
buf = torch.zeros(nt * 16, dtype=torch.int64, device=dev)
_lib.call("slk_debug_chol_trace", ctypes.c_void_p(buf.data_ptr()))
factor()
torch.cuda.synchronize()
_lib.call("slk_debug_chol_trace", None)
if rank == 0:
    tr = buf.cpu().numpy().reshape(nt, 16)
    tr = tr[tr[:, 2] > 0]                       # tasks this rank ran (tickets are per rank)
    t0 = tr[:, 2].min()
    span = (tr[:, 3].max() - t0) / 1e3
    off = tr[tr[:, 0] != tr[:, 1]]
    dg = tr[tr[:, 0] == tr[:, 1]]
    kl = (off[:, 5] - off[:, 4]) / np.maximum(off[:, 1], 1)          # clocks per 64-wide k step
    print(f"n={n} T={T} ranks={world}: factor {ms:.2f} ms (gather + tile kernel + export), tile-kernel span on rank 0 "
          f"{span / 1e3:.2f} ms, tasks on rank 0 {len(tr)}")
    for lo, hi in ((1, T // 4), (T // 4, T // 2), (T // 2, 3 * T // 4), (3 * T // 4, T)):
        sel = (off[:, 1] >= lo) & (off[:, 1] < hi)
        if sel.any():
            dur = (off[sel, 3] - off[sel, 2]) / 1e3
            print(f"  columns {lo:4d}-{hi:4d}: {int(sel.sum()):6d} off-diagonal tasks, k-loop {np.median(kl[sel]):8.0f} clk per k step "
                  f"(median; incl. waiting for tiles), task duration median {np.median(dur):8.1f} us")
    # chain: start-to-start distance of consecutive diagonal tasks this rank owns
    if len(dg) > 2:
        dg = dg[np.argsort(dg[:, 1])]
        gaps = np.diff(dg[:, 3]) / 1e3 / np.diff(dg[:, 1])
        print(f"  diagonal tiles on rank 0: {len(dg)}; end-to-end gap per column: median {np.median(gaps):.1f} us, "
              f"90th pct {np.percentile(gaps, 90):.1f} us")
if pws is not None:
    pws.close()
if world > 1:
    dist.destroy_process_group()
