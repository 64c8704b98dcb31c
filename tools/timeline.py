"""Phase timeline of one graph replay of the layer-set pass (development aid).

    python tools/timeline.py [--e2e]

Stream-ordered %globaltimer stamps (slk_debug_timestamp) around the phases of every layer inside the
multi-stream CUDA graph: 0 start, 1 after copy-in, 2 after scale search, 3 after ordering / permutation
(stage A end), 4 / 5 begin / end of the layer's batched factor launch (stored on the group's first
layer), 6 sweep begin, 7 layer end.  Prints per-phase windows and how many layers sit in each phase
per half millisecond."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from sleekit_b200 import codebook, ops  # noqa: E402
from sleekit_b200 import workloads as wl  # noqa: E402
from sleekit_b200.pipeline import LayerSetQuantizer  # noqa: E402


def main():
    e2e = "--e2e" in sys.argv
    per_op = "--ops" in sys.argv
    only = [a.split("=")[1] for a in sys.argv if a.startswith("--only=")]
    dev = torch.device("cuda", 0)
    shapes = wl.layer_shapes("opt-125m")
    if only:
        shapes = [sh for sh in shapes if (sh[1] >= 2048) == (only[0] == "big")]
    cb = codebook.UniformCodebook(8, -1, 1)
    Wd, Hd = [], []
    for i, (r, n) in enumerate(shapes):
        Wd.append(torch.from_numpy(wl.synthetic_weight(r, n, i)).to(dev))
        x = torch.from_numpy(wl.synthetic_calibration(n, i, 2048)).to(dev)
        h = torch.zeros((n, n), dtype=torch.float32, device=dev)
        m = torch.zeros(n, dtype=torch.float32, device=dev)
        ops.hessian_accum(x, h, m, 0.0, 2048)
        Hd.append(h)
    lsq = LayerSetQuantizer(cb, scaling_mode="diag", streams=72, batch_k2=True)
    L = len(shapes)
    trace = torch.zeros((L, 8), dtype=torch.int64, device=dev)
    if e2e:
        plan = lsq.host_plan(shapes)
        for i in range(L):
            plan.W[i][...] = Wd[i].cpu().numpy()
            plan.H[i][...] = Hd[i].cpu().numpy()
        plan.run()
        lsq.trace = trace
        plan._graph = None
        plan.run()
        torch.cuda.synchronize()
        trace.zero_()
        plan.run()
    else:
        for _ in range(3):
            lsq(Wd, Hd, keep_outputs=False)
        torch.cuda.synchronize()
        lsq.trace = trace
        strace = None
        if "--sweeptrace" in sys.argv:
            import ctypes
            from sleekit_b200 import _lib
            strace = torch.zeros((128, 16), dtype=torch.int64, device=dev)
            _lib.call("slk_debug_sweep_trace", ctypes.c_void_p(strace.data_ptr()))
        if per_op:
            ops.TRACE = {"buf": torch.zeros(8192, dtype=torch.int64, device=dev), "names": [], "tag": None}
        g, errs, _ = lsq.capture(Wd, Hd)
        optrace = ops.TRACE
        ops.TRACE = None
        g.replay()
        torch.cuda.synchronize()
        trace.zero_()
        g.replay()
    torch.cuda.synchronize()
    t = trace.cpu().numpy().astype(np.float64)
    t0 = t[t > 0].min()
    t = np.where(t > 0, (t - t0) / 1e6, np.nan)     # ms
    print(f"pass length {np.nanmax(t):.3f} ms; groups {lsq.last_groups}")
    names = ["start", "copied", "searched", "prepared", "k2 begin", "k2 end", "sweep begin", "end"]
    for cls, sel in (("n=3072", [i for i, s in enumerate(shapes) if s[1] == 3072]),
                     ("n=768 r=3072", [i for i, s in enumerate(shapes) if s == (3072, 768)]),
                     ("n=768 r=768", [i for i, s in enumerate(shapes) if s == (768, 768)])):
        print(cls)
        for c, nm in enumerate(names):
            col = t[sel, c]
            if np.all(np.isnan(col)):
                continue
            print(f"  {nm:12s} min {np.nanmin(col):7.3f}  median {np.nanmedian(col):7.3f}  max {np.nanmax(col):7.3f}")
    # occupancy of phases per 0.5 ms
    edges = np.arange(0, np.nanmax(t) + 0.5, 0.5)
    print("bucket  search  prepare  wait-k2  sweep")
    for a in edges:
        b = a + 0.5
        def overlap(c0, c1):
            s, e = t[:, c0], t[:, c1]
            return float(np.nansum(np.clip(np.minimum(e, b) - np.maximum(s, a), 0, None)) / 0.5)
        print(f"{a:5.1f}  {overlap(1, 2):6.1f}  {overlap(2, 3):6.1f}  {overlap(3, 6):6.1f}  {overlap(6, 7):6.1f}")
    if not e2e and strace is not None:
        st_ = strace.cpu().numpy()
        st_ = st_[st_[:, 0] > 0]
        print("sweep macro kernel, CTA 0 of the last launch that wrote the trace, cycles per 32-column block:")
        print("  block  wait-barrier  tail+reduce  U_JJ-multiply  leaf||look-ahead  total")
        for b in range(len(st_)):
            r_ = st_[b]
            nxt = st_[b + 1][0] if b + 1 < len(st_) else r_[5]
            print(f"  {b:4d}  {r_[1]-r_[0]:10d}  {r_[2]-r_[1]:10d}  {r_[3]-r_[2]:10d}  {r_[4]-r_[3]:10d}  {nxt-r_[0]:8d}")
    if per_op and not e2e:
        ot = optrace["buf"].cpu().numpy().astype(np.float64)
        names = optrace["names"]
        ot = (ot[: len(names)] - t0) / 1e6
        by_stream = {}
        for k, (nm, tag, sid) in enumerate(names):
            by_stream.setdefault(sid, []).append((ot[k], nm))
        # pick three streams: the one finishing last, a median one, and one long-chain layer
        ends = sorted(by_stream.items(), key=lambda kv: max(x[0] for x in kv[1]))
        for label, (sid, evs) in (("last stream", ends[-1]), ("median stream", ends[len(ends) // 2]), ("first stream", ends[0])):
            print(f"--- {label} ({len(evs)} calls)")
            prev = None
            for tm, nm in sorted(evs):
                print(f"   {tm:8.3f}  (+{0.0 if prev is None else tm - prev:6.3f})  {nm}")
                prev = tm
        # time spent waiting + running per op name (sum of deltas to the previous stamp on the same stream)
        tot = {}
        for sid, evs in by_stream.items():
            evs = sorted(evs)
            for (t_a, _), (t_b, nm) in zip(evs[:-1], evs[1:]):
                tot[nm] = tot.get(nm, 0.0) + (t_b - t_a)
        print("sum over streams of (stamp - previous stamp on the stream), ms:")
        for nm, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            print(f"   {nm:34s} {v:9.3f}")
    k2 = [(t[i, 4], t[i, 5]) for i in range(L) if not np.isnan(t[i, 4])]
    print("factor launches (begin, end):", [(round(a, 3), round(b, 3)) for a, b in sorted(k2)])


if __name__ == "__main__":
    main()
