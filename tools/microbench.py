"""Per-kernel timing on one GPU with CUDA events (development aid; numbers quoted in DESIGN.md
come from bench.py and ncu, not from here).

    python tools/microbench.py [leaf|sweep|hinv|search|error|xtx|all]
"""

import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sleekit_b200 import _lib, codebook, ops  # noqa: E402
from sleekit_b200 import workloads as wl  # noqa: E402


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    cb = codebook.UniformCodebook(8, -1, 1)
    dev = "cuda"
    if what in ("fastdiv", "all"):
        divs = np.array([2 / 7, 2 / 3, 0.75, 0.1, 1.0, 123.456, 1e-3, 5e-18], dtype=np.float32)
        d = torch.tensor(divs, device=dev)
        bad = torch.zeros(d.numel(), dtype=torch.int64, device=dev)
        _lib.call("slk_selftest_fastdiv_f32", ctypes.c_void_p(d.data_ptr()), d.numel(), ctypes.c_void_p(bad.data_ptr()),
                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        print("fastdiv mismatches", dict(zip(divs.tolist(), bad.tolist())))
    if what == "one":
        r, n = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (768, 768)
        W, H, _ = wl.synthetic_layer(r, n, 0, samples=2048 if n <= 1024 else 4096)
        Wd, Hd = torch.from_numpy(W).to(dev), torch.from_numpy(H).to(dev)
        for _ in range(2):
            damp = ops.damp_value(Hd, 0.01)
            order = ops.argsort(ops.order_keys(Hd, damp, None))
            u64, u32, info = ops.hinv(Hd, order, damp)
            sc = ops.scale_search(Wd, cb, torch.linspace(0.05, 1, 100, device=dev), Hd.diagonal().contiguous())[0]
            Ws = ops.scale_rows(Wd, sc, 0)
            ops.gptq_sweep(Ws, u64, u32, cb)
            ops.hweighted_error(Wd, Ws, Hd)
            torch.cuda.synchronize()
        print("one pass done", r, n)
        return
    if what in ("leaf", "all"):
        for r in (768, 3072):
            n = 32
            q0 = (torch.randn(r, n, device=dev) * 0.3).contiguous()
            u64 = torch.triu(torch.rand(n, n, dtype=torch.float64, device=dev) * 0.1) + torch.eye(n, dtype=torch.float64, device=dev)
            u32 = u64.float()
            e = torch.empty_like(q0)

            def f():
                ops.gptq_sweep(q0.clone(), u64, u32, cb, e=e)

            def g():
                q0.clone()

            print(f"leaf r={r}: {timeit(f) - timeit(g):8.1f} us")
    if what in ("sweep", "hinv", "all"):
        for r, n in ((768, 768), (3072, 768), (768, 3072), (1024, 4096)):
            W, H, _ = wl.synthetic_layer(r, n, 0, samples=2048 if n <= 1024 else 4096)
            Wd, Hd = torch.from_numpy(W).to(dev), torch.from_numpy(H).to(dev)
            damp = ops.damp_value(Hd, 0.01)
            order = ops.argsort(ops.order_keys(Hd, damp, None))
            t_h = timeit(lambda: ops.hinv(Hd, order, damp), reps=5)
            u64, u32, info = ops.hinv(Hd, order, damp)
            sc = ops.scale_search(Wd, cb, torch.linspace(0.05, 1, 100, device=dev), Hd.diagonal().contiguous())[0]
            Ws = ops.scale_rows(Wd, sc, 0)
            t_s = timeit(lambda: ops.gptq_sweep(Ws.clone(), u64, u32, cb), reps=5)
            t_a = timeit(lambda: ops.argsort(ops.order_keys(Hd, damp, None)), reps=5)
            t_c = timeit(lambda: ops.chol_factor(Hd, order, damp), reps=5)
            r32, rt32, ud32, info2 = ops.chol_factor(Hd, order, damp)
            t_r = timeit(lambda: ops.gptq_sweep_r(Ws.clone(), r32, rt32, ud32, cb), reps=5)
            print(f"[{r}x{n}] hinv {t_h:9.1f} us   sweep {t_s:9.1f} us   argsort {t_a:7.1f} us  info={int(info.item())}"
                  f"   chol {t_c:9.1f} us  sweep_r {t_r:9.1f} us  info={int(info2.item())}")
    if what == "sweepr":
        for r, n in ((768, 768), (3072, 768), (768, 3072), (1024, 4096)):
            W, H, _ = wl.synthetic_layer(r, n, 0, samples=2048)
            Wd, Hd = torch.from_numpy(W).to(dev), torch.from_numpy(H).to(dev)
            damp = ops.damp_value(Hd, 0.01)
            order = ops.argsort(ops.order_keys(Hd, damp, None))
            r32, rt32, ud32, info2 = ops.chol_factor(Hd, order, damp)
            sc = ops.scale_search(Wd, cb, torch.linspace(0.05, 1, 100, device=dev), Hd.diagonal().contiguous())[0]
            Ws = ops.scale_rows(Wd, sc, 0)
            sums = torch.empty((r, 2), device=dev)
            dd = torch.empty_like(Ws)
            t_c = timeit(lambda: Ws.clone(), reps=20)
            t_0 = timeit(lambda: ops.gptq_sweep_r(Ws.clone(), r32, rt32, ud32, cb, d=dd), reps=20)
            t_1 = timeit(lambda: ops.gptq_sweep_r(Ws.clone(), r32, rt32, ud32, cb, d=dd, err_sums=sums), reps=20)
            print(f"[{r}x{n}] sweep_r {t_0 - t_c:9.1f} us   with err_sums {t_1 - t_c:9.1f} us")
        return
    if what in ("search", "all"):
        for r, n in ((768, 768), (3072, 768), (768, 3072)):
            W, H, _ = wl.synthetic_layer(r, n, 0, samples=256)
            Wd, Hd = torch.from_numpy(W).to(dev), torch.from_numpy(H).to(dev)
            hd = Hd.diagonal().contiguous()
            f = torch.linspace(0.05, 1, 100, device=dev)
            t = timeit(lambda: ops.scale_search(Wd, cb, f, hd))
            print(f"[{r}x{n}] scale_search {t:9.1f} us  {r * n * 100 / t / 1e3:8.1f} G evals/s  eff {4 * 100 * r * n / t / 1e3:8.1f} GB/s")
    if what in ("error", "all"):
        for r, n in ((768, 768), (3072, 768), (768, 3072)):
            Wd = torch.randn(r, n, device=dev)
            Qd = torch.randn(r, n, device=dev)
            Hd = torch.randn(n, n, device=dev)
            t = timeit(lambda: ops.hweighted_error(Wd, Qd, Hd))
            print(f"[{r}x{n}] hweighted_error {t:9.1f} us  {2 * r * n * n / t / 1e6:8.2f} TFLOP/s")
    if what in ("round", "all"):
        for count in (1 << 24, 1 << 27):
            x = torch.randn(count, device=dev) * 0.6
            t = timeit(lambda: ops.round_to_codebook(x, cb)[0])
            print(f"round value  n={count}: {t:9.1f} us  {8 * count / t / 1e3:8.1f} GB/s (4 B read + 4 B written per weight)")
            t = timeit(lambda: ops.round_to_codebook(x, cb, want_val=False, want_idx=True)[1])
            print(f"round index  n={count}: {t:9.1f} us  {5 * count / t / 1e3:8.1f} GB/s (4 B read + 1 B written per weight)")
            w2 = x.view(-1, 4096)
            s_ = torch.rand(w2.shape[0], device=dev) + 0.5
            t = timeit(lambda: ops.scale_rows(w2, s_, 0))
            print(f"scale rows   n={count}: {t:9.1f} us  {8 * count / t / 1e3:8.1f} GB/s")
    if what in ("xtx", "all"):
        for S, n in ((2048, 768), (2048, 3072), (8192, 4096)):
            X = torch.randn(S, n, device=dev)
            Hd = torch.zeros(n, n, device=dev)
            m = torch.zeros(n, device=dev)
            t = timeit(lambda: ops.hessian_accum(X, Hd, m, 0.5, 2 * S), reps=5)
            print(f"xtx S={S} n={n}: {t:9.1f} us  {2 * S * n * n / t / 1e6:8.2f} TFLOP/s")


if __name__ == "__main__":
    main()
