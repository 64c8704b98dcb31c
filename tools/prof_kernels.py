"""One kernel scenario per invocation, as an ncu target (profiles/README.md lists the commands).

    python tools/prof_kernels.py <case> [reps]

cases:  gemm4096   tcgen05 3xTF32 GEMM 4096^3 (slk_tc_gemm_f32, incl. the split passes)
        k1         K1 X^T X, S = 8192, n = 4096 (statistics.py:76-87)
        k6         K6 ((W-Q) H (W-Q)^T row sums, [11008, 4096] (obq.py:89-95)
        c5gemm     the K = 2048 super-block push of the n = 28672 sweep: M = 8192, N = 8192, K = 2048
        k4         K4 rounding of 2^27 values, value and index forms (codebook.py:43-65)
        search     K5 diag-H scale search, [3072, 768] and [768, 3072] (scaling.py:98-134)
        chol12     K2 batched: twelve n = 3072 factorisations in one launch (obq.py:38-55)
        chol1      K2 single n = 3072 / n = 768 factorisation
        sweep      K3 sweep of a [768, 3072] and a [3072, 768] layer (obq.py:106-137)
        fullh      full-H scale search, [1024, 4096], 3-entry codebook (scaling.py:84-95, 2-D H)
        ls         K7 local search, [11008, 4096], 10 moves (obq.py:234-358)
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleekit_b200 import _lib, codebook, ops, scaling  # noqa: E402
from sleekit_b200 import workloads as wl  # noqa: E402


def P(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def layer(r, n, lid=0, S=2048):
    W = torch.from_numpy(wl.synthetic_weight(r, n, lid)).cuda()
    X = torch.from_numpy(wl.synthetic_calibration(n, lid, S)).cuda()
    H = torch.zeros((n, n), dtype=torch.float32, device="cuda")
    m = torch.zeros(n, dtype=torch.float32, device="cuda")
    ops.hessian_accum(X, H, m, 0.0, S)
    return W, H, m


def tc_gemm(M, N, K, reps):
    lib = _lib.load()
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    out = torch.zeros(M, N, device="cuda")
    nbytes = lib.slk_tc_gemm_ws_bytes(M, N, K)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for _ in range(reps):
        _lib.call("slk_tc_gemm_f32", 0, P(A), None, K, P(B), K, P(out), N, None, 0, M, N, K, 1.0, 0.0, 1.0, P(ws), nbytes,
                  None, C.c_void_p(torch.cuda.current_stream().cuda_stream))


def main():
    case = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cb8 = codebook.UniformCodebook(8, -1, 1)
    if case == "gemm4096":
        tc_gemm(4096, 4096, 4096, reps)
    elif case == "c5gemm":
        tc_gemm(8192, 8192, 2048, reps)
    elif case == "k1":
        x = torch.randn(8192, 4096, device="cuda")
        h = torch.zeros(4096, 4096, device="cuda")
        m = torch.zeros(4096, device="cuda")
        for k in range(reps):
            ops.hessian_accum(x, h, m, k / (k + 1.0), (k + 1.0) * 8192)
    elif case == "k6":
        W, H, _ = layer(11008, 4096)
        Q = ops.round_to_codebook(W * 30.0, cb8)[0] / 30.0
        for _ in range(reps):
            ops.hweighted_error(W, Q, H)
    elif case == "k4":
        x = torch.randn(1 << 27, device="cuda") * 0.6
        for _ in range(reps):
            ops.round_to_codebook(x, cb8, want_val=True, want_idx=False)
            ops.round_to_codebook(x, cb8, want_val=False, want_idx=True)
    elif case == "search":
        f = torch.linspace(0.05, 1.0, 100, device="cuda")
        for r, n in ((3072, 768), (768, 3072)):
            W, H, _ = layer(r, n)
            hd = H.diagonal().contiguous()
            for _ in range(reps):
                ops.scale_search(W, cb8, f, hd)
    elif case in ("chol12", "chol1"):
        sizes = [3072] * 12 if case == "chol12" else [3072, 768]
        hs, orders, damps = [], [], []
        for k, n in enumerate(sizes):
            _, H, _ = layer(8, n, k)
            dv = ops.damp_value(H, 0.01)
            hs.append(H)
            damps.append(dv)
            orders.append(ops.argsort(ops.order_keys(H, dv, None)))
        for _ in range(reps):
            if case == "chol12":
                ops.chol_factor_batched(hs, orders, damps)
            else:
                for h, o, d in zip(hs, orders, damps):
                    ops.chol_factor(h, o, d)
    elif case == "sweep":
        for r, n in ((768, 3072), (3072, 768)):
            W, H, _ = layer(r, n)
            sc = scaling.search_scale_device(W, cb8, H.diagonal().contiguous())
            for _ in range(reps):
                scaling.quantize_scaled_device(W, sc, cb8, H, "diag", 0.01, 0, want_err=True)
    elif case == "fullh":
        cb3 = codebook.UniformCodebook(3, -1, 1)
        W, H, m = layer(1024, 4096)
        Hc = ops.remove_input_bias(H, m)
        for _ in range(reps):
            scaling.search_scale_device(W, cb3, Hc)
    elif case == "ls":
        cb4 = codebook.UniformCodebook(4, -1, 1)
        W, H, _ = layer(11008, 4096)
        sc = scaling.search_scale_device(W, cb4, H.diagonal().contiguous())
        Ws = ops.scale_rows(W, sc, 0)
        Q0 = ops.round_to_codebook(Ws, cb4)[0]
        for _ in range(reps):
            ops.local_search(Ws, Q0.clone(), H, cb4, 10)
    else:
        raise SystemExit(f"unknown case {case}")
    torch.cuda.synchronize()
    print("ok", case)


if __name__ == "__main__":
    main()
