"""Development check of the tcgen05 GEMM core against an fp64 reference (run under `timeout`)."""

import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleekit_b200 import _lib  # noqa: E402


def P(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def run(epi, A, B, Cmat=None, R=None, A2=None, alpha=1.0, keep=0.0, count=1.0):
    lib = _lib.load()
    M, K = A.shape
    N = B.shape[0]
    nbytes = lib.slk_tc_gemm_ws_bytes(M, N, K)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    if epi == 3:
        tiles = (N + 127) // 128
        out = torch.zeros(M, tiles, device="cuda")
        ldc = tiles
    else:
        out = Cmat
        ldc = Cmat.stride(0)
    _lib.call("slk_tc_gemm_f32", epi, P(A), P(A2), A.stride(0), P(B), B.stride(0), P(out), ldc, P(R),
              R.stride(0) if R is not None else 0, M, N, K, alpha, keep, count, P(ws), nbytes, None,
              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return out


def main():
    torch.manual_seed(0)
    ok = True
    for (M, N, K) in [(128, 128, 32), (128, 128, 64), (256, 128, 96), (768, 768, 768), (200, 136, 100), (3072, 768, 768),
                      (768, 3072, 3072)]:
        A = torch.randn(M, K, device="cuda")
        B = torch.randn(N, K, device="cuda")
        ref = (A.double() @ B.double().T)
        Cm = torch.zeros(M, N, device="cuda")
        out = run(0, A, B, Cm)
        err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
        sg = (A @ B.T)
        err32 = (sg.double() - ref).abs().max().item() / ref.abs().max().item()
        print(f"STORE  {M}x{N}x{K}: rel max err {err:.3e}  (torch fp32 matmul: {err32:.3e})")
        ok &= err < 5e-6
        # accumulate with alpha = -1
        C0 = torch.randn(M, N, device="cuda")
        Cm = C0.clone()
        run(1, A, B, Cm, alpha=-1.0)
        err = (Cm.double() - (C0.double() - ref)).abs().max().item() / ref.abs().max().item()
        print(f"ACCUM  {M}x{N}x{K}: rel max err {err:.3e}")
        ok &= err < 5e-6
        if N == K:
            R = torch.randn(M, N, device="cuda")
            part = run(3, A, B, R=R)
            got = part.sum(dim=1).double()
            want = (ref * R.double()).sum(dim=1)
            err = (got - want).abs().max().item() / want.abs().max().item()
            print(f"ROWDOT {M}x{N}x{K}: rel max err {err:.3e}")
            ok &= err < 2e-5
    # timing
    for (M, N, K) in [(768, 768, 768), (3072, 768, 768), (768, 3072, 3072), (4096, 4096, 4096)]:
        A = torch.randn(M, K, device="cuda")
        B = torch.randn(N, K, device="cuda")
        Cm = torch.zeros(M, N, device="cuda")
        for _ in range(2):
            run(0, A, B, Cm)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run(0, A, B, Cm)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"time {M}x{N}x{K}: {ms * 1e3:.1f} us incl. split, {2 * M * N * K / ms / 1e9:.1f} TFLOP/s algorithmic")
    print("TC_CHECK", "PASS" if ok else "FAIL")


if __name__ == "__main__":
    main()
