"""Host-side profile (cProfile) of the reference-facing numpy call sequence of experiments/compare.py:84-95 --
compute_scaling + quantize_with_scaling + quantization_error per layer, pageable numpy in and out -- over the
BASELINE configs[1] layer set (what bench.py reports as e2e_numpy_api).

    python tools/numpy_api_profile.py [--blocks 12] [--top 40]
"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleekit_b200 import codebook, obq, scaling, workloads as wl  # noqa: E402
from sleekit_b200 import _convert as cv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=12)
    ap.add_argument("--top", type=int, default=40)
    args = ap.parse_args()
    shapes = [(768, 768)] * 4 + [(3072, 768), (768, 3072)]
    shapes = shapes * args.blocks
    cb = codebook.UniformCodebook(8, -1, 1)
    Ws, Hs = [], []
    for i, (r, n) in enumerate(shapes):
        W, H, _ = wl.synthetic_layer(r, n, i % 6, samples=2048)
        Ws.append(W)
        Hs.append(H)

    def step():
        for W, H in zip(Ws, Hs):
            sc = scaling.compute_scaling(W, cb, H, mode="diag", grid_size=100)
            q = scaling.quantize_with_scaling(W, sc, cb, H=H, damp=0.01, nb_ls_moves=0)
            obq.quantization_error(W, q, H=H)

    step()
    step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step()
    torch.cuda.synchronize()
    print(f"{len(shapes)} layers: {(time.perf_counter() - t0) * 1e3:.1f} ms per pass; cache hits {cv.CACHE_HITS} misses {cv.CACHE_MISSES}")
    pr = cProfile.Profile()
    pr.enable()
    step()
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(args.top)
    print(s.getvalue())
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(args.top)
    print(s.getvalue())


if __name__ == "__main__":
    main()
