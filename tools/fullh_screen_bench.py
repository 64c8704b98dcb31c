"""Full-H scale search (compute_min_mse_scaling with a 2-D H, scaling.py:98-134) on the BASELINE configs[2]
layer shapes: every grid point evaluated by the 3xTF32 product (fullh_topk = 0) against the screened search
(one bf16 or TF32 pass over all points + exact evaluation of the 4 / 8 / 16 best-ranked ones), with both tile
widths of the screening product.  Prints ms per call (CUDA events, after warm-up) and whether scales and errors are
bit-identical to the unscreened result.

    python tools/fullh_screen_bench.py [--reps 3]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleekit_b200 import codebook, ops, workloads as wl  # noqa: E402
from sleekit_b200.scaling import _factors  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--shape", default="", help="rows,cols: only this shape")
    ap.add_argument("--topk", default="", help="only this candidate count")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cb = codebook.UniformCodebook(3, -1, 1)
    f = _factors(0.05, 1.0, 100, dev)
    out = []
    shapes = [tuple(int(v) for v in args.shape.split(","))] if args.shape else [(1024, 1024), (4096, 1024), (1024, 4096)]
    # (candidates, screening tile width, bf16 operands, CTAs per SM, compacted candidate list)
    variants = [(0, 256, 1, 2, 1), (8, 256, 1, 2, 1), (8, 256, 1, 2, 0), (8, 256, 1, 1, 0), (8, 128, 1, 2, 0),
                (4, 256, 1, 2, 0), (16, 256, 1, 2, 1), (16, 256, 1, 2, 0), (8, 256, 0, 2, 0)]
    if args.topk:
        variants = [(int(args.topk), 256, 1, 2, 1)]
    for r, n in shapes:
        W, H, _ = wl.synthetic_layer(r, n, 31, samples=2048)
        Wd, Hd = torch.from_numpy(W).to(dev), torch.from_numpy(H).to(dev)
        base = None
        for topk, bn, bf16, ctas, compact in variants:
            ops.set_option("fullh_ctas", ctas)
            ops.set_option("fullh_compact", compact)
            ops.set_option("fullh_topk", topk)
            ops.set_option("fullh_bn", bn)
            ops.set_option("fullh_bf16", bf16)
            for _ in range(2):
                sc, err = ops.scale_search_fullh(Wd, cb, f, Hd, want_err=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                sc, err = ops.scale_search_fullh(Wd, cb, f, Hd, want_err=True)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            if base is None:
                base = (sc.clone(), err.clone())
            rec = {"rows": r, "cols": n, "grid_points": 100, "fullh_topk": topk, "screen_tile_n": bn if topk else None, "screen_ctas_per_sm": ctas if topk else None,
                   "candidates": ("within 2^-5 of the best-ranked, compacted" if compact else "fixed per row") if topk else None,
                   "screen_operands": ("bf16" if bf16 else "tf32") if topk else None,
                   "ms": round(ms, 3), "algorithmic_tflops": round(2.0 * 100 * r * n * n / ms / 1e9, 1),
                   "scales_identical": bool(torch.equal(sc, base[0])), "errors_identical": bool(torch.equal(err, base[1]))}
            out.append(rec)
            print(json.dumps(rec), flush=True)
    ops.set_option("fullh_topk", 8)
    ops.set_option("fullh_bn", 256)
    ops.set_option("fullh_bf16", 1)
    ops.set_option("fullh_ctas", 2)
    ops.set_option("fullh_compact", 1)
    return out


if __name__ == "__main__":
    main()
