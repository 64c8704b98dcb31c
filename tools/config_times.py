"""Device time of one representative layer of every BASELINE config on one GPU (development aid;
the bench line is configs[1]).  python tools/config_times.py [--cpu] """
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleekit_b200 import codebook, obq, ops, scaling  # noqa: E402
from sleekit_b200 import workloads as wl  # noqa: E402

CASES = [
    # name, r, n, codebook, scaling, bias-corrected H, local-search moves
    ("C1 opt-125m [768,768] 3-bit diag+GPTQ", 768, 768, 8, "diag", False, 0),
    ("C2 opt-125m fc1 [3072,768]", 3072, 768, 8, "diag", False, 0),
    ("C2 opt-125m fc2 [768,3072]", 768, 3072, 8, "diag", False, 0),
    ("C3 opt-350m fc2 [1024,4096] 1.5-bit full-H + H-mm^T", 1024, 4096, 3, "hessian", True, 0),
    ("C3 opt-350m fc1 [4096,1024] 1.5-bit full-H + H-mm^T", 4096, 1024, 3, "hessian", True, 0),
    ("C4 llama-2-7b gate [11008,4096] 2-bit GPTQ + 10 moves", 11008, 4096, 4, "diag", False, 10),
    ("C4 llama-2-7b down [4096,11008] 2-bit GPTQ + 10 moves", 4096, 11008, 4, "diag", False, 10),
    ("C5 llama-3-70b down [8192,28672] 3-bit diag+GPTQ", 8192, 28672, 8, "diag", False, 0),
]


def main():
    dev = "cuda"
    for name, r, n, c, mode, corr, moves in CASES:
        cb = codebook.UniformCodebook(c, -1, 1)
        g = np.random.default_rng(1)
        Wd = torch.from_numpy(0.02 * g.standard_normal((r, n), dtype=np.float32)).to(dev)
        X = torch.from_numpy(wl.synthetic_calibration(n, 7, 2048)).to(dev)
        H = torch.zeros((n, n), dtype=torch.float32, device=dev)
        m = torch.zeros(n, dtype=torch.float32, device=dev)
        ops.hessian_accum(X, H, m, 0.0, 2048)
        Hq = ops.remove_input_bias(H, m) if corr else H
        Hs = Hq if mode == "hessian" else Hq.diagonal().contiguous()

        def one():
            sc = scaling.search_scale_device(Wd, cb, Hs)
            # layer error from the sweep's residuals when no local-search move follows (obq.gptq_device)
            q, (err, _) = scaling.quantize_scaled_device(Wd, sc, cb, Hq, "diag", 0.01, moves, want_err=True)
            return err

        one()
        torch.cuda.synchronize()
        ops.PROFILE = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        err = one()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        ph = {k: round(v[0], 3) for k, v in ops.profile_totals_ms(ops.PROFILE).items()}
        ops.PROFILE = None
        print(json.dumps({"case": name, "ms": round(ms, 3), "Mweights_per_s": round(r * n / ms / 1e3, 1),
                          "layer_error": float(err), "phases_ms": ph}), flush=True)
        del Wd, X, H, Hq, Hs
        torch.cuda.empty_cache()


def obq_case():
    """compute_obq_scaling: 100 full sweeps per layer sharing one factor."""
    dev = "cuda"
    for r, n in ((768, 768), (3072, 768), (768, 3072)):
        cb = codebook.UniformCodebook(8, -1, 1)
        g = np.random.default_rng(1)
        Wd = torch.from_numpy(0.02 * g.standard_normal((r, n), dtype=np.float32)).to(dev)
        X = torch.from_numpy(wl.synthetic_calibration(n, 7, 2048)).to(dev)
        H = torch.zeros((n, n), dtype=torch.float32, device=dev)
        m = torch.zeros(n, dtype=torch.float32, device=dev)
        ops.hessian_accum(X, H, m, 0.0, 2048)
        scaling.obq_scale_device(Wd, cb, H)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sc, info = scaling.obq_scale_device(Wd, cb, H)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"case": f"compute_obq_scaling [{r},{n}] 100 grid points", "ms": round(e0.elapsed_time(e1), 2),
                          "sweeps_per_s": round(100 / (e0.elapsed_time(e1) * 1e-3), 1)}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "obq":
        obq_case()
    else:
        main()
