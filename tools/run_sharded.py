"""Row- and sample-sharded quantization of one big layer (BASELINE config 5) under torchrun.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/run_sharded.py [r n S]

Every rank generates the same synthetic layer (seeded), keeps its row slice of W and its sample
slice of X, runs sleekit_b200.pipeline.quantize_layer_sharded and reports device time (max over
ranks).  With --check, rank 0 also quantizes the first rows on its own from the all-reduced H and
compares (rows never interact, so the sharded rows must match the single-GPU rows)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleekit_b200 import codebook, ops  # noqa: E402
from sleekit_b200 import dist as sdist  # noqa: E402
from sleekit_b200 import workloads as wl  # noqa: E402
from sleekit_b200.pipeline import ShardedLayerQuantizer  # noqa: E402
from sleekit_b200.scaling import quantize_scaled_device, search_scale_device  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("shape", nargs="*", type=int, default=[8192, 28672, 2048])
    ap.add_argument("--codebook", type=int, default=8)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--replicated-factor", action="store_true", help="every rank factors H on its own (round-1 behaviour)")
    args = ap.parse_args()
    r, n, S = (args.shape + [2048])[:3]
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cb = codebook.UniformCodebook(args.codebook, -1, 1)
    a, b = sdist.row_partition(r, world)[rank]
    sa, sb = sdist.row_partition(S, world)[rank]
    g = np.random.default_rng(1000)
    W = torch.from_numpy((0.02 * g.standard_normal((r, n), dtype=np.float32))[a:b].copy()).to(dev)
    X = torch.from_numpy(wl.synthetic_calibration(n, 0, S)[sa:sb].copy()).to(dev)
    slq = ShardedLayerQuantizer(n, cb, group=None, dist_factor=not args.replicated_factor)
    slq(W, X)                       # warm-up (first-touch of the workspaces, NCCL channels)
    times = []
    for _ in range(args.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        q, sc, err, H, mean = slq(W, X)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t))
    slq(W, X, timing=True)
    phases = {k: round(v, 3) for k, v in slq.phases_ms.items()}
    ok = None
    fac_ok = None
    if args.check and slq.dist_factor:
        # the distributed factor against this rank's own single-GPU factor of the same matrix: the tile
        # tasks and their arithmetic are the same, only who runs them differs -> bit-identical
        dampval = ops.damp_value(H, 0.01)
        order = ops.argsort(ops.order_keys(H, dampval, None))
        d = slq._factor_fn(H, order, dampval)
        s1 = ops.chol_factor(H, order, dampval)
        fac_ok = bool(torch.equal(d[0], s1[0]) and torch.equal(d[2], s1[2]) and int(d[3].item()) == 0
                      and torch.equal(torch.tril(d[1][0]), torch.tril(s1[1][0])))
        del d, s1
        t = torch.tensor([1.0 if fac_ok else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        fac_ok = bool(t.item() == 1.0)
    if args.check and rank == 0:
        rows = min(64, b - a)
        sc1 = search_scale_device(W[:rows].contiguous(), cb, H.diagonal().contiguous())
        q1 = quantize_scaled_device(W[:rows].contiguous(), sc1, cb, H)
        ok = bool(torch.equal(sc1, sc[:rows]) and torch.equal(q1, q[:rows]))
        if world == 1:
            # the sharded driver takes the row errors from the sweep's residuals: check against the K6 product
            e6 = float(ops.mean(ops.hweighted_error(W, q, H)))
            ok = ok and abs(float(err) - e6) <= 1e-4 * abs(e6)
    if rank == 0:
        print(json.dumps({"config": f"[{r},{n}] rows sharded over {world} GPU(s), S={S} samples sharded, "
                                    f"{args.codebook}-entry codebook, diag-H scaling + GPTQ + layer error",
                          "n_gpus": world, "ms": times, "weights_per_s": r * n / (min(times) * 1e-3),
                          "layer_error": float(err), "rows_match_single_gpu": ok,
                          "dist_factor": slq.dist_factor, "dist_factor_equals_single_gpu_factor": fac_ok,
                          "phases_ms_rank0": phases,
                          "mem_gb": torch.cuda.max_memory_allocated() / 2**30}), flush=True)
    slq.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
