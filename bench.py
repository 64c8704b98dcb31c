#!/usr/bin/env python
"""bench.py -- GPTQ weights quantized per second (BASELINE.json's metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4|c5|xtx]

Default workload = BASELINE.json configs[1] ("c2"): all 72 OPT-125M-shaped linear layers, 3-bit
uniform codebook, diag-H scale-grid search (100 points) + GPTQ (diag ordering, 1 % damp) + layer
error.  One "step" = one pass of the hot path over the whole layer set, with W and H resident in HBM
(`value`), and again from / to HOST buffers (`e2e`).  N > 1 (torchrun): every rank quantizes its own
layer set (layers are independent; no data-path collective) -> weak scaling.  The same line carries
`sharded_c5`: BASELINE configs[4], ONE Llama-3-70B-shaped [8192, 28672] layer with its rows and
calibration samples sharded over the N ranks -- K1 on the local samples, NCCL all-reduce of the
packed statistics, the fp64 factor distributed over the GPUs through NVLink peer stores, row-local
scale search and sweep -- strong scaling, per-phase device times, all inside its timed region.

Other configs (one JSON line each, same keys): c3 = OPT-350M-shaped layers, 3-entry codebook, full-H
scale search on the bias-corrected Hessian; c4 = Llama-2-7B MLP layers, 4-entry codebook, GPTQ +
10 best-first local-search moves; c5 = the sharded layer as the line's own workload; xtx = the
calibration product X^T X alone (TFLOP/s).

`--impl reference` times the reference's own CPU implementation (the unmodified package installed
in baseline/_ref when present -- kind "reference" -- else the numpy restatement in oracle/ -- kind
"port") on the host cores, on a bounded sample of the same workload.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# independent layers run as parallel graph branches / streams: give them distinct hardware queues
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

METRIC = "gptq_weights_quantized_per_s"
UNIT = "weights/s"
GRID = 100
DAMP = 0.01
SAMPLES = 2048

CONFIGS = {
    # model (workloads.layer_shapes), layers per block, default blocks, codebook, scaling mode, H - m m^T, moves
    "c2": dict(model="opt-125m", per_block=6, blocks=12, codebook=8, scaling="diag", bias=False, moves=0,
               name="BASELINE configs[1]"),
    "c3": dict(model="opt-350m", per_block=6, blocks=2, codebook=3, scaling="hessian", bias=True, moves=0,
               name="BASELINE configs[2]"),
    "c4": dict(model="llama2-7b-mlp", per_block=3, blocks=1, codebook=4, scaling="diag", bias=False, moves=10,
               name="BASELINE configs[3]"),
}
C5_ROWS, C5_COLS = 8192, 28672


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5", "xtx"])
    ap.add_argument("--blocks", type=int, default=0, help="transformer blocks of the config's model (0 = the config's default)")
    ap.add_argument("--layers", type=int, default=0, help="debug: only the first L layers")
    ap.add_argument("--cpu-row-div", type=int, default=0,
                    help="CPU arm: 1/div of each layer's rows for the row-linear phases (0 = config default; 1 = every row)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="skip the sharded_c5 measurement of the default line")
    ap.add_argument("--sharded-rows", type=int, default=C5_ROWS)
    ap.add_argument("--sharded-cols", type=int, default=C5_COLS)
    ap.add_argument("--replicated-factor", action="store_true", help="sharded layer: every rank factors H on its own")
    ap.add_argument("--only", default="", choices=["", "big", "small"], help="debug: only layers with n >= 2048 / n < 2048")
    ap.add_argument("--streams", type=int, default=72, help="CUDA streams the independent layers are spread over")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--model-order", action="store_true", help="issue layers in model order instead of longest chains first")
    return ap.parse_args()


def config_shapes(args):
    from sleekit_b200 import workloads as wl

    cfg = CONFIGS[args.config]
    shapes = wl.layer_shapes(cfg["model"])
    blocks = args.blocks or cfg["blocks"]
    shapes = shapes[: blocks * cfg["per_block"]]
    if args.layers:
        shapes = shapes[: args.layers]
    if args.only:
        shapes = [sh for sh in shapes if (sh[1] >= 2048) == (args.only == "big")]
    return cfg, shapes


def workload_name(cfg, nlayers):
    bits = {8: "3-bit", 3: "1.5-bit", 4: "2-bit"}.get(cfg["codebook"], "")
    search = {"diag": "diag-H scale search", "hessian": "full-H scale search", "mse": "MSE scale search"}[cfg["scaling"]]
    return (f"{cfg['name']}: {cfg['model']}: {nlayers} linear layers, {cfg['codebook']}-entry uniform codebook ({bits}), "
            f"{search} ({GRID} pts){' on H - m m^T' if cfg['bias'] else ''} + GPTQ (diag order, {DAMP:g} damp)"
            f"{' + %d best-first local-search moves' % cfg['moves'] if cfg['moves'] else ''} + layer error, "
            f"S={SAMPLES} synthetic calibration rows")


# ---------------------------------------------------------------------------
# CPU arm: the reference itself (baseline/_ref) or its numpy restatement (oracle/) on a bounded sample
# ---------------------------------------------------------------------------


class CpuArm:
    """The reference's per-layer call sequence (experiments/compare.py:84-95, statistics.py:160-186) on the
    host: [remove_input_bias] -> compute_scaling -> quantize_with_scaling (GPTQ, local search) ->
    quantization_error."""

    def __init__(self):
        ref = os.path.join(ROOT, "baseline", "_ref")
        self.kind = "port"
        if os.path.isdir(os.path.join(ref, "sleekit")) and "sleekit" not in sys.modules:
            sys.path.insert(0, ref)
            try:
                import sleekit.codebook as rcb
                import sleekit.obq as robq
                import sleekit.scaling as rsc

                if os.path.realpath(rcb.__file__).startswith(os.path.realpath(ref)):
                    self.kind, self.rcb, self.robq, self.rsc = "reference", rcb, robq, rsc
            except Exception:
                pass
            finally:
                sys.path.remove(ref)
        if self.kind == "port":
            from oracle import sleekit_oracle as orc

            self.orc = orc

    def describe(self):
        return ("the unmodified reference package (baseline/_ref, pip --target install of /root/reference)"
                if self.kind == "reference" else "oracle/sleekit_oracle.py, the numpy restatement of the reference")

    def codes(self, q, sc, c):
        """codebook indices of de-scaled weights q under row scales sc (codebook.py:43-54)."""
        if self.kind == "reference":
            return self.rcb.UniformCodebook(c, -1, 1).quantize_index(q / sc[:, None])
        return self.orc.UniformGrid(c, -1, 1).index(self.orc.divide_rows(q, sc, 0))

    def layer(self, W, H, mean, cfg, row_div=1):
        """Returns dict(est = seconds for the whole layer, measured = seconds spent, q, sc, err, rows)."""
        r, n = W.shape
        rs = max(1, r // row_div)
        Ws = np.ascontiguousarray(W[:rs])
        c = cfg["codebook"]
        t0 = time.perf_counter()
        if self.kind == "reference":
            cb = self.rcb.UniformCodebook(c, -1, 1)
            Hq = self.robq.remove_input_bias(H, mean) if cfg["bias"] else H
            t1 = time.perf_counter()
            sc = self.rsc.compute_scaling(Ws, cb, H=Hq, mode=cfg["scaling"], grid_size=GRID)
            t2 = time.perf_counter()
            q = self.rsc.quantize_with_scaling(Ws, sc, cb, H=Hq, act_order="diag", damp=DAMP, nb_ls_moves=cfg["moves"])
            t3 = time.perf_counter()
            err = self.robq.quantization_error(Ws, q, H=Hq)
            t4 = time.perf_counter()
            factor = 0.0
            if rs < r:   # the fp64 factor does not depend on the rows: time it alone to extrapolate the rest
                Hd = Hq.astype(np.float32) + DAMP * Hq.diagonal().mean() * np.eye(n)
                tf = time.perf_counter()
                self.robq.compute_hessian_chol(Hd)
                factor = time.perf_counter() - tf
        else:
            orc = self.orc
            grid = orc.UniformGrid(c, -1, 1)
            Hq = orc.strip_input_bias(H, mean) if cfg["bias"] else H
            t1 = time.perf_counter()
            sc = orc.choose_scale(Ws, grid, Hq, mode=cfg["scaling"], points=GRID)
            t2 = time.perf_counter()
            q = orc.quantize_scaled(Ws, sc, grid, H=Hq, rule="diag", damp=DAMP, ls_moves=cfg["moves"])
            t3 = time.perf_counter()
            err = orc.mean_error(Ws, q, Hq)
            t4 = time.perf_counter()
            factor = 0.0
            if rs < r:
                Hd = Hq.astype(np.float32) + DAMP * Hq.diagonal().mean() * np.eye(n)
                tf = time.perf_counter()
                orc.inverse_upper_factor(Hd)
                factor = time.perf_counter() - tf
        total = t4 - t0
        fixed = (t1 - t0) + min(factor, t3 - t2)            # bias removal + factor: independent of the rows
        est = fixed + (total - fixed) * (r / rs)
        return dict(est=est, measured=total + factor, q=q, sc=sc, err=float(err), rows=rs)


def cpu_sample(arm, cfg, shapes, inputs, row_div):
    """One sample step: each distinct layer shape of the first block once.  inputs(lid) -> (W, H, mean).
    Returns (extrapolated whole-block seconds, weights of the block, measured seconds, per-layer results)."""
    block = shapes[: cfg["per_block"]]
    cache, est_total, measured, results = {}, 0.0, 0.0, {}
    for lid, (r, n) in enumerate(block):
        if (r, n) not in cache:
            W, H, m = inputs(lid)
            res = arm.layer(W, H, m, cfg, row_div)
            cache[(r, n)] = res["est"]
            measured += res["measured"]
            results[lid] = res
        est_total += cache[(r, n)]
    return est_total, sum(r * n for r, n in block), measured, results


def default_row_div(config):
    # c2 runs every row of its three sampled layers (~15 s of CPU); the larger configs a row subset
    return {"c2": 1, "c3": 32, "c4": 256}.get(config, 1)


def sample_text(cfg, shapes, row_div, arm):
    block = shapes[: cfg["per_block"]]
    return (f"per step: block 0 of the workload ({len(block)} layers, {sum(r * n for r, n in block)} weights): one layer "
            f"of each distinct shape ({sorted(set(block))}) is timed"
            + (" in full (every row)" if row_div == 1 else
               f" on the first 1/{row_div} of its rows for the row-linear phases (scale search, sweep, local search, "
               f"error: scaled back), bias removal and the fp64 factor in full")
            + f", repeated shapes counted by multiplicity; {arm.describe()}, numpy/OpenBLAS on all host cores")


_CPU_INPUTS = {}


def synthetic_inputs(shapes):
    from sleekit_b200 import workloads as wl

    def get(lid):
        r, n = shapes[lid]
        if lid not in _CPU_INPUTS:
            _CPU_INPUTS[lid] = wl.synthetic_layer(r, n, lid, samples=SAMPLES)
        return _CPU_INPUTS[lid]

    return get


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config in ("c5", "xtx"):
        print(json.dumps({"impl": "reference", "unavailable": f"--config {args.config} has no bounded CPU sample "
                          "(the n = 28672 fp64 factor alone takes ~13 min on the host); see cpu_baseline of c2"}))
        return
    cfg, shapes = config_shapes(args)
    arm = CpuArm()
    row_div = args.cpu_row_div or default_row_div(args.config)
    cores = os.cpu_count()
    inputs = synthetic_inputs(shapes)
    for _ in range(args.warmup):
        cpu_sample(arm, cfg, shapes, inputs, row_div)
    ests = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        est, weights, _, _ = cpu_sample(arm, cfg, shapes, inputs, row_div)
        ests.append(est)
    wall = time.perf_counter() - t0
    value = weights / (sum(ests) / len(ests))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg, len(shapes))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": arm.kind,
                         "sample": sample_text(cfg, shapes, row_div, arm)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------


class ClockSampler:
    """SM clock and throttle reasons sampled WHILE the timed region runs: an NVML polling thread (5 ms period;
    nvidia_ml_py), with `nvidia-smi -lms` as fallback.  stop(window) prefers the samples inside the timed window."""

    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20))

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []            # (time.perf_counter(), sm MHz, max MHz, reasons bitmask)
        self.thread = None
        self.stop_flag = False
        self.proc = None
        self.path = None

    def _poll(self, handle, nv):
        mx = nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM)
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                self.samples.append((time.perf_counter(), float(sm), float(mx), int(rs)))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import threading

            import pynvml as nv

            nv.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.gpu
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                idx = int(vis.split(",")[self.gpu])
            handle = nv.nvmlDeviceGetHandleByIndex(idx)
            self.thread = threading.Thread(target=self._poll, args=(handle, nv), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, window=None):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": None}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            rows = self.samples
            inside = [r for r in rows if window and window[0] <= r[0] <= window[1]]
            use = inside or rows
            if use:
                bits = 0
                for r in use:
                    bits |= r[3]
                out.update(sm_mhz=statistics.median(r[1] for r in use), sm_max_mhz=max(r[2] for r in use),
                           reasons=sorted(n for n, b in self.REASONS if bits & b), samples=len(use),
                           samples_inside_timed_region=len(inside), source="NVML, 5 ms period")
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if val.lower() == "active":
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       source="nvidia-smi -lms 50 (started before the warm-up; whole loaded window)")
        return out


# ---------------------------------------------------------------------------
# helpers of our arm
# ---------------------------------------------------------------------------


class Dist:
    def __init__(self):
        import torch

        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        import torch

        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """steps calls of fn between barrier + synchronize; returns (device ms, wall ms), max over ranks."""
        import torch

        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        self.last_window = (t0, time.perf_counter())
        wall = 1e3 * (time.perf_counter() - t0)
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms, wall], dtype=torch.float64, device=self.dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1])
        return ms, wall


def measured_peaks(dev):
    """MEASURED_PEAKS.json (driver-written) + the two peaks it lacks, measured here with cuBLAS: dense TF32
    (torch.matmul with TF32 allowed) and FP64 GEMM."""
    import torch

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    out = {"hbm_gbs": float(peaks.get("hbm_gbs", 6650.0)),
           "hbm_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
           "bf16_tflops": float(peaks.get("bf16_tflops", 1590.0))}

    def gemm_rate(dtype, n, reps, tf32=False):
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            a = torch.randn(n, n, dtype=dtype, device=dev)
            b = torch.randn(n, n, dtype=dtype, device=dev)
            torch.matmul(a, b)
            torch.cuda.synchronize()
            best = 0.0
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    torch.matmul(a, b)
                e1.record()
                torch.cuda.synchronize()
                best = max(best, reps * 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
            return best
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old

    out["fp64_tflops"] = gemm_rate(torch.float64, 4096, 3)
    out["tf32_tflops"] = gemm_rate(torch.float32, 8192, 5, tf32=True)
    out["fp64_source"] = "cuBLAS fp64 GEMM 4096^3 measured in this run (MEASURED_PEAKS.json has no fp64 figure)"
    out["tf32_source"] = "cuBLAS TF32 GEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no TF32 figure)"
    return out


def ncu_traffic():
    for name in ("r2_ncu_traffic.json", "r1_ncu_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))
        except Exception:
            continue
    return {}


def phase_rooflines(phases, profile, shapes, groups, cfg, peaks):
    """One roofline entry per timed operation of the serial pass (SURVEY 8d work models); the dominant one
    (largest share of the serial device time) becomes the line's `roofline`."""
    total_ms = sum(v[0] for v in phases.values()) or 1.0
    traffic = ncu_traffic()
    rn2 = sum(float(r) * n * n for r, n in shapes)
    rn = sum(float(r) * n for r, n in shapes)
    n3 = sum(float(n) ** 3 for r, n in shapes)
    out = {}
    for name, (ms, calls) in phases.items():
        if ms <= 0:
            continue
        ent = {"kernel": name, "launches_timed": calls, "avg_launch_ms": ms / calls, "share_of_serial_device_time": ms / total_ms,
               "traffic": None}
        if name == "chol_factor":
            work = n3 / 3.0
            ent.update(bound="tensor", unit="TFLOP/s", achieved=work / (ms * 1e-3) / 1e12, peak=peaks["fp64_tflops"],
                       peak_source=peaks["fp64_source"],
                       note=("K2: fp64 damp + permute + Cholesky factor, n^3/3 flop per matrix (factor only: the sweep's R "
                             "form needs no inverse), tile-task kernel on the FP64 tensor path (DMMA); equal-sized "
                             "matrices are factored by ONE launch whose tile tasks share a ticket queue; achieved = "
                             "algorithmic flop of all timed launches / their total time"))
            if groups:
                per = [s_.elapsed_time(e_) for s_, e_ in profile.get(name, [])]
                if len(per) == len(groups):
                    ent["by_launch"] = [{"n": n_, "matrices": b_, "ms": round(t_, 4),
                                         "achieved": round(b_ * n_ ** 3 / 3.0 / (t_ * 1e-3) / 1e12, 3)}
                                        for (n_, b_), t_ in zip(groups, per)]
            tot_b, ok = 0.0, True
            for r_, n_ in shapes:
                key = f"chol_factor:n={n_}"
                if key in traffic:
                    tot_b += traffic[key]["dram_bytes_per_launch"]
                else:
                    ok = False
            if ok and calls:
                ent["traffic"] = tot_b / calls
        elif name == "gptq_sweep":
            ent.update(bound="tensor", unit="TFLOP/s", achieved=rn2 / (ms * 1e-3) / 1e12, peak=peaks["tf32_tflops"],
                       peak_source=peaks["tf32_source"],
                       note=("K3: algorithmic r*n^2 flop per layer (the lazy-batch trailing updates); the tensor-core part "
                             "issues 3x that as TF32 MMAs (fp32-faithful split), the in-block part is a dependent chain"))
        elif name in ("hweighted_error", "local_search"):
            ent.update(bound="tensor", unit="TFLOP/s", achieved=2.0 * rn2 / (ms * 1e-3) / 1e12, peak=peaks["tf32_tflops"],
                       peak_source=peaks["tf32_source"], note="2*r*n^2 algorithmic flop per layer ((W-Q) H product)")
        elif name == "scale_search_fullh":
            ent.update(bound="tensor", unit="TFLOP/s", achieved=2.0 * GRID * rn2 / (ms * 1e-3) / 1e12,
                       peak=peaks["tf32_tflops"], peak_source=peaks["tf32_source"],
                       note=("full-H scale search: achieved = the REFERENCE's work, 2*G*r*n^2 flop per layer (scaling.py:84-95 "
                             "per grid point), over the measured time.  The kernel path issues less: one bf16 tcgen05 pass over "
                             "all G grid points (ranking only) plus the fp32-faithful 3xTF32 product for the 8 best-ranked "
                             "points of each row -- results bit-identical to evaluating every point "
                             "(test_full_h_search_screening_is_exact).  Issued: (G + 24) / G x the algorithmic flop, G/(G+24) "
                             "of it on the bf16 pipe (ncu, profiles/r2_ncu_raw_fullh_screen.csv: tensor pipe 47.7 % active "
                             "in the bf16 pass, 59.8 % in the exact pass); peak = the TF32 rate the unscreened search ran at"))
        elif name == "scale_search":
            ent.update(bound="hbm", unit="GB/s", achieved=4.0 * GRID * rn / (ms * 1e-3) / 1e9, peak=peaks["hbm_gbs"],
                       peak_source=peaks["hbm_source"],
                       note=("effective GB/s: algorithmic bytes = 4*G bytes per weight (the reference's G passes over W, "
                             "scaling.py:127-133); the fused kernel reads W twice (8 B/weight of real traffic) and is bound "
                             "by instruction issue"))
        else:
            ent.update(bound="hbm", unit="GB/s", achieved=8.0 * rn / (ms * 1e-3) / 1e9, peak=peaks["hbm_gbs"],
                       peak_source=peaks["hbm_source"], note="algorithmic bytes = one read + one write of W per layer")
        ent["frac"] = ent["achieved"] / ent["peak"]
        out[name] = ent
    return out


def xtx_measure(dev, cases, reps=5):
    """K1 (statistics.py:76-87) alone: algorithmic 2*S*n^2 flop per product; returns one entry per case."""
    import torch
    from sleekit_b200 import ops

    out = []
    for S, n in cases:
        x = torch.randn(S, n, dtype=torch.float32, device=dev)
        h = torch.zeros((n, n), dtype=torch.float32, device=dev)
        m = torch.zeros(n, dtype=torch.float32, device=dev)
        ops.hessian_accum(x, h, m, 0.0, S)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(reps):
            ops.hessian_accum(x, h, m, k / (k + 1.0), (k + 1.0) * S)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        # issued TF32 MMA flop: upper tiles only, 3 MMAs per product (fp32-faithful split)
        t = (n + 127) // 128
        issued = 3.0 * 2.0 * S * 128 * 128 * (t * (t + 1) / 2)
        out.append({"S": S, "n": n, "ms": ms, "tflops_algorithmic": 2.0 * S * n * n / (ms * 1e-3) / 1e12,
                    "tflops_issued_tf32": issued / (ms * 1e-3) / 1e12})
        del x, h, m
    return out


def sharded_layer(args, D, steps, warmup, peaks):
    """BASELINE configs[4]: one [rows, cols] layer, rows of W and calibration samples sharded over the ranks."""
    import torch
    from sleekit_b200 import codebook
    from sleekit_b200 import dist as sdist
    from sleekit_b200.pipeline import ShardedLayerQuantizer

    r, n = args.sharded_rows, args.sharded_cols
    dev = D.dev
    a, b = sdist.row_partition(r, D.world)[D.rank]
    sa, sb = sdist.row_partition(SAMPLES, D.world)[D.rank]
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + D.rank)
    W = 0.02 * torch.randn((b - a, n), generator=g, dtype=torch.float32, device=dev)
    # calibration rows of SURVEY 8d's recipe, generated on the device: a rank-64 correlated part plus noise,
    # log-normal per-channel scales, non-zero mean; the channel scales and the mixing matrix are common to all ranks
    gc = torch.Generator(device=dev)
    gc.manual_seed(2000)
    mix = torch.randn((64, n), generator=gc, dtype=torch.float32, device=dev)
    chan = torch.exp(torch.randn(n, generator=gc, dtype=torch.float32, device=dev))
    X = torch.randn((sb - sa, 64), generator=g, dtype=torch.float32, device=dev) @ mix
    X += 0.3 * torch.randn((sb - sa, n), generator=g, dtype=torch.float32, device=dev)
    X = (X * chan + 0.5).contiguous()
    del mix
    cb = codebook.UniformCodebook(8, -1, 1)
    slq = ShardedLayerQuantizer(n, cb, scaling_mode="diag", act_order="diag", damp=DAMP, dist_factor=not args.replicated_factor)
    out = {}

    def step():
        out["res"] = slq(W, X)

    for _ in range(max(1, warmup)):
        step()
    sampler = ClockSampler(D.local)
    sampler.start()
    ms, wall = D.timed(step, steps)
    clocks = sampler.stop(D.last_window)
    slq(W, X, timing=True)
    phases = dict(slq.phases_ms)
    if D.world > 1:
        keys = sorted(phases)
        t = torch.tensor([phases[k] for k in keys], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        phases = {k: float(v) for k, v in zip(keys, t)}
    q, sc, err, H, mean = out["res"]
    # parity: the first rows of rank 0 quantized on their own from the all-reduced H (rows never interact)
    from sleekit_b200.scaling import quantize_scaled_device, search_scale_device

    rows = min(32, b - a)
    sc1 = search_scale_device(W[:rows].contiguous(), cb, H.diagonal().contiguous())
    q1 = quantize_scaled_device(W[:rows].contiguous(), sc1, cb, H)
    same = torch.tensor([1.0 if (torch.equal(sc1, sc[:rows]) and torch.equal(q1, q[:rows])) else 0.0], device=dev)
    if D.world > 1:
        torch.distributed.all_reduce(same, op=torch.distributed.ReduceOp.MIN)
    ms_layer = ms / steps
    from sleekit_b200 import ops as _ops

    L = int(_ops.sym_packed_len(n)) if D.world > 1 else 0
    res = {
        "workload": (f"BASELINE configs[4]: one Llama-3-70B-shaped [{r},{n}] layer, 8-entry codebook, diag-H scale search + GPTQ + "
                     f"layer error; rows of W and the S={SAMPLES} calibration rows sharded over {D.world} GPU(s); the timed "
                     f"region holds K1 on the local samples, the NCCL all-reduce of the packed statistics, the fp64 factor"
                     f"{' distributed over the GPUs (tile rows cyclic, tiles pushed through NVLink peer stores)' if slq.dist_factor else ' (replicated on every rank)'}"
                     f", the row-local scale search and sweep, and the error all-reduce"),
        "n_gpus": D.world, "scaling": "strong", "steps": steps, "ms_per_layer": ms_layer,
        "value": r * n / (ms_layer * 1e-3), "unit": UNIT, "layer_error": float(err),
        "phases_ms_max_over_ranks": {k: round(v, 3) for k, v in phases.items()},
        "dist_factor": slq.dist_factor, "allreduce_bytes": 4 * (L + n) if D.world > 1 else 0,
        "rows_equal_single_gpu_rows": bool(same.item() == 1.0), "clocks": clocks,
        "factor_tflops_fp64": (n ** 3 / 3.0) / (phases.get("factor", 0) * 1e-3) / 1e12 if phases.get("factor") else None,
        "factor_frac_of_fp64_peak_x_gpus": ((n ** 3 / 3.0) / (phases["factor"] * 1e-3) / 1e12 / (peaks["fp64_tflops"] * D.world)
                                            if phases.get("factor") else None),
        "mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
    }
    slq.close()
    del W, X, q, sc, H, mean, out
    torch.cuda.empty_cache()
    return res


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------


def run_ours(args):
    import torch

    D = Dist()
    rank, world, dev = D.rank, D.world, D.dev
    from sleekit_b200 import _convert as cv
    from sleekit_b200 import codebook, obq, ops, scaling
    from sleekit_b200 import workloads as wl

    peaks = measured_peaks(dev)

    if args.config == "xtx":
        cases = [(2048, 768), (2048, 3072), (2048, 4096), (8192, 4096), (8192, 8192), (2048, 28672)]
        res = xtx_measure(dev, cases)
        if rank == 0:
            best = max(res, key=lambda e: e["tflops_algorithmic"])
            print(json.dumps({"metric": "xtx_tflops", "value": best["tflops_algorithmic"], "unit": "TFLOP/s", "n_gpus": world,
                              "steps": 5, "warmup": 1, "ms_per_step": best["ms"], "higher_is_better": True, "scaling": "weak",
                              "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                              "config": {"workload": "K1 X^T X / S (statistics.py:76-87), fp32-faithful 3xTF32 on tcgen05, "
                                                     f"upper tiles only; best of the cases (S={best['S']}, n={best['n']})"},
                              "cases": res,
                              "roofline": {"bound": "tensor", "achieved": best["tflops_issued_tf32"], "peak": peaks["tf32_tflops"],
                                           "unit": "TFLOP/s", "frac": best["tflops_issued_tf32"] / peaks["tf32_tflops"],
                                           "traffic": None, "peak_source": peaks["tf32_source"],
                                           "note": "issued TF32 MMA flop (3 MMAs per product, upper tiles) / time"},
                              "peaks": peaks}), flush=True)
        return

    if args.config == "c5":
        res = sharded_layer(args, D, args.steps, max(args.warmup, 1), peaks)
        if rank == 0:
            ph = res["phases_ms_max_over_ranks"]
            print(json.dumps({"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                              "warmup": max(args.warmup, 1), "ms_per_step": res["ms_per_layer"], "higher_is_better": True,
                              "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                              "config": {"workload": res["workload"], "l2": "H alone (3.3 GB) exceeds the 126 MB L2"},
                              "clocks": res["clocks"], "sharded_c5": res,
                              "roofline": {"bound": "tensor", "kernel": "chol_factor", "achieved": res["factor_tflops_fp64"],
                                           "peak": peaks["fp64_tflops"] * world, "unit": "TFLOP/s",
                                           "frac": res["factor_frac_of_fp64_peak_x_gpus"], "traffic": None,
                                           "peak_source": peaks["fp64_source"] + f" x {world} GPUs",
                                           "note": "n^3/3 fp64 flop of the factor / its phase time (gather + tile kernel + "
                                                   f"export + two barriers), {ph.get('factor')} ms"},
                              "peaks": peaks}), flush=True)
        return

    cfg, shapes = config_shapes(args)
    L = len(shapes)
    weights = sum(r * n for r, n in shapes)
    cb = codebook.UniformCodebook(cfg["codebook"], -1, 1)

    # ---- inputs: W on host (pinned) and device; H = X^T X / S built on the device by K1 ----
    Wh, Hh, Mh, Wd, Hd, Md = [], [], [], [], [], []
    xtx_ms, xtx_flop = 0.0, 0.0
    for i, (r, n) in enumerate(shapes):
        lid = i + 1000 * rank                      # every rank owns a different layer set (weak scaling)
        w = torch.from_numpy(wl.synthetic_weight(r, n, lid)).pin_memory()
        x = torch.from_numpy(wl.synthetic_calibration(n, lid, SAMPLES)).to(dev)
        h = torch.zeros((n, n), dtype=torch.float32, device=dev)
        m = torch.zeros(n, dtype=torch.float32, device=dev)
        ops.hessian_accum(x, h, m, 0.0, SAMPLES)    # untimed first touch
        h.zero_(); m.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.hessian_accum(x, h, m, 0.0, SAMPLES)
        e1.record()
        torch.cuda.synchronize()
        xtx_ms += e0.elapsed_time(e1)
        xtx_flop += 2.0 * SAMPLES * n * n
        Wh.append(w)
        Wd.append(w.to(dev))
        Hd.append(h)
        Md.append(m)
        Hh.append(h.cpu().pin_memory())
        Mh.append(m.cpu())
        del x
    means = Md if cfg["bias"] else None
    errs = torch.zeros(L, dtype=torch.float32, device=dev)

    from sleekit_b200.pipeline import LayerSetQuantizer

    def make(streams, big_first):
        return LayerSetQuantizer(cb, scaling_mode=cfg["scaling"], act_order="diag", damp=DAMP, nb_ls_moves=cfg["moves"],
                                 grid_size=GRID, streams=streams, big_first=big_first, bias_correction=cfg["bias"],
                                 batch_k2=True if os.environ.get("SLK_BATCH_K2") is None else None)

    lsq = make(args.streams, not args.model_order)

    def step_device():
        lsq(Wd, Hd, errs_out=errs, keep_outputs=False, means=means)

    # ---- warm-up, then a profiled SERIAL pass: per-operation device time ---------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    # one stream: an event pair then brackets exactly one operation's kernels (on overlapping streams
    # they would include each other); the factorisations are batched exactly as in the timed pass
    serial = make(1, not args.model_order)
    fullh_uncertified = None
    if cfg["scaling"].startswith("hessian"):
        # the screened full-H search counts the rows whose minimum it cannot certify as global (0 expected)
        scaling.FULLH_UNCERTIFIED = torch.zeros(1, dtype=torch.int32, device=dev)
    serial(Wd, Hd, errs_out=errs, keep_outputs=False, means=means)
    torch.cuda.synchronize()
    if scaling.FULLH_UNCERTIFIED is not None:
        fullh_uncertified = int(scaling.FULLH_UNCERTIFIED.item())
        scaling.FULLH_UNCERTIFIED = None
    ops.PROFILE = {}
    serial(Wd, Hd, errs_out=errs, keep_outputs=False, means=means)
    torch.cuda.synchronize()
    phases = ops.profile_totals_ms(ops.PROFILE)
    serial_profile = ops.PROFILE
    ops.PROFILE = None
    groups = getattr(serial, "last_groups", None)
    rooflines = phase_rooflines(phases, serial_profile, shapes, groups, cfg, peaks)
    top = max(rooflines, key=lambda k: rooflines[k]["share_of_serial_device_time"]) if rooflines else None

    graph = None
    if not args.no_graph:
        # one pass recorded as a CUDA graph: the streams become parallel branches, replay has no
        # per-kernel CPU launch cost (the eager pass is bound by ~4 us of CPU per launch)
        graph, gerrs, _ = lsq.capture(Wd, Hd, means=means)

        def step_device():
            graph.replay()

        for _ in range(2):
            step_device()
        torch.cuda.synchronize()
        errs = gerrs

    # ---- timed region (device resident) ------------------------------------------------------
    sampler = ClockSampler(D.local)
    sampler.start()
    ms, wall = D.timed(step_device, args.steps)
    clocks = sampler.stop(D.last_window)
    # a replay launches the kernels recorded at capture time; count them with one eager pass
    launches0 = ops.launch_count()
    serial(Wd, Hd, errs_out=torch.zeros_like(errs), keep_outputs=False, means=means)
    torch.cuda.synchronize()
    launches = (ops.launch_count() - launches0) * args.steps
    ms_per_step = ms / args.steps
    value = world * weights / (ms_per_step * 1e-3)
    layer_err = float(errs.mean().item())

    # ---- where the step's time goes, measured INSIDE the timed configuration --------------------------
    # The timed step is a multi-stream graph: kernels of different layers overlap, so no event pair brackets
    # one kernel.  A second capture of the same pass carries stream-ordered %globaltimer stamps
    # (slk_debug_timestamp) around every layer's phases and around every factor launch; one untimed replay of
    # it gives each factor launch's duration as it runs beside the other layers' work.
    in_step = None
    if graph is not None and lsq.batch_k2 and L > 1:
        trace = torch.zeros((L, 8), dtype=torch.int64, device=dev)
        lsq.trace = trace
        g2, _, _ = lsq.capture(Wd, Hd, means=means)
        lsq.trace = None
        g2.replay()
        torch.cuda.synchronize()
        trace.zero_()
        g2.replay()
        torch.cuda.synchronize()
        tt = trace.cpu().numpy().astype(np.float64)
        t0 = tt[tt > 0].min()
        tt = np.where(tt > 0, (tt - t0) / 1e6, np.nan)                      # ms since the first stamp
        pass_ms = float(np.nanmax(tt))
        launches_k2 = []
        for (n_, b_), head in zip(lsq.last_groups, lsq.last_group_heads):
            b0, b1 = float(tt[head, 4]), float(tt[head, 5])
            launches_k2.append({"n": n_, "matrices": b_, "begin_ms": round(b0, 3), "end_ms": round(b1, 3), "ms": round(b1 - b0, 4),
                                "achieved": round(b_ * n_ ** 3 / 3.0 / ((b1 - b0) * 1e-3) / 1e12, 3),
                                "share_of_step": round((b1 - b0) / pass_ms, 3)})
        in_step = {"pass_ms_with_stamps": round(pass_ms, 3),
                   "scale_search_window_ms": [round(float(np.nanmin(tt[:, 1])), 3), round(float(np.nanmax(tt[:, 2])), 3)],
                   "sweep_window_ms": [round(float(np.nanmin(tt[:, 6])), 3), round(float(np.nanmax(tt[:, 7])), 3)],
                   "factor_launches": launches_k2,
                   "note": "device %globaltimer stamps on the launching streams, one untimed replay of the same graph with the "
                           "stamps captured in; the factor launches run beside the other layers' scale searches and sweeps"}
        del g2
    err_check = None
    if not cfg["moves"]:
        # the timed pass takes each layer's error from the sweep's residuals (sum E^2 - damp * sum D^2,
        # obq.gptq_device); check it here, untimed, against the explicit ((W-Q) H (W-Q)^T) product (K6)
        errs_fused = errs.clone()
        errs_k6 = torch.zeros_like(errs)
        obq.USE_SWEEP_ERROR = False
        serial(Wd, Hd, errs_out=errs_k6, keep_outputs=False, means=means)
        obq.USE_SWEEP_ERROR = True
        torch.cuda.synchronize()
        err_check = {"max_rel_diff_vs_product": float(((errs_fused - errs_k6).abs() / errs_k6.abs()).max().item()),
                     "layers": L, "note": "layer error from the sweep residuals vs the explicit K6 product, per layer"}

    # ---- end to end with HOST buffers: H2D and D2H inside the timed region -----------------------
    # (a) the layer-set plan: inputs in page-locked host memory, every layer a graph branch
    #     H2D(W, H) -> hot path -> D2H(Q, err), so the copy engines run under other layers' kernels
    # (a') the same plan returning codes (uint8) + row scales instead of de-scaled fp32 weights
    # (b) the reference's per-call numpy API on pageable arrays (experiments/compare.py:84-95)
    e2e = e2e_codes = e2e_numpy = None
    Wnp = [w.numpy() for w in Wh]
    Hnp = [h.numpy() for h in Hh]
    Mnp = [m.numpy() for m in Mh]
    if not args.no_e2e:
        for kind in ("weights", "codes"):
            plan = lsq.host_plan(shapes, outputs=kind)
            for i in range(L):
                plan.W[i][...] = Wnp[i]
                plan.H[i][...] = Hnp[i]
                if cfg["bias"]:
                    plan.M[i][...] = Mnp[i]
            plan.run()
            plan.run()
            ems, ewall = D.timed(lambda: plan.run(sync=True), args.steps)
            ent = {"value": world * weights / (ewall / args.steps * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": plan.h2d_bytes, "d2h_bytes_per_step": plan.d2h_bytes,
                   "ms_per_step": ewall / args.steps, "layer_error_mean": float(plan.err.mean()),
                   "api": (f"LayerSetQuantizer.host_plan(shapes, outputs='{kind}').run(): W and H of all layers in pinned host "
                           "buffers -> " + ("de-scaled quantized fp32 weights" if kind == "weights" else
                                            "codebook indices (uint8) + fp32 row scales")
                           + " and layer errors in pinned host buffers; wall clock, synchronised every step")}
            if kind == "weights":
                e2e = ent
            else:
                e2e_codes = ent
            del plan
        Harg = [(obq.remove_input_bias(Hnp[i], Mnp[i]) if cfg["bias"] else Hnp[i]) for i in range(L)]

        def step_numpy():
            # the reference-facing call sequence of experiments/compare.py:84-95, host numpy in and out
            for i in range(L):
                sc = scaling.compute_scaling(Wnp[i], cb, Harg[i], mode=cfg["scaling"], grid_size=GRID)
                q = scaling.quantize_with_scaling(Wnp[i], sc, cb, H=Harg[i], damp=DAMP, nb_ls_moves=cfg["moves"])
                obq.quantization_error(Wnp[i], q, H=Harg[i])

        step_numpy()
        cv.H2D_BYTES = cv.D2H_BYTES = 0
        nsteps = max(1, min(args.steps, 2))
        ems, ewall = D.timed(step_numpy, nsteps)
        e2e_numpy = {"value": world * weights / (ewall / nsteps * 1e-3), "unit": UNIT,
                     "h2d_bytes_per_step": cv.H2D_BYTES // nsteps, "d2h_bytes_per_step": cv.D2H_BYTES // nsteps,
                     "ms_per_step": ewall / nsteps, "steps": nsteps,
                     "api": "compute_scaling + quantize_with_scaling + quantization_error on pageable numpy arrays "
                            "(the drop-in sleekit.* functions; repeated operands are uploaded once per step through the "
                            "identity-keyed device cache of _convert)"}

    # ---- the sharded layer (configs[4]) on the same ranks ----------------------------------------------
    sharded = None
    if args.config == "c2" and not args.no_sharded and not args.layers and not args.only:
        try:
            sharded = sharded_layer(args, D, 2, 1, peaks)
        except Exception as ex:   # never lose the main line to the secondary measurement
            sharded = {"error": repr(ex)[:300]}

    xtx_big = xtx_measure(dev, [(8192, 4096)], reps=3) if args.config == "c2" else None
    # K1 over the whole layer set at once: the calibration products of independent layers as parallel branches of
    # one CUDA graph (what a calibration pass over a model does), instead of one launch after the other
    xtx_set = None
    if args.config == "c2" and not args.layers and not args.only:
        try:
            xs = [torch.from_numpy(wl.synthetic_calibration(n, i + 1000 * rank, SAMPLES)).to(dev) for i, (r, n) in enumerate(shapes)]
            hs = [torch.zeros((n, n), dtype=torch.float32, device=dev) for r, n in shapes]
            ms_ = [torch.zeros(n, dtype=torch.float32, device=dev) for r, n in shapes]
            sts = lsq.streams

            def xtx_pass():
                main = torch.cuda.current_stream()
                ev0 = torch.cuda.Event()
                ev0.record(main)
                for i in range(L):
                    st_ = sts[i % len(sts)]
                    st_.wait_event(ev0)
                    with torch.cuda.stream(st_):
                        ops.hessian_accum(xs[i], hs[i], ms_[i], 0.0, SAMPLES)
                for st_ in sts[: min(L, len(sts))]:
                    ev1 = torch.cuda.Event()
                    ev1.record(st_)
                    main.wait_event(ev1)

            xtx_pass()
            torch.cuda.synchronize()
            gx = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gx):
                xtx_pass()
            gx.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                gx.replay()
            e1.record()
            torch.cuda.synchronize()
            set_ms = e0.elapsed_time(e1) / 3
            xtx_set = {"ms": set_ms, "tflops": xtx_flop / (set_ms * 1e-3) / 1e12,
                       "note": f"the {L} products as parallel branches of one CUDA graph (one replay = the whole set)"}
            del xs, hs, ms_, gx
        except Exception as ex:
            xtx_set = {"error": repr(ex)[:200]}

    if rank != 0:
        return

    # ---- CPU arm on the same layers + parity of the timed GPU results against it ----------------------
    cpu_baseline = parity = None
    if not args.no_cpu_baseline:
        arm = CpuArm()
        row_div = args.cpu_row_div or default_row_div(args.config)
        outs, scs, _ = make(1, False)(Wd, Hd, keep_outputs=True, means=means)
        torch.cuda.synchronize()
        # identical inputs on both sides: the host copy of the Hessian K1 built on the device
        est, bw, meas, results = cpu_sample(arm, cfg, shapes, lambda lid: (Wnp[lid], Hnp[lid], Mnp[lid]), row_div)
        cpu_baseline = {"value": bw / est, "unit": UNIT, "cores": os.cpu_count(), "kind": arm.kind,
                        "sample": sample_text(cfg, shapes, row_div, arm) + f"; {meas:.1f} s of CPU work"}
        per = []
        for lid, res in results.items():
            rows = res["rows"]
            qg, sg = outs[lid][:rows].cpu().numpy(), scs[lid][:rows].cpu().numpy()
            same_scale = float((sg == res["sc"]).mean())
            ok = sg == res["sc"]                                  # compare codes where both sides use the same scale
            ca = float((arm.codes(qg[ok], sg[ok], cfg["codebook"]) == arm.codes(res["q"][ok], res["sc"][ok], cfg["codebook"])).mean()) if ok.any() else None
            Hq = obq.remove_input_bias(Hnp[lid], Mnp[lid]) if cfg["bias"] else Hnp[lid]
            eg = float(obq.quantization_error(Wnp[lid][:rows], qg, Hq))
            per.append({"layer": lid, "shape": list(shapes[lid]), "rows_compared": rows, "scale_equal_frac": same_scale,
                        "code_agreement": ca, "layer_error_gpu": eg, "layer_error_cpu": res["err"],
                        "layer_error_rel_diff": abs(eg - res["err"]) / abs(res["err"])})
        parity = {"against": arm.kind, "layers": per,
                  "code_agreement_min": min(p["code_agreement"] for p in per if p["code_agreement"] is not None),
                  "scale_equal_frac_min": min(p["scale_equal_frac"] for p in per),
                  "layer_error_rel_diff_max": max(p["layer_error_rel_diff"] for p in per),
                  "bar": "BASELINE north_star: layer error within 1e-3 relative, code agreement reported"}
        del outs, scs

    roofline = rooflines.get(top) if top else None
    big = max(in_step["factor_launches"], key=lambda e: e["ms"]) if in_step and in_step["factor_launches"] else None
    if big and (top == "chol_factor" or big["share_of_step"] >= 0.4):
        # the kernel with the largest share of the TIMED step: the batched factor launch of the widest layers
        tr_ = ncu_traffic().get(f"chol_factor_batched:n={big['n']}x{big['matrices']}")
        roofline = {"kernel": "chol_dag_kernel (slk_chol_factor_batched_f32, %d matrices of n = %d in one launch)" % (big["matrices"], big["n"]),
                    "bound": "tensor", "unit": "TFLOP/s", "achieved": big["achieved"], "peak": peaks["fp64_tflops"],
                    "frac": big["achieved"] / peaks["fp64_tflops"], "peak_source": peaks["fp64_source"],
                    "avg_launch_ms": big["ms"], "launches_timed": 1, "share_of_step": big["share_of_step"],
                    "traffic": tr_["dram_bytes_per_launch"] if tr_ else None,
                    "by_launch": in_step["factor_launches"],
                    "note": ("dominant kernel of the timed step by duration.  Algorithmic work n^3/3 fp64 flop per matrix (factor "
                             "only: the sweep's R form needs no inverse) / the launch's duration INSIDE the step (device "
                             "%globaltimer stamps on its stream; it runs with one CTA per SM while the other layers' scale "
                             "searches and sweeps share the SMs).  FP64 tensor path (DMMA); peak = cuBLAS fp64 GEMM measured in "
                             "this run.  Alone on the GPU the same launch reaches the figure in profiles/ (ncu: DMMA pipe 56 % "
                             "active).  rooflines_all_phases has every operation of the serial profile pass.")}
    issued3 = {"gptq_sweep", "hweighted_error", "local_search"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg, L), "layers": L, "weights_per_rank": weights,
                   "parallelism": f"independent layer sets x{world}" if world > 1 else "single GPU",
                   "streams": args.streams, "cuda_graph": graph is not None, "batched_factor": bool(lsq.batch_k2),
                   "l2": f"inputs (W+H ~{sum(4 * r * n + 4 * n * n for r, n in shapes) / 1e9:.2f} GB per rank) are larger than "
                         "the 126 MB L2; no flush needed"},
        "clocks": clocks, "e2e": e2e, "e2e_codes": e2e_codes, "e2e_numpy_api": e2e_numpy, "gpu_launches": launches,
        "roofline": roofline, "in_step": in_step,
        "rooflines_all_phases": {k: {kk: v[kk] for kk in ("bound", "achieved", "peak", "unit", "frac", "share_of_serial_device_time")}
                                 for k, v in rooflines.items()},
        "tensor_issue_note": f"3xTF32: kernels {sorted(issued3 & set(rooflines))} issue 3 TF32 MMA flop per algorithmic flop",
        "cpu_baseline": cpu_baseline, "parity": parity,
        "layer_error_mean": layer_err, "layer_error_check": err_check, "fullh_uncertified_rows": fullh_uncertified,
        "serial_phases_ms_per_step": {k: round(v[0], 3) for k, v in sorted(phases.items(), key=lambda kv: -kv[1][0])},
        "xtx": {"tflops": xtx_flop / (xtx_ms * 1e-3) / 1e12 if xtx_ms else None, "ms_total": xtx_ms,
                "note": (f"K1 X^T X over the {L} calibration matrices (S={SAMPLES}), one launch each, algorithmic 2*S*n^2 flop; "
                         "tcgen05 3xTF32, upper tiles only, incl. the hi/lo split+transpose pass"),
                "s8192_n4096": xtx_big[0] if xtx_big else None, "layer_set_concurrent": xtx_set},
        "sharded_c5": sharded, "peaks": peaks,
        "wall_ms_per_step": wall / args.steps,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
