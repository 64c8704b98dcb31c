"""Layer-set driver: quantize many independent layers with the per-layer hot path.

The reference's scripts loop over layers one by one (experiments/compare.py:50-135).  Layers
never interact, and small layers (OPT-125M: n = 768) are bound by the latency of their serial
chains (n dependent sweep columns, n/64 dependent factor panels), not by throughput.  This
driver therefore issues consecutive layers round-robin on a few CUDA streams so that one
layer's serial chain overlaps the others' -- same kernels, same results, no collective.
"""

from __future__ import annotations

import torch

from . import _lib, ops
from .scaling import quantize_scaled_device
from .statistics import _device_scaling


def issue_order(shapes, order="big"):
    """Order in which the layers' work is issued, for shapes [(rows, cols), ...].
    "big": longest serial chains first (the fp64 factor chain and the sweep chain grow with the
    number of columns) -- best when the inputs are already on the device, and measured best from
    pinned host buffers too (OPT-125M set: 19.4 ms against 19.8 / 20.2 ms for the other two);
    "interleaved": the same sorted list dealt out as one long-chain layer followed by its share of the
    short ones, so that host->device copies deliver work for all SMs from the start;
    "model": as given."""
    L = len(shapes)
    if order == "model":
        return list(range(L))
    by_size = sorted(range(L), key=lambda k: (-shapes[k][1], -shapes[k][0], k))
    if order == "big":
        return by_size
    if order != "interleaved":
        raise ValueError(f"unknown issue order {order!r}")
    nmin = min(sh[1] for sh in shapes)
    big = [k for k in by_size if shapes[k][1] >= 2 * nmin]
    small = [k for k in by_size if shapes[k][1] < 2 * nmin]
    if not big or not small:
        return by_size
    out, per, taken = [], len(small) / len(big), 0
    for i, b in enumerate(big):
        out.append(b)
        upto = round((i + 1) * per)
        out.extend(small[taken:upto])
        taken = upto
    out.extend(small[taken:])
    return out


class LayerSetQuantizer:
    """Holds the side streams; call it with lists of device tensors."""

    def __init__(self, codebook, scaling_mode="diag", act_order="diag", damp=0.01, nb_ls_moves=0, grid_size=100,
                 min_factor=0.05, max_factor=1.0, streams=8, big_first=True):
        ops.require_cuda()
        self.cb = codebook
        self.scaling_mode, self.act_order = scaling_mode, act_order
        self.damp, self.nb_ls_moves = damp, nb_ls_moves
        self.grid_size, self.min_factor, self.max_factor = grid_size, min_factor, max_factor
        self.streams = [torch.cuda.Stream() for _ in range(max(1, int(streams)))]
        self.big_first = bool(big_first)

    def _one(self, W, H):
        sc = _device_scaling(W, self.cb, H, self.scaling_mode, self.grid_size, self.min_factor, self.max_factor)
        # layer error (obq.py:89-103): from the sweep's own residuals when possible (gptq_device)
        q, (err, _) = quantize_scaled_device(W, sc, self.cb, H, self.act_order, self.damp, self.nb_ls_moves,
                                             want_err=True)
        return q, sc, err

    def _issue_order(self, shapes, order=None):
        return issue_order(shapes, order or ("big" if self.big_first else "model"))

    def __call__(self, Ws, Hs, errs_out=None, keep_outputs=True, _in_capture=False, _pre=None, _post=None, _order=None):
        """Ws[i] [r_i, n_i] fp32, Hs[i] [n_i, n_i] fp32 on the device.  Returns (quantized
        weights, scales, errors) -- errors as one fp32 device vector.  Nothing synchronises.
        _pre(i) / _post(i, q) run on layer i's stream before / after its kernels (HostPlan uses
        them for the host<->device copies)."""
        L = len(Ws)
        dev = Ws[0].device
        errs = errs_out if errs_out is not None else torch.empty(L, dtype=torch.float32, device=dev)
        main = torch.cuda.current_stream()
        start = torch.cuda.Event()
        start.record(main)
        outs, scales = [None] * L, [None] * L
        S = len(self.streams)
        # longest serial chains first (the fp64 factor chain grows with n, the sweep chain with n):
        # their launches then enter the queues ahead of the short layers that fill the gaps
        issue = self._issue_order([tuple(w.shape) for w in Ws], _order)
        for slot, i in enumerate(issue):
            st = self.streams[slot % S]
            if slot < S:
                st.wait_event(start)
            with torch.cuda.stream(st):
                if _pre is not None:
                    _pre(i)
                q, sc, e = self._one(Ws[i], Hs[i])
                errs[i:i + 1].copy_(e.reshape(1))
                if _post is not None:
                    _post(i, q)
                if keep_outputs:
                    outs[i], scales[i] = q, sc
                    if not _in_capture:
                        q.record_stream(main)
                        sc.record_stream(main)
        for st in self.streams[: min(S, L)]:
            done = torch.cuda.Event()
            done.record(st)
            main.wait_event(done)
        return outs, scales, errs

    def capture(self, Ws, Hs):
        """Record one pass over the layer set into a CUDA graph (the side streams become parallel
        branches) and return (graph, errors, quantized weights).  graph.replay() re-runs the whole
        pass on the current contents of Ws / Hs with no per-kernel CPU launch cost; outputs are
        rewritten in place at every replay.  Call once after a warm-up pass."""
        L = len(Ws)
        errs = torch.empty(L, dtype=torch.float32, device=Ws[0].device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            outs, _, _ = self(Ws, Hs, errs_out=errs, keep_outputs=True, _in_capture=True)
        return graph, errs, outs

    def host_plan(self, shapes, symmetric_h=True):
        """Pinned host buffers + device buffers + one CUDA graph for a fixed list of layer shapes
        [(r, n), ...]; see HostPlan."""
        return HostPlan(self, shapes, symmetric_h)


class HostPlan:
    """Host-buffer entry point of the layer-set driver (the experiments' per-layer loop,
    experiments/compare.py:50-135, over host arrays).

    The plan owns page-locked host buffers ``W[i]``, ``H[i]`` (inputs, numpy views the caller
    fills) and ``Q[i]``, ``err`` (outputs).  ``run()`` replays ONE CUDA graph in which every layer
    is a branch: H2D copy of its W and H -> scale search -> GPTQ -> layer error -> D2H copy of
    the quantized weights; the copy engines therefore work under the kernels of the other
    layers.  Same kernels and results as the device-tensor path.  H is a Hessian, hence symmetric:
    by default only its block upper triangle is sent over PCIe and mirrored on the device
    (slk_upload_symmetric_f32; 53-63 % of the bytes); symmetric_h=False sends the whole matrix."""

    def __init__(self, lsq, shapes, symmetric_h=True):
        ops.require_cuda()
        self.lsq = lsq
        # H is a Hessian X^T X / n: symmetric.  With symmetric_h only its block upper triangle crosses
        # PCIe (ops.upload_symmetric) and the device mirrors it; pass False for arbitrary matrices.
        self.symmetric_h = bool(symmetric_h)
        import os
        # issue order of the plan (LayerSetQuantizer._issue_order); measured on the OPT-125M set from pinned
        # host buffers: big 19.4 ms, interleaved 19.8 ms, model 20.2 ms per pass -> the quantizer's default
        self.order = os.environ.get("SLK_PLAN_ORDER") or None
        self.shapes = [(int(r), int(n)) for r, n in shapes]
        dev = ops.device()
        f32 = torch.float32
        self._Wp = [torch.empty((r, n), dtype=f32).pin_memory() for r, n in self.shapes]
        self._Hp = [torch.empty((n, n), dtype=f32).pin_memory() for r, n in self.shapes]
        self._Qp = [torch.empty((r, n), dtype=f32).pin_memory() for r, n in self.shapes]
        self._errp = torch.empty(len(self.shapes), dtype=f32).pin_memory()
        self.W = [t.numpy() for t in self._Wp]
        self.H = [t.numpy() for t in self._Hp]
        self.Q = [t.numpy() for t in self._Qp]
        self.err = self._errp.numpy()
        self._Wd = [torch.empty((r, n), dtype=f32, device=dev) for r, n in self.shapes]
        self._Hd = [torch.empty((n, n), dtype=f32, device=dev) for r, n in self.shapes]
        self._errd = torch.empty(len(self.shapes), dtype=f32, device=dev)
        self._graph = None
        lib = _lib.load()
        self.h2d_bytes = sum(4 * r * n + (int(lib.slk_upload_symmetric_bytes(n, ops.symmetric_block_rows(n)))
                                          if self.symmetric_h else 4 * n * n) for r, n in self.shapes)
        self.d2h_bytes = sum(4 * r * n for r, n in self.shapes) + 4 * len(self.shapes)

    def _pre(self, i):
        self._Wd[i].copy_(self._Wp[i], non_blocking=True)
        if self.symmetric_h:
            ops.upload_symmetric(self._Hp[i], self._Hd[i])
        else:
            self._Hd[i].copy_(self._Hp[i], non_blocking=True)

    def _post(self, i, q):
        self._Qp[i].copy_(q, non_blocking=True)

    def _pass(self, in_capture):
        self.lsq(self._Wd, self._Hd, errs_out=self._errd, keep_outputs=False, _in_capture=in_capture,
                 _pre=self._pre, _post=self._post, _order=self.order)
        self._errp.copy_(self._errd, non_blocking=True)

    def run(self, sync=True):
        """One pass over the layer set from the current contents of W / H; fills Q and err."""
        if self._graph is None:
            # first call: identity Hessians would do as well -- the caller's data is used so that
            # the warm-up pass is already a valid result; then record the graph
            self._pass(False)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._pass(True)
            self._graph = g
        self._graph.replay()
        if sync:
            torch.cuda.current_stream().synchronize()
        return self.Q, self.err


def quantize_layer_sharded(W_rows, X_rows, codebook, scaling_mode="diag", act_order="diag", damp=0.01, nb_ls_moves=0,
                           bias_correction=False, grid_size=100, min_factor=0.05, max_factor=1.0, group=None):
    """One big layer over the ranks of a process group (BASELINE config 5).

    W_rows [r_local, n]: this rank's contiguous slice of the output rows of W (dist.row_partition).
    X_rows [S_local, n]: this rank's slice of the calibration samples.
    1. every rank folds its samples into running statistics (K1) and ONE all-reduce of
       [n*n + n + 1] floats makes H and the mean identical everywhere (statistics.py:76-87);
    2. scale search, ordering, fp64 factor, sweep, local search and the row errors run on the local
       rows with no communication (rows never interact); the factor is recomputed on every rank
       (deterministic, hence identical) -- it is the term that does not shard;
    3. the only other exchanges: the column residual sums of the err / sqerr orderings (n floats)
       and the scalar layer error.
    Returns (quantized rows [r_local, n], scales [r_local], layer error (0-dim tensor), H, mean)."""
    import torch.distributed as tdist

    from . import dist as sdist

    ops.require_cuda()
    dev = W_rows.device
    n = W_rows.shape[1]
    H = torch.zeros((n, n), dtype=torch.float32, device=dev)
    mean = torch.zeros(n, dtype=torch.float32, device=dev)
    count = int(X_rows.shape[0])
    if count:
        ops.hessian_accum(X_rows.contiguous(), H, mean, 0.0, count)
    H, mean, count = sdist.allreduce_statistics(H, mean, count, group)
    Hq = ops.remove_input_bias(H, mean) if bias_correction else H
    sc = _device_scaling(W_rows, codebook, Hq, scaling_mode, grid_size, min_factor, max_factor)
    reduce_cols = (lambda c: sdist.allreduce_column_sums(c, group)) if act_order in ("err", "sqerr") else None
    q, (_, rows_err) = quantize_scaled_device(W_rows, sc, codebook, Hq, act_order, damp, nb_ls_moves,
                                              colsum_reduce=reduce_cols, want_err=True)
    total_rows = torch.tensor([W_rows.shape[0]], dtype=torch.int64, device=dev)
    if tdist.is_available() and tdist.is_initialized() and tdist.get_world_size(group) > 1:
        tdist.all_reduce(total_rows, group=group)
    err = sdist.allreduce_row_error_mean(rows_err, int(total_rows.item()), group)
    return q, sc, err, Hq, mean
