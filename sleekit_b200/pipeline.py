"""Layer-set driver: quantize many independent layers with the per-layer hot path.

The reference's scripts loop over layers one by one (experiments/compare.py:50-135).  Layers
never interact, and small layers (OPT-125M: n = 768) are bound by the latency of their serial
chains (n dependent sweep columns, n/64 dependent factor panels), not by throughput.  This
driver therefore issues consecutive layers round-robin on a few CUDA streams so that one
layer's serial chain overlaps the others' -- same kernels, same results, no collective.
"""

from __future__ import annotations

import os

import torch

from . import _lib, obq, ops
from .scaling import quantize_scaled_device
from .statistics import _device_scaling


def issue_order(shapes, order="big"):
    """Order in which the layers' work is issued, for shapes [(rows, cols), ...].
    "big": longest serial chains first (the fp64 factor chain and the sweep chain grow with the
    number of columns) -- best when the inputs are already on the device, and measured best from
    pinned host buffers too (OPT-125M set: 19.4 ms against 19.8 / 20.2 ms for the other two);
    "interleaved": the same sorted list dealt out as one long-chain layer followed by its share of the
    short ones, so that host->device copies deliver work for all SMs from the start;
    "small": shortest chains first (their inputs are small: work reaches the GPU at once);
    "model": as given."""
    L = len(shapes)
    if order == "model":
        return list(range(L))
    by_size = sorted(range(L), key=lambda k: (-shapes[k][1], -shapes[k][0], k))
    if order == "big":
        return by_size
    if order == "small":
        return by_size[::-1]
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
                 min_factor=0.05, max_factor=1.0, streams=8, big_first=True, batch_k2=None, k2_group=None,
                 bias_correction=False):
        """batch_k2: factor the Hessians of equal-sized layers in batched launches whose tile tasks share
        one ticket queue (ops.chol_factor_batched) instead of one factorisation per layer; default on
        when there is more than one stream (SLK_BATCH_K2=0/1 overrides).  k2_group: at most this many
        layers per batched launch (SLK_K2_GROUP)."""
        ops.require_cuda()
        self.bias_correction = bool(bias_correction)      # H - m m^T (obq.py:14-25) before everything else
        env = os.environ.get("SLK_BATCH_K2")
        self.batch_k2 = (int(streams) > 1 if batch_k2 is None else bool(batch_k2)) if env is None else env != "0"
        self.k2_group = int(os.environ.get("SLK_K2_GROUP") or k2_group or 64)
        self.k2_single_from = 1 << 30      # layers at least this wide are factored one launch each (HostPlan "small" mode)
        # the factor launches and the layers with the longest chains run on high-priority streams: their
        # CTAs are scheduled ahead of the short layers' whenever both are ready (SLK_PRIORITY=0: off)
        real = os.environ.get("SLK_REAL_STREAMS", "1") != "0"
        mk = (lambda cnt, prio: ops.create_streams(cnt, prio)) if real else (
            lambda cnt, prio: [torch.cuda.Stream(priority=prio) for _ in range(cnt)])
        self.streams = mk(max(1, int(streams)), 0)
        self.use_priority = os.environ.get("SLK_PRIORITY", "1") != "0"
        hp = -1 if self.use_priority else 0
        self.k2_streams = mk(4, hp)
        self.hi_streams = mk(max(1, int(streams)), hp) if self.use_priority else None
        # CTAs per macro-block sweep launch while a layer SET is in flight (fewer, taller CTAs: less SM-slot
        # time per layer; measured on the sixty n = 768 layers: 6.2 ms with 24, 6.6 with 48, 7.1 with 96)
        self.sweep_ctas = int(os.environ.get("SLK_SET_SWEEP_CTAS", "24")) if int(streams) > 1 else 0
        self.infos = None      # optional int32 CUDA tensor [layers]: receives every layer's factorisation status
        self.trace = None      # development aid: int64 CUDA tensor [layers, 8] of phase time stamps (tools/timeline.py)
        self.cb = codebook
        self.scaling_mode, self.act_order = scaling_mode, act_order
        self.damp, self.nb_ls_moves = damp, nb_ls_moves
        self.grid_size, self.min_factor, self.max_factor = grid_size, min_factor, max_factor
        self.big_first = bool(big_first)

    def _hessian(self, H, mean):
        if not self.bias_correction:
            return H
        assert mean is not None, "bias_correction needs the per-layer input means"
        return ops.remove_input_bias(H, mean)

    def _one(self, W, H):
        sc = _device_scaling(W, self.cb, H, self.scaling_mode, self.grid_size, self.min_factor, self.max_factor)
        # layer error (obq.py:89-103): from the sweep's own residuals when possible (gptq_device)
        gs = obq.gptq_prepare(W, H, self.cb, self.act_order, self.damp, self.nb_ls_moves, row_scale=sc, want_err=True)
        q, (err, _) = obq.gptq_finish(gs)
        return q, sc, err, gs.info

    def _issue_order(self, shapes, order=None):
        return issue_order(shapes, order or ("big" if self.big_first else "model"))

    def __call__(self, Ws, Hs, errs_out=None, keep_outputs=True, _in_capture=False, _pre=None, _post=None, _order=None,
                 means=None):
        """Ws[i] [r_i, n_i] fp32, Hs[i] [n_i, n_i] fp32 on the device.  Returns (quantized
        weights, scales, errors) -- errors as one fp32 device vector.  Nothing synchronises.
        _pre(i) / _post(i, q) run on layer i's stream before / after its kernels (HostPlan uses
        them for the host<->device copies)."""
        L = len(Ws)
        dev = Ws[0].device
        errs = errs_out if errs_out is not None else torch.empty(L, dtype=torch.float32, device=dev)
        ops.set_option("sweep_ctas", self.sweep_ctas)
        try:
            return self._run(Ws, Hs, errs, keep_outputs, _in_capture, _pre, _post, _order, means)
        finally:
            ops.set_option("sweep_ctas", 0)

    def _run(self, Ws, Hs, errs, keep_outputs, _in_capture, _pre, _post, _order, means):
        L = len(Ws)
        if self.batch_k2 and L > 1 and obq.USE_CHOL_FORM:
            return self._staged(Ws, Hs, errs, keep_outputs, _in_capture, _pre, _post, _order, means)
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
                q, sc, e, info = self._one(Ws[i], self._hessian(Hs[i], means[i] if means is not None else None))
                errs[i:i + 1].copy_(e.reshape(1))
                if self.infos is not None:
                    self.infos[i:i + 1].copy_(info.reshape(1))
                if _post is not None:
                    _post(i, q, sc)
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

    def _staged(self, Ws, Hs, errs, keep_outputs, _in_capture, _pre, _post, _order, means=None):
        """The same pass in three stages so that the fp64 factorisations can be batched:
        A (per layer, its stream): [copy in] -> scale search -> damp, ordering, scaled + permuted weights;
        B (per group of equal-sized layers, a factor stream): ONE batched tile-task Cholesky launch;
        C (per layer, its stream): sweep -> un-permute / de-scale -> layer error -> [copy out].
        Same kernels and arithmetic as the per-layer path: identical results."""
        L = len(Ws)
        main = torch.cuda.current_stream()
        start = torch.cuda.Event()
        start.record(main)
        outs, scales = [None] * L, [None] * L
        S = len(self.streams)
        issue = self._issue_order([tuple(w.shape) for w in Ws], _order)
        stream_of, stage, done_a = {}, {}, {}
        Hs = list(Hs)
        nmax = max(int(w.shape[1]) for w in Ws)
        nmin = min(int(w.shape[1]) for w in Ws)
        used = []
        for slot, i in enumerate(issue):
            long_chain = (self.hi_streams is not None and S > 1 and nmax >= 2 * nmin
                          and int(Ws[i].shape[1]) * 2 > nmax)
            st = (self.hi_streams if long_chain else self.streams)[slot % S]
            stream_of[i] = st
            if st not in used:
                used.append(st)
                st.wait_event(start)
            with torch.cuda.stream(st):
                tr = self.trace
                if tr is not None:
                    ops.timestamp(tr, i, 0)
                if _pre is not None:
                    _pre(i)
                if tr is not None:
                    ops.timestamp(tr, i, 1)
                Hs[i] = self._hessian(Hs[i], means[i] if means is not None else None)
                sc = _device_scaling(Ws[i], self.cb, Hs[i], self.scaling_mode, self.grid_size, self.min_factor,
                                     self.max_factor)
                if tr is not None:
                    ops.timestamp(tr, i, 2)
                stage[i] = (sc, obq.gptq_prepare(Ws[i], Hs[i], self.cb, self.act_order, self.damp, self.nb_ls_moves,
                                                 row_scale=sc, want_err=True))
                if tr is not None:
                    ops.timestamp(tr, i, 3)
                ev = torch.cuda.Event()
                ev.record(st)
                done_a[i] = ev
        # groups of equal n in issue order
        groups, cur = [], []
        for i in issue:
            limit = 1 if int(Ws[i].shape[1]) >= self.k2_single_from else self.k2_group
            if cur and (Ws[cur[0]].shape[1] != Ws[i].shape[1] or len(cur) >= limit):
                groups.append(cur)
                cur = []
            cur.append(i)
        if cur:
            groups.append(cur)
        self.last_groups = [(int(Ws[m[0]].shape[1]), len(m)) for m in groups]   # (n, matrices) per factor launch
        self.last_group_heads = [m[0] for m in groups]                          # layer whose trace row holds the launch's stamps
        keep = []                                        # factor tensors stay alive until the join below
        for g, members in enumerate(groups):
            ks = self.k2_streams[g % len(self.k2_streams)]
            if g < len(self.k2_streams):
                ks.wait_event(start)
            for i in members:
                ks.wait_event(done_a[i])
            with torch.cuda.stream(ks):
                if self.trace is not None:
                    ops.timestamp(self.trace, members[0], 4)
                if len(members) == 1:
                    i = members[0]
                    facs = [ops.chol_factor(Hs[i], stage[i][1].order, stage[i][1].dampval)]
                else:
                    facs = ops.chol_factor_batched([Hs[i] for i in members], [stage[i][1].order for i in members],
                                                   [stage[i][1].dampval for i in members])
                if self.trace is not None:
                    ops.timestamp(self.trace, members[0], 5)
                ev = torch.cuda.Event()
                ev.record(ks)
            keep.append(facs)
            for i, fac in zip(members, facs):
                st = stream_of[i]
                st.wait_event(ev)
                with torch.cuda.stream(st):
                    sc, gs = stage[i]
                    if self.trace is not None:
                        ops.timestamp(self.trace, i, 6)
                    q, (e, _) = obq.gptq_finish(gs, fac)
                    errs[i:i + 1].copy_(e.reshape(1))
                    if self.infos is not None:
                        self.infos[i:i + 1].copy_(gs.info.reshape(1))
                    if _post is not None:
                        _post(i, q, sc)
                    if self.trace is not None:
                        ops.timestamp(self.trace, i, 7)
                    if keep_outputs:
                        outs[i], scales[i] = q, sc
                        if not _in_capture:
                            q.record_stream(main)
                            sc.record_stream(main)
        for st in used + list(self.k2_streams[: min(len(self.k2_streams), len(groups))]):
            done = torch.cuda.Event()
            done.record(st)
            main.wait_event(done)
        stage.clear()
        del keep
        return outs, scales, errs

    def capture(self, Ws, Hs, means=None):
        """Record one pass over the layer set into a CUDA graph (the side streams become parallel
        branches) and return (graph, errors, quantized weights).  graph.replay() re-runs the whole
        pass on the current contents of Ws / Hs with no per-kernel CPU launch cost; outputs are
        rewritten in place at every replay.  Call once after a warm-up pass."""
        L = len(Ws)
        errs = torch.empty(L, dtype=torch.float32, device=Ws[0].device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            outs, _, _ = self(Ws, Hs, errs_out=errs, keep_outputs=True, _in_capture=True, means=means)
        return graph, errs, outs

    def host_plan(self, shapes, symmetric_h=True, outputs="weights"):
        """Pinned host buffers + device buffers + one CUDA graph for a fixed list of layer shapes
        [(r, n), ...]; see HostPlan."""
        return HostPlan(self, shapes, symmetric_h, outputs)


class HostPlan:
    """Host-buffer entry point of the layer-set driver (the experiments' per-layer loop,
    experiments/compare.py:50-135, over host arrays).

    The plan owns page-locked host buffers ``W[i]``, ``H[i]`` (inputs, numpy views the caller
    fills) and ``Q[i]``, ``err`` (outputs).  ``run()`` replays ONE CUDA graph in which every layer
    is a branch: H2D copy of its W and H -> scale search -> GPTQ -> layer error -> D2H copy of
    the quantized weights; the copy engines therefore work under the kernels of the other
    layers.  Same kernels and results as the device-tensor path.  H is a Hessian, hence symmetric:
    by default only its block upper triangle is sent over PCIe and mirrored on the device
    (slk_upload_symmetric_f32; 53-63 % of the bytes); symmetric_h=False sends the whole matrix.
    outputs="weights": Q[i] are the de-scaled fp32 quantized weights, what the reference's
    quantize_with_scaling returns (scaling.py:80-81).  outputs="codes": the plan returns instead the
    codebook indices ``codes[i]`` (uint8 / uint16 [r, n], codebook.py:43-54 of Q / scale) and the row scales
    ``scales[i]`` [r] -- the packed form a deployment stores; a quarter of the device->host bytes.
    With the quantizer's bias_correction the plan also owns ``M[i]`` (input means, obq.py:14-25)."""

    def __init__(self, lsq, shapes, symmetric_h=True, outputs="weights"):
        ops.require_cuda()
        assert outputs in ("weights", "codes")
        self.outputs = outputs
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
        cdt = ops.device_codebook(lsq.cb).index_dtype
        if outputs == "weights":
            self._Qp = [torch.empty((r, n), dtype=f32).pin_memory() for r, n in self.shapes]
            self._Sp = []
        else:
            self._Qp = [torch.empty((r, n), dtype=cdt).pin_memory() for r, n in self.shapes]
            self._Sp = [torch.empty(r, dtype=f32).pin_memory() for r, n in self.shapes]
        self._Mp = [torch.empty(n, dtype=f32).pin_memory() for r, n in self.shapes] if lsq.bias_correction else []
        self._Md = [torch.empty(n, dtype=f32, device=dev) for r, n in self.shapes] if lsq.bias_correction else None
        self.M = [t.numpy() for t in self._Mp]
        self._errp = torch.empty(len(self.shapes), dtype=f32).pin_memory()
        self.W = [t.numpy() for t in self._Wp]
        self.H = [t.numpy() for t in self._Hp]
        self.Q = [t.numpy() for t in self._Qp] if outputs == "weights" else None
        self.codes = [t.numpy() for t in self._Qp] if outputs == "codes" else None
        self.scales = [t.numpy() for t in self._Sp] if outputs == "codes" else None
        self.err = self._errp.numpy()
        self._Wd = [torch.empty((r, n), dtype=f32, device=dev) for r, n in self.shapes]
        self._Hd = [torch.empty((n, n), dtype=f32, device=dev) for r, n in self.shapes]
        self._errd = torch.empty(len(self.shapes), dtype=f32, device=dev)
        self._infod = torch.zeros(len(self.shapes), dtype=torch.int32, device=dev)
        self._infop = torch.zeros(len(self.shapes), dtype=torch.int32).pin_memory()
        self.info = self._infop.numpy()       # per layer: 0, or the first non-positive pivot of its factorisation
        self._graph = None
        # high priority: the mirror kernel of the symmetric upload must not queue behind the layers' kernels
        self.copy_stream = ops.create_streams(1, -1)[0]
        self._copy_joined = False
        # the plan's layers arrive one after the other over PCIe: per-layer factorisations start as soon as
        # a layer's Hessian has landed, a batched launch only after the last one of its group
        # (measured: 19.7 ms per OPT-125M pass per layer, 27 ms batched); SLK_PLAN_BATCH_K2=1 batches anyway
        mode = os.environ.get("SLK_PLAN_BATCH_K2", "0")
        self.batch_k2 = mode != "0"
        self.batch_small_only = mode == "small"   # batch the narrow layers in groups of 12, factor the wide ones per layer
        lib = _lib.load()
        self.h2d_bytes = sum(4 * r * n + (int(lib.slk_upload_symmetric_bytes(n, ops.symmetric_block_rows(n)))
                                          if self.symmetric_h else 4 * n * n) for r, n in self.shapes)
        self.h2d_bytes += sum(4 * n for r, n in self.shapes) if lsq.bias_correction else 0
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in self._Qp + self._Sp) + 4 * len(self.shapes)

    def _pre(self, i):
        # All host->device copies go through ONE copy stream in the issue order of the layers, so the
        # inputs of the longest chains (issued first) really arrive first; a layer's stream waits for
        # its own copies only.  (Left to 72 independent streams the DMA engine interleaves the copies
        # and the big layers' Hessians land last.)
        cur = torch.cuda.current_stream()
        cs = self.copy_stream
        if not self._copy_joined:
            ev0 = torch.cuda.Event()
            ev0.record(cur)
            cs.wait_event(ev0)
            self._copy_joined = True
        with torch.cuda.stream(cs):
            self._Wd[i].copy_(self._Wp[i], non_blocking=True)
            if self.symmetric_h:
                ops.upload_symmetric_copy(self._Hp[i], self._Hd[i])   # DMA only: the copy queue never waits for a kernel
            else:
                self._Hd[i].copy_(self._Hp[i], non_blocking=True)
            if self._Md is not None:
                self._Md[i].copy_(self._Mp[i], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        cur.wait_event(ev)
        if self.symmetric_h:
            ops.mirror_symmetric(self._Hd[i])                         # on the layer's own stream

    def _post(self, i, q, sc):
        if self.outputs == "weights":
            self._Qp[i].copy_(q, non_blocking=True)
        else:
            # codebook.py:43-54 on Q / scale: two HBM passes, then a quarter of the bytes over PCIe
            idx = ops.round_to_codebook(ops.scale_rows(q, sc, 0), self.lsq.cb, want_val=False, want_idx=True)[1]
            self._Qp[i].copy_(idx, non_blocking=True)
            self._Sp[i].copy_(sc, non_blocking=True)

    def _pass(self, in_capture):
        self._copy_joined = False
        saved = (self.lsq.batch_k2, self.lsq.k2_single_from, self.lsq.k2_group)
        self.lsq.batch_k2 = self.batch_k2
        if self.batch_small_only:
            self.lsq.k2_single_from, self.lsq.k2_group = 2048, 12
        try:
            self._pass_inner(in_capture)
        finally:
            self.lsq.batch_k2, self.lsq.k2_single_from, self.lsq.k2_group = saved

    def _pass_inner(self, in_capture):
        self.lsq.infos = self._infod
        try:
            self.lsq(self._Wd, self._Hd, errs_out=self._errd, keep_outputs=False, _in_capture=in_capture,
                     _pre=self._pre, _post=self._post, _order=self.order, means=self._Md)
        finally:
            self.lsq.infos = None
        self._errp.copy_(self._errd, non_blocking=True)
        self._infop.copy_(self._infod, non_blocking=True)

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
            bad = [i for i, v in enumerate(self.info) if v != 0]
            if bad:
                # the reference raises from np.linalg.cholesky (obq.py:49-50); the plan reports which layers
                import numpy as np

                raise np.linalg.LinAlgError(f"Matrix is not positive definite (layers {bad[:8]}{'...' if len(bad) > 8 else ''})")
        return (self.Q if self.outputs == "weights" else (self.codes, self.scales)), self.err


class ShardedLayerQuantizer:
    """One big layer over the ranks of a process group (BASELINE config 5).

    Rank g holds a contiguous slice of the output rows of W (dist.row_partition) and a slice of the
    calibration samples.  Per call:
    1. K1 folds the local samples into running statistics; the count-weighted block upper triangle of
       H and the mean are summed over the ranks by ONE all-reduce (dist.allreduce_statistics), after
       which H and the mean are identical everywhere (statistics.py:76-87);
    2. scale search, ordering, sweep, local search and row errors run on the local rows with no
       communication (rows never interact: obq.py:106-137, :264-346, scaling.py:127-133);
    3. the fp64 factor of H_opt -- the one step that is not row-parallel (obq.py:38-55) -- is SHARED:
       tile row i of the tile-task Cholesky belongs to rank i % P, finished tiles are pushed into every
       peer's workspace through NVLink peer stores (ops.chol_factor_dist), so every rank ends with the
       complete factor at ~1/P of the fp64 work (dist_factor=False: every rank factors on its own);
    4. the only other exchanges: the column residual sums of the err / sqerr orderings (n floats) and
       the scalar layer error.
    phases_ms (after a call with timing=True): device time per phase on this rank."""

    def __init__(self, n, codebook, scaling_mode="diag", act_order="diag", damp=0.01, nb_ls_moves=0,
                 bias_correction=False, grid_size=100, min_factor=0.05, max_factor=1.0, group=None, dist_factor=True):
        import torch.distributed as tdist

        ops.require_cuda()
        self.n, self.cb, self.group = int(n), codebook, group
        self.scaling_mode, self.act_order, self.damp, self.nb_ls_moves = scaling_mode, act_order, damp, nb_ls_moves
        self.bias_correction = bias_correction
        self.grid_size, self.min_factor, self.max_factor = grid_size, min_factor, max_factor
        self.world = tdist.get_world_size(group) if (tdist.is_available() and tdist.is_initialized()) else 1
        self.dist_factor = bool(dist_factor) and self.world > 1
        self.pws = ops.PeerWorkspace(ops.chol_dist_ws_bytes(self.n), group) if self.dist_factor else None
        # the packed statistics are summed by our own NVLink kernel over a peer-visible buffer
        # (SLK_PEER_ALLREDUCE=0: NCCL all-reduce instead)
        self.pbuf = (ops.PeerBuffer(ops.sym_packed_len(self.n) + self.n, group)
                     if (self.world > 1 and os.environ.get("SLK_PEER_ALLREDUCE", "1") != "0") else None)
        self._token = torch.zeros(1, dtype=torch.float32, device=ops.device())
        self.phases_ms = {}
        self.info = None

    def _barrier(self):
        import torch.distributed as tdist

        tdist.all_reduce(self._token, group=self.group)        # stream-ordered: no host synchronisation

    def _factor_fn(self, Hd, order, dampval):
        fac = ops.chol_factor_dist(Hd, order, dampval, self.pws, self._barrier)
        import torch.distributed as tdist

        tdist.all_reduce(fac[3], op=tdist.ReduceOp.MAX, group=self.group)   # a non-PD pivot seen by any rank
        return fac

    def close(self):
        if self.pws is not None:
            self.pws.close()
            self.pws = None
        if self.pbuf is not None:
            self.pbuf.close()
            self.pbuf = None

    def __call__(self, W_rows, X_rows, timing=False):
        """W_rows [r_local, n], X_rows [S_local, n] fp32 on the device.  Returns (quantized rows,
        scales, layer error (0-dim), H, mean)."""
        import torch.distributed as tdist

        from . import dist as sdist

        dev = W_rows.device
        n = self.n
        marks = []

        def mark(name):
            if timing:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))

        mark("start")
        H = torch.zeros((n, n), dtype=torch.float32, device=dev)
        mean = torch.zeros(n, dtype=torch.float32, device=dev)
        count = int(X_rows.shape[0])
        if count:
            ops.hessian_accum(X_rows.contiguous(), H, mean, 0.0, count)
        mark("statistics")
        H, mean, count = sdist.allreduce_statistics(H, mean, count, self.group, self.pbuf, self._barrier)
        mark("allreduce")
        Hq = ops.remove_input_bias(H, mean) if self.bias_correction else H
        sc = _device_scaling(W_rows, self.cb, Hq, self.scaling_mode, self.grid_size, self.min_factor, self.max_factor)
        mark("scale_search")
        reduce_cols = ((lambda c: sdist.allreduce_column_sums(c, self.group))
                       if self.act_order in ("err", "sqerr") else None)
        st = obq.gptq_prepare(W_rows, Hq, self.cb, self.act_order, self.damp, self.nb_ls_moves,
                              colsum_reduce=reduce_cols, row_scale=sc, want_err=True)
        mark("order_permute")
        fac = None
        if st.chol_form:
            fac = self._factor_fn(Hq, st.order, st.dampval) if self.dist_factor else ops.chol_factor(Hq, st.order, st.dampval)
            self.info = fac[3]        # device int32 [1]: 0, or the first non-positive pivot (max over the ranks); obq.py:49-50
        mark("factor")
        q, (_, rows_err) = obq.gptq_finish(st, fac)
        mark("sweep")
        total_rows = torch.tensor([W_rows.shape[0]], dtype=torch.int64, device=dev)
        if self.world > 1:
            tdist.all_reduce(total_rows, group=self.group)
        err = sdist.allreduce_row_error_mean(rows_err, int(total_rows.item()), self.group)
        mark("error")
        if timing:
            torch.cuda.synchronize()
            self.phases_ms = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks[:-1], marks[1:])}
        return q, sc, err, Hq, mean


def quantize_layer_sharded(W_rows, X_rows, codebook, scaling_mode="diag", act_order="diag", damp=0.01, nb_ls_moves=0,
                           bias_correction=False, grid_size=100, min_factor=0.05, max_factor=1.0, group=None):
    """One-shot form of ShardedLayerQuantizer with the factor replicated on every rank (no peer
    workspace to set up); see the class for the distributed factor."""
    slq = ShardedLayerQuantizer(W_rows.shape[1], codebook, scaling_mode, act_order, damp, nb_ls_moves, bias_correction,
                                grid_size, min_factor, max_factor, group, dist_factor=False)
    return slq(W_rows, X_rows)
