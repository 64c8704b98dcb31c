"""Layer-set driver: quantize many independent layers with the per-layer hot path.

The reference's scripts loop over layers one by one (experiments/compare.py:50-135).  Layers
never interact, and small layers (OPT-125M: n = 768) are bound by the latency of their serial
chains (n dependent sweep columns, n/64 dependent factor panels), not by throughput.  This
driver therefore issues consecutive layers round-robin on a few CUDA streams so that one
layer's serial chain overlaps the others' -- same kernels, same results, no collective.
"""

from __future__ import annotations

import torch

from . import ops
from .scaling import quantize_scaled_device
from .statistics import _device_scaling


class LayerSetQuantizer:
    """Holds the side streams; call it with lists of device tensors."""

    def __init__(self, codebook, scaling_mode="diag", act_order="diag", damp=0.01, nb_ls_moves=0, grid_size=100,
                 min_factor=0.05, max_factor=1.0, streams=8):
        ops.require_cuda()
        self.cb = codebook
        self.scaling_mode, self.act_order = scaling_mode, act_order
        self.damp, self.nb_ls_moves = damp, nb_ls_moves
        self.grid_size, self.min_factor, self.max_factor = grid_size, min_factor, max_factor
        self.streams = [torch.cuda.Stream() for _ in range(max(1, int(streams)))]

    def _one(self, W, H):
        sc = _device_scaling(W, self.cb, H, self.scaling_mode, self.grid_size, self.min_factor, self.max_factor)
        q = quantize_scaled_device(W, sc, self.cb, H, self.act_order, self.damp, self.nb_ls_moves)
        err = ops.mean(ops.hweighted_error(W, q, H))
        return q, sc, err

    def __call__(self, Ws, Hs, errs_out=None, keep_outputs=True, _in_capture=False):
        """Ws[i] [r_i, n_i] fp32, Hs[i] [n_i, n_i] fp32 on the device.  Returns (quantized
        weights, scales, errors) -- errors as one fp32 device vector.  Nothing synchronises."""
        L = len(Ws)
        dev = Ws[0].device
        errs = errs_out if errs_out is not None else torch.empty(L, dtype=torch.float32, device=dev)
        main = torch.cuda.current_stream()
        start = torch.cuda.Event()
        start.record(main)
        outs, scales = [None] * L, [None] * L
        S = len(self.streams)
        for i in range(L):
            st = self.streams[i % S]
            if i < S:
                st.wait_event(start)
            with torch.cuda.stream(st):
                q, sc, e = self._one(Ws[i], Hs[i])
                errs[i:i + 1].copy_(e.reshape(1))
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
