"""Drop-in for ``sleekit.statistics`` (reference: sleekit/statistics.py).

``Sleekit`` keeps the GPTQ-compatible surface (``add_batch / quantize / export /
free`` and the three presets).  Statistics are accumulated and the layer is
quantized on the CUDA device by the sleekit_b200 kernels; a layer that lives on
the CPU gets its statistics kept on the current CUDA device and its quantized
weights written back to where the layer lives.
"""

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .codebook import UniformCodebook
from .scaling import obq_scale_device, quantize_scaled_device, search_scale_device


class Sleekit:
    """Statistics of a layer, with an API compatible with GPTQ (statistics.py:12-199)."""

    def __init__(self, layer):
        self.layer = layer
        weight = layer.weight
        if not isinstance(self.layer, (nn.Linear, nn.Conv1d, nn.Conv2d)):
            raise ValueError(f"Unsupported layer type {type(self.layer)}")
        if isinstance(self.layer, (nn.Conv1d, nn.Conv2d)):
            weight = weight.flatten(1)
        n = weight.shape[1]
        dev = self.compute_device
        self.mean = torch.zeros(n, dtype=torch.float32, device=dev)
        self.hessian = torch.zeros((n, n), dtype=torch.float32, device=dev)
        self.count = 0

    @property
    def device(self):
        return self.layer.weight.device

    @property
    def compute_device(self):
        w = self.layer.weight.device
        return w if w.type == "cuda" else ops.device()

    def _prepare_input(self, inp):
        """2-D [features, samples] view of a batch (statistics.py:37-74).  This is the GPTQ-compatible
        interface's own reshape / unfold sequence (Linear: flatten + transpose; Conv: F.unfold with the layer's
        kernel / dilation / padding / stride, then [features, batch * positions]); it is layer plumbing that
        stays torch's and necessarily mirrors the reference step by step -- the arithmetic that follows it
        (K1, ops.hessian_accum) is what this package replaces."""
        if isinstance(self.layer, nn.Linear):
            inp = inp.reshape((-1, inp.shape[-1]))
            inp = inp.t()
        elif isinstance(self.layer, nn.Conv2d):
            if inp.ndim == 3:
                inp = torch.unsqueeze(inp, 0)
            inp = F.unfold(inp, self.layer.kernel_size, self.layer.dilation, self.layer.padding, self.layer.stride)
            inp = inp.permute([1, 0, 2]).flatten(1)
        elif isinstance(self.layer, nn.Conv1d):
            if inp.ndim == 2:
                inp = torch.unsqueeze(inp, 0)
            inp = torch.unsqueeze(inp, -1)
            inp = F.unfold(inp, (self.layer.kernel_size[0], 1), (self.layer.dilation[0], 1),
                           (self.layer.padding[0], 0), (self.layer.stride[0], 1))
            inp = inp.permute([1, 0, 2]).flatten(1)
        else:
            raise RuntimeError(f"Unsupported layer type {type(self.layer)}")
        assert inp.ndim == 2
        return inp.float()

    def add_batch(self, inp, out=None):
        """Fold a batch into the running means of x and x x^T (statistics.py:76-87), kernel K1."""
        inp = self._prepare_input(inp)
        added = inp.shape[1]
        keep = self.count / (self.count + added)
        self.count += added
        # K1 wants samples as rows: [S, n] row-major (for nn.Linear this is the batch as given)
        x = inp.t().to(self.compute_device).contiguous()
        ops.hessian_accum(x, self.hessian, self.mean, keep, self.count)

    def export(self, path, npy_format=False):
        """Write bias / weight / mean / hessian as .pt or .npy (statistics.py:89-105)."""
        os.makedirs(path, exist_ok=True)
        items = {"bias": self.layer.bias, "weight": self.layer.weight, "mean": self.mean, "hessian": self.hessian}
        for name, t in items.items():
            t = t.detach().cpu()
            if npy_format:
                import numpy as np

                np.save(os.path.join(path, name + ".npy"), t.numpy())
            else:
                torch.save(t, os.path.join(path, name + ".pt"))

    def quantize_basic(self, nbits):
        """Plain GPTQ settings (statistics.py:107-118)."""
        return self.quantize(nbits, scaling_mode="mse", order_mode="diag", bias_correction=False, damp=0.01,
                             nb_ls_moves=0)

    def quantize_sleekit_light(self, nbits):
        """Sleekit "light" (statistics.py:120-131)."""
        return self.quantize(nbits, scaling_mode="diag", order_mode="sqerr", bias_correction=True, damp=0.03,
                             nb_ls_moves=0)

    def quantize_sleekit_heavy(self, nbits):
        """Sleekit "heavy" (statistics.py:133-144)."""
        return self.quantize(nbits, scaling_mode="hessian", order_mode="sqerr", bias_correction=True, damp=0.03,
                             nb_ls_moves=100)

    def quantize(self, nbits, scaling_mode="mse", order_mode="diag", bias_correction=False, damp=0.01,
                 nb_ls_moves=0, grid_size=100, min_factor=0.05, max_factor=1.0):
        """Quantize the layer in place (statistics.py:146-190); everything stays on the device."""
        cb = UniformCodebook(2**nbits, -1, 1)
        dev = self.compute_device
        H, mean = self.hessian, self.mean
        if bias_correction:
            H = ops.remove_input_bias(H, mean)
        weight = self.layer.weight.data.flatten(1).to(dev, torch.float32).contiguous()
        sc = _device_scaling(weight, cb, H, scaling_mode, grid_size, min_factor, max_factor)
        quant = quantize_scaled_device(weight, sc, cb, H, order_mode, damp, nb_ls_moves, check=True)
        self.layer.weight.data = quant.reshape(self.layer.weight.shape).to(self.layer.weight.device,
                                                                           self.layer.weight.dtype)
        if bias_correction:
            delta = ops.bias_delta(weight, quant, mean)
            self.layer.bias.data += delta.to(self.layer.bias.device, self.layer.bias.dtype)

    def free(self):
        """Drop the statistics (statistics.py:192-199)."""
        self.layer = None
        self.mean = None
        self.hessian = None
        self.count = 0


def _device_scaling(weight, cb, H, mode, grid_size, min_factor, max_factor):
    """compute_scaling (scaling.py:193-238) on device tensors, axis 0."""
    if mode == "max":
        return ops.row_noclip_scale(weight, float(cb.min()), float(cb.max()))
    if mode == "norm":
        return ops.row_rms_scale(weight)
    if mode == "obq":
        return obq_scale_device(weight, cb, H, min_factor=min_factor, max_factor=max_factor, grid_size=grid_size)[0]
    if mode == "mse":
        Hs = None
    elif mode.startswith("hessian"):
        Hs = H
        if len(mode) > 7:
            from .scaling import _add_to_diagonal

            Hs = _add_to_diagonal(H, 0.01 * float(mode[7:]))
    elif mode.startswith("diag"):
        from .scaling import _diag_with_penalty

        Hs = _diag_with_penalty(H, 0.01 * float(mode[4:]) if len(mode) > 4 else None)
    else:
        raise RuntimeError(f"Unknown scaling mode {mode}")
    return search_scale_device(weight, cb, Hs, min_factor, max_factor, grid_size)
