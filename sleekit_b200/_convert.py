"""numpy <-> device plumbing for the drop-in API.

The reference's functions take and return host numpy arrays.  The drop-in
accepts those (copied to the current CUDA device, results copied back) and,
additionally, CUDA ``torch`` tensors, which pass through with no copy and no
synchronisation (results stay on the device).
"""

from __future__ import annotations

import numpy as np
import torch

from . import ops

_FLOATS = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}


def is_tensor(x):
    return isinstance(x, torch.Tensor)


def float_dtype_of(x):
    """numpy float dtype the reference would compute in for this input."""
    if is_tensor(x):
        return np.dtype(np.float64) if x.dtype == torch.float64 else np.dtype(np.float32)
    dt = np.asarray(x).dtype
    if dt == np.float64 or dt.kind in "iub":
        return np.dtype(np.float64)  # lists / ints behave as float64 under numpy
    return np.dtype(np.float32) if dt in (np.float32, np.float16) else np.dtype(np.float64)


H2D_BYTES = 0  # bytes copied host->device / device->host by this module (bench.py reads these)
D2H_BYTES = 0


def to_dev(x, dtype=None):
    """Contiguous CUDA tensor holding x (optionally cast)."""
    global H2D_BYTES
    dev = ops.device()
    if is_tensor(x):
        if not x.is_cuda:
            H2D_BYTES += x.numel() * x.element_size()
        t = x if x.is_cuda else x.to(dev)
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.contiguous()
    a = np.asarray(x)
    if a.dtype.kind not in "f" or a.dtype == np.float16:
        a = a.astype(np.float64 if a.dtype.kind in "iub" else np.float32)
    a = np.ascontiguousarray(a)
    H2D_BYTES += a.nbytes
    t = torch.from_numpy(a).to(dev, non_blocking=False)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t


def to_host(t):
    global D2H_BYTES
    D2H_BYTES += t.numel() * t.element_size()
    return t.detach().cpu().numpy()


def back(t, like):
    """Return t in the container kind of `like` (numpy in -> numpy out)."""
    return t if is_tensor(like) else to_host(t)


def torch_float(np_dtype):
    return _FLOATS[np.dtype(np_dtype)]
