"""Drop-in for ``sleekit.scaling`` (reference: sleekit/scaling.py), on CUDA kernels."""

import numpy as np
import torch

from . import _convert as cv
from . import ops
from .obq import (  # noqa: F401  (the reference re-exports these names through its imports)
    _quantize_opt_block,
    compute_hessian_chol,
    compute_hessian_order,
    quantize_opt,
    gptq_device,
    _device_order,
    _sweep_leaf,
)

__all__ = [
    "np", "apply_scaling", "apply_scaling_in_place", "compute_norm_scaling", "compute_non_saturating_scaling",
    "quantize_with_scaling", "compute_min_mse_scaling", "compute_obq_scaling", "compute_scaling",
    "_broadcast_scaling", "_compute_mse", "_quantize_opt_block", "compute_hessian_chol",
    "compute_hessian_order", "quantize_opt",
]


def _broadcast_scaling(data, scale, axis):
    """Shape a 1-D scale so it broadcasts along ``axis`` (scaling.py:11-18)."""
    assert scale.ndim == 1
    shape = [1] * data.ndim
    shape[axis] = -1
    return scale.reshape(shape)


def _axis_view(shape, axis):
    axis = axis % len(shape)
    outer = int(np.prod(shape[:axis], dtype=np.int64))
    inner = int(np.prod(shape[axis + 1:], dtype=np.int64))
    return outer, int(shape[axis]), inner


def _scale_dev(x, s, axis, mode):
    outer, length, inner = _axis_view(tuple(x.shape), axis)
    assert s.ndim == 1 and s.numel() == length
    return ops.scale_axis(x, s, outer, length, inner, mode)


def apply_scaling(data, scale, axis=0):
    """data / scale broadcast along ``axis`` (scaling.py:21-25)."""
    assert scale.ndim == 1
    dt = cv.torch_float(np.result_type(cv.float_dtype_of(data), cv.float_dtype_of(scale)))
    out = _scale_dev(cv.to_dev_cached(data, dt), cv.to_dev_cached(scale, dt), axis, 0)
    return cv.back(out, data)


def apply_scaling_in_place(data, scale, axis=0):
    """In-place variant (scaling.py:28-32): the result is written back into ``data``."""
    res = apply_scaling(data, scale, axis)
    if cv.is_tensor(data):
        data.copy_(res)
    else:
        data[...] = res
        cv.invalidate(data)


def _rows_first(x, axis):
    """[axis, everything else] contiguous 2-D view of x (scaling.py:39, 50, 119-120)."""
    axis = axis % x.ndim
    if axis != 0:
        x = x.movedim(axis, 0)
    return x.reshape(x.shape[0], -1).contiguous()


def compute_norm_scaling(data, axis=0):
    """sqrt(max(mean(x^2) over the other axes, 1e-16)) (scaling.py:35-41)."""
    x = cv.to_dev_cached(data)
    return cv.back(ops.row_rms_scale(_rows_first(x, axis)), data)


def compute_non_saturating_scaling(data, codebook, axis=0):
    """Smallest scale with no clipping (scaling.py:44-55)."""
    if codebook.min() >= 0 or codebook.max() <= 0:
        raise RuntimeError("Codebook should have both negative and positive values.")
    x = cv.to_dev_cached(data)
    out = ops.row_noclip_scale(_rows_first(x, axis), float(codebook.min()), float(codebook.max()))
    return cv.back(out, data)


def quantize_scaled_device(Wd, sd, quantizer, Hd=None, act_order="diag", damp=0.01, nb_ls_moves=0, check=False,
                           colsum_reduce=None, want_err=False, factor_fn=None):
    """quantize_with_scaling on device tensors; returns de-scaled quantized weights (with want_err and
    a Hessian: (weights, (mean layer error, row errors)), see gptq_device)."""
    if Hd is not None:
        # scaling.py:73, 75-77, 80: the two row scalings ride on the column permutations
        return gptq_device(Wd, Hd, quantizer, act_order, damp, nb_ls_moves, check=check,
                           colsum_reduce=colsum_reduce, row_scale=sd, want_err=want_err, factor_fn=factor_fn)
    assert not want_err
    x = ops.scale_rows(Wd, sd, 0)                                          # scaling.py:73
    x = ops.round_to_codebook(x, quantizer)[0]                              # scaling.py:79
    return ops.scale_rows(x, sd, 1)                                         # scaling.py:80


def quantize_with_scaling(data, scale, quantizer, H=None, act_order="diag", damp=0.01, nb_ls_moves=0):
    """Scale rows, quantize (GPTQ when H is given), scale back (scaling.py:58-81).
    Returns de-scaled weights, never codes."""
    assert data.ndim == 2
    assert scale.ndim == 1
    assert data.shape[0] == (scale.numel() if cv.is_tensor(scale) else scale.size)
    Wd = cv.to_dev_cached(data, torch.float32)
    sd = cv.to_dev_cached(scale, torch.float32)
    Hd = cv.to_dev_cached(H, torch.float32) if H is not None else None
    out = quantize_scaled_device(Wd, sd, quantizer, Hd, act_order, damp, nb_ls_moves, check=not cv.is_tensor(data))
    return cv.back(out, data)


def _compute_mse(H, E):
    """Row errors for no / diagonal / full Hessian (scaling.py:84-95): sum E^2, sum h_j E_j^2 or
    ((E @ H) * E).sum(-1), in the promoted dtype of E and H."""
    if H is None:
        Ed = cv.to_dev_cached(E)
        out = ops.row_wsq(Ed.reshape(-1, Ed.shape[-1]).contiguous()).reshape(Ed.shape[:-1])
    elif H.ndim == 1:
        assert E.shape[1] == H.shape[0]
        out = ops.row_wsq(cv.to_dev_cached(E), cv.to_dev_cached(H))     # dtype promotion inside (np.square(E) in E's dtype)
    else:
        assert H.ndim == 2 and E.shape[1] == H.shape[0] and H.shape[1] == H.shape[0]
        dt = cv.torch_float(np.result_type(cv.float_dtype_of(E), cv.float_dtype_of(H)))
        out = ops.hweighted_error(cv.to_dev_cached(E, dt), None, cv.to_dev_cached(H, dt))
    return cv.back(out, E)


_FACTOR_CACHE = {}


def _factors(min_factor, max_factor, grid_size, dev):
    """The fp32 grid of scaling.py:124 on the device (cached: a few hundred bytes per setting, and
    no host->device copy is then needed inside a CUDA-graph capture)."""
    key = (float(min_factor), float(max_factor), int(grid_size), str(dev))
    t = _FACTOR_CACHE.get(key)
    if t is None:
        grid = np.linspace(min_factor, max_factor, grid_size, dtype=np.float32)  # scaling.py:124
        t = _FACTOR_CACHE[key] = torch.from_numpy(grid).to(dev)
    return t


FULLH_UNCERTIFIED = None   # optional int32 CUDA tensor [1]: the full-H searches add the rows they could not certify


def search_scale_device(Wd, codebook, Hd=None, min_factor=0.05, max_factor=1.0, grid_size=100):
    """compute_min_mse_scaling for a device [r, n] fp32 matrix, axis 0."""
    f = _factors(min_factor, max_factor, grid_size, Wd.device)
    if Hd is None or Hd.ndim == 1:
        return ops.scale_search(Wd, codebook, f, Hd)[0]
    if FULLH_UNCERTIFIED is not None:
        sc, _, bad = ops.scale_search_fullh(Wd, codebook, f, Hd, want_check=True)
        FULLH_UNCERTIFIED.add_(bad.clamp_min(0))
        return sc
    return ops.scale_search_fullh(Wd, codebook, f, Hd)[0]


def compute_min_mse_scaling(data, codebook, axis=0, H=None, min_factor=0.05, max_factor=1.0, grid_size=100):
    """Grid search of the per-row scale minimising the (H-weighted) squared error (scaling.py:98-134)."""
    if codebook.min() >= 0 or codebook.max() <= 0:
        raise RuntimeError("Codebook should have both negative and positive values.")
    flat = _rows_first(cv.to_dev_cached(data, torch.float32), axis)
    Hd = None
    if H is not None:
        Hd = cv.to_dev_cached(H)
        if Hd.dtype not in (torch.float32, torch.float64):
            Hd = Hd.float()
        assert Hd.shape[-1] == flat.shape[1]
    out = search_scale_device(flat, codebook, Hd, min_factor, max_factor, grid_size)
    return cv.back(out, data)


def obq_scale_device(Wd, codebook, Hd, damp=0.01, act_order="diag", min_factor=0.05, max_factor=1.0, grid_size=100,
                     max_rows_per_launch=1 << 16):
    """compute_obq_scaling on device tensors: every grid point is a full sweep sharing one
    ordering and one factor, i.e. the sweep kernel over (grid points x rows) independent rows."""
    r, n = Wd.shape
    base = ops.row_noclip_scale(Wd, float(codebook.min()), float(codebook.max()))   # scaling.py:164
    dampval = ops.damp_value(Hd, damp)                                                # scaling.py:167
    ws = ops.scale_rows(Wd, base, 0)
    if act_order == "none":
        order = torch.arange(n, dtype=torch.int64, device=Wd.device)
    elif act_order in ("diag", "err", "sqerr"):
        col = ops.col_resid_sums(ws, codebook, squared=(act_order == "sqerr")) if act_order != "diag" else None
        order = ops.argsort(ops.order_keys(Hd, dampval, col))                         # scaling.py:168-170
    else:
        hopt = Hd.to(torch.float64)
        hopt.diagonal().add_(dampval.to(torch.float64))
        order = _device_order(ws, hopt, codebook, act_order)
    Wp = ops.permute_cols(Wd, order)                                                  # scaling.py:171
    Hp = Hd[order][:, order].contiguous()                                             # scaling.py:172 (plumbing gather)
    from . import obq as _obq

    chol_form = _obq.USE_CHOL_FORM
    if chol_form:
        r32, rt, ud32, info = ops.chol_factor(Hd, order, dampval)                     # scaling.py:173-174 (factor only)
    else:
        u64, u32, info = ops.hinv(Hd, order, dampval)                                 # scaling.py:173-174
    f = _factors(min_factor, max_factor, grid_size, Wd.device)
    best_err = torch.full((r,), float("inf"), dtype=torch.float32, device=Wd.device)
    best_f = torch.full((r,), float("inf"), dtype=torch.float32, device=Wd.device)
    chunk = max(1, min(grid_size, max_rows_per_launch // max(r, 1)))
    for g0 in range(0, grid_size, chunk):
        fs = f[g0:g0 + chunk]
        gc = fs.numel()
        scale = (fs.unsqueeze(1) * base.unsqueeze(0)).reshape(-1).contiguous()        # scaling.py:181
        Wrep = Wp.unsqueeze(0).expand(gc, r, n).reshape(gc * r, n).contiguous()
        Q = ops.scale_rows(Wrep, scale, 0)                                            # scaling.py:182
        if chol_form:
            ops.gptq_sweep_r(Q, r32, rt, ud32, codebook)                              # scaling.py:183-184
        else:
            ops.gptq_sweep(Q, u64, u32, codebook, 32, 8)
        Q = ops.scale_rows(Q, scale, 1)                                               # scaling.py:185
        err = ops.hweighted_error(Q, Wrep, Hp).reshape(gc, r)                         # scaling.py:186
        for k in range(gc):                                                           # scaling.py:187-189
            better = err[k] < best_err
            best_err = torch.where(better, err[k], best_err)
            best_f = torch.where(better, fs[k], best_f)
    return base * best_f, info


def compute_obq_scaling(data, codebook, axis, H, damp=0.01, act_order="diag", min_factor=0.05, max_factor=1.0,
                        grid_size=100):
    """Scale search where every grid point is evaluated after a full GPTQ sweep (scaling.py:137-190)."""
    if codebook.min() >= 0 or codebook.max() <= 0:
        raise RuntimeError("Codebook should have both negative and positive values.")
    Wd = _rows_first(cv.to_dev_cached(data, torch.float32), axis)
    Hd = cv.to_dev_cached(H, torch.float32)
    out, info = obq_scale_device(Wd, codebook, Hd, damp, act_order, min_factor, max_factor, grid_size)
    if not cv.is_tensor(data):
        from .obq import _raise_if_not_pd

        _raise_if_not_pd(info)
    return cv.back(out, data)


def compute_scaling(data, codebook, H, mode="mse", axis=0, min_factor=0.05, max_factor=1.0, grid_size=100):
    """Dispatcher over the scaling heuristics (scaling.py:193-238)."""
    if mode == "max":
        return compute_non_saturating_scaling(data, codebook, axis)
    if mode == "norm":
        return compute_norm_scaling(data, axis)
    if mode == "obq":
        return compute_obq_scaling(data, codebook, axis, H=H, grid_size=grid_size, min_factor=min_factor,
                                   max_factor=max_factor)
    if mode == "mse":
        Hs = None
    elif mode.startswith("hessian"):
        Hs = H
        if len(mode) > 7:
            penalty = 0.01 * float(mode[7:])
            Hs = _add_to_diagonal(H, penalty)
    elif mode.startswith("diag"):
        Hs = _diag_with_penalty(H, 0.01 * float(mode[4:]) if len(mode) > 4 else None)
    else:
        raise RuntimeError(f"Unknown scaling mode {mode}")
    return compute_min_mse_scaling(data, codebook, axis, H=Hs, grid_size=grid_size, min_factor=min_factor,
                                   max_factor=max_factor)


def _add_to_diagonal(H, penalty):
    """H + penalty * mean(diag H) * eye, promoted to fp64 as np.eye does (scaling.py:222)."""
    Hd = cv.to_dev_cached(H)
    mean_diag = Hd.diagonal().mean()          # stays in H's dtype
    out = Hd.to(torch.float64).clone()
    out.diagonal().add_((penalty * mean_diag).to(Hd.dtype).to(torch.float64))
    return out


def _diag_with_penalty(H, penalty):
    """H.diagonal() [+ penalty * its mean], in H's dtype (scaling.py:224-227)."""
    d = cv.to_dev_cached(H).diagonal().contiguous()
    if penalty is not None:
        d = d + (penalty * d.mean()).to(d.dtype)
    return d
