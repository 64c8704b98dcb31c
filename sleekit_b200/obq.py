"""Drop-in for ``sleekit.obq`` (reference: sleekit/obq.py), on CUDA kernels.

Public names, argument order, defaults, array layouts and error behaviour
follow the reference.  Host numpy arrays in -> host numpy arrays out; CUDA
torch tensors pass through and stay on the device (no synchronisation).
"""

import numpy as np
import torch

from . import _convert as cv
from . import ops

__all__ = [
    "np", "random_psd_matrix", "remove_input_bias", "remove_dead_values", "compute_hessian_chol",
    "compute_hessian_order", "channelwise_error", "quantization_error", "quantize_opt", "compute_gain",
    "LocalSearchQuantizer", "quantize_local_search", "_quantize_opt_core", "_quantize_opt_block",
]

MAX_LEAF = 32  # widest column block the leaf kernel sweeps (one lane per column)
# quantize_opt's fused path: Cholesky factor + R-form sweep (True) or full inverse factor + U-form sweep
import os as _os

USE_CHOL_FORM = _os.environ.get("SLK_CHOL_FORM", "1") != "0"
# layer error of the fused path from the sweep's own residuals (gptq_device, want_err); 0: always the K6 product
USE_SWEEP_ERROR = _os.environ.get("SLK_SWEEP_ERROR", "1") != "0"


def random_psd_matrix(size, rank, damp=0.0):
    """Wishart-distributed PSD test matrix, fp64 (obq.py:4-11).  Host RNG helper, not hot path."""
    factor = np.random.randn(size, rank).astype(np.float32)
    gram = factor @ factor.T
    return gram + damp * np.linalg.norm(gram, ord=2, axis=1) * np.eye(size)


def remove_input_bias(H, input_bias):
    """H - outer(m, m) (obq.py:14-25)."""
    assert H.ndim == 2
    assert input_bias.ndim == 1
    assert H.shape[0] == H.shape[1]
    assert H.shape[0] == input_bias.shape[0]
    dt = cv.torch_float(np.result_type(cv.float_dtype_of(H), cv.float_dtype_of(input_bias)))
    out = ops.remove_input_bias(cv.to_dev_cached(H, dt), cv.to_dev_cached(input_bias, dt))
    return cv.back(out, H)


def remove_dead_values(H, W):
    """In place: dead inputs (zero diagonal) get the mean diagonal and zero weights (obq.py:28-35).
    An O(n) mutation of caller-owned arrays, done where they live."""
    if cv.is_tensor(H):
        d = H.diagonal()
        fill = d.mean()
        dead = d == 0
        d[dead] = fill
        W[:, dead] = 0
        return
    d = H.diagonal()
    fill = d.mean()
    dead = d == 0
    H[dead, dead] = fill
    W[:, dead] = 0
    cv.invalidate(H)      # device copies of the arrays just patched are stale
    cv.invalidate(W)


def _raise_if_not_pd(info):
    bad = int(info.item())
    if bad != 0:
        raise np.linalg.LinAlgError("Matrix is not positive definite")


def compute_hessian_chol(H):
    """Upper factor U with inv(H) = U^T U, fp64 (obq.py:38-55)."""
    h = cv.to_dev_cached(H, torch.float64)
    assert h.ndim == 2 and h.shape[0] == h.shape[1]
    u64, _, info = ops.hinv(h, want64=True, want32=False)
    if not cv.is_tensor(H):
        _raise_if_not_pd(info)
    return cv.back(u64, H)


def _device_order(Wd, Hd, quantizer, act_order, diag64=None):
    """argsort keys of obq.py:58-86 for a device W (fp32) and the fp64 diagonal of H_opt."""
    n = Wd.shape[1]
    if act_order == "none":
        return torch.arange(n, dtype=torch.int64, device=Wd.device)
    d = diag64 if diag64 is not None else Hd.diagonal().to(torch.float64)
    if act_order == "diag":
        keys = -d
    elif act_order in ("err", "sqerr"):
        col = ops.col_resid_sums(Wd, quantizer, squared=(act_order == "sqerr"))
        keys = -d * col.to(torch.float64)
    elif act_order in ("inv_diag", "combined_diag"):
        u64, _, _ = ops.hinv(Hd.to(torch.float64).contiguous(), want64=True, want32=False)
        inv_diag = (u64 * u64).sum(dim=0)  # diag(U^T U)
        keys = inv_diag if act_order == "inv_diag" else -d / inv_diag
    elif act_order == "pivot":
        return ops.pivot_order(Hd.to(torch.float64).contiguous())       # obq.py:140-166
    else:
        raise RuntimeError(f"Invalid act_order value {act_order}")
    return ops.argsort(keys.contiguous())


def compute_hessian_order(W, H, quantizer, act_order):
    """Column ordering heuristics (obq.py:58-86); returns int64 indices."""
    if act_order not in ("err", "sqerr", "combined_diag", "inv_diag", "pivot", "diag", "none"):
        raise RuntimeError(f"Invalid act_order value {act_order}")
    Wd = cv.to_dev_cached(W, torch.float32)
    Hd = cv.to_dev_cached(H)
    order = _device_order(Wd, Hd, quantizer, act_order)
    return cv.back(order, W)


def channelwise_error(W, Q, H):
    """((W-Q) @ H * (W-Q)).sum(-1) per row (obq.py:89-95)."""
    dt_np = np.result_type(cv.float_dtype_of(W), cv.float_dtype_of(Q), cv.float_dtype_of(H))
    dt = cv.torch_float(dt_np)
    # the residual is formed in the dtype of W and Q, then promoted with H (numpy semantics)
    wq = cv.torch_float(np.result_type(cv.float_dtype_of(W), cv.float_dtype_of(Q)))
    Wd, Qd = cv.to_dev_cached(W, wq), cv.to_dev_cached(Q, wq)
    lead = Wd.shape[:-1]
    Wd, Qd = Wd.reshape(-1, Wd.shape[-1]), Qd.reshape(-1, Qd.shape[-1])
    Hd = cv.to_dev_cached(H, dt)
    if wq == dt:
        err = ops.hweighted_error(Wd, Qd, Hd)
    else:
        err = ops.hweighted_error((Wd - Qd).to(dt), None, Hd)
    return cv.back(err.reshape(lead), W)


def quantization_error(W, Q, H):
    """Mean of channelwise_error, a scalar of the promoted dtype (obq.py:98-103)."""
    if cv.is_tensor(W):
        return ops.mean(channelwise_error(W, Q, H).reshape(-1).contiguous())
    dt_np = np.result_type(cv.float_dtype_of(W), cv.float_dtype_of(Q), cv.float_dtype_of(H))
    rows = channelwise_error(cv.to_dev_cached(W), cv.to_dev_cached(Q), cv.to_dev_cached(H))
    m = ops.mean(rows.reshape(-1).contiguous())
    return np.dtype(dt_np).type(m.item())  # a numpy scalar, so f-strings print as the reference's do


def _sweep_leaf(min_block_size):
    return int(min(max(int(min_block_size), 1), MAX_LEAF))


def _quantize_opt_block(Q, E, Hinv, quantizer, min_block_size, num_blocks):
    """Blocked sweep, in place on Q and E (obq.py:121-137).  Leaves wider than 32 columns are
    split further (the result is the same algebra; the reference's own test_blockobq shows the
    blocking does not matter)."""
    if cv.is_tensor(Q):
        u64 = cv.to_dev_cached(Hinv, torch.float64)
        ops.gptq_sweep(Q, u64, u64.to(torch.float32), quantizer, _sweep_leaf(min_block_size), num_blocks, e=E,
                       exact_leaf=True)
        return
    q = cv.to_dev(Q, torch.float32)
    u64 = cv.to_dev_cached(Hinv, torch.float64)
    q, e = ops.gptq_sweep(q, u64, u64.to(torch.float32), quantizer, _sweep_leaf(min_block_size), num_blocks,
                          exact_leaf=True)
    Q[...] = cv.to_host(q)
    E[...] = cv.to_host(e)
    cv.invalidate(Q)
    cv.invalidate(E)


def _quantize_opt_core(Q, E, Hinv, quantizer):
    """Unblocked sweep, in place (obq.py:106-118)."""
    _quantize_opt_block(Q, E, Hinv, quantizer, MAX_LEAF, 8)


class GptqStage:
    """State between the two halves of gptq_device: everything up to the fp64 factor has been enqueued
    (damp value, ordering, scaled / permuted weights); the factor and the sweep follow in gptq_finish.
    The layer-set driver (pipeline.LayerSetQuantizer) runs the first half of many layers, factors
    their Hessians in batched launches (ops.chol_factor_batched) and then finishes each layer."""

    __slots__ = ("W_in", "Wd", "Hd", "quantizer", "act_order", "dampval", "order", "Q", "fuse", "row_scale",
                 "nb_ls_moves", "leaf", "num_blocks", "want_err", "chol_form", "info")


def gptq_prepare(Wd, Hd, quantizer, act_order="diag", damp=0.01, nb_ls_moves=0, min_block_size=32, num_blocks=8,
                 colsum_reduce=None, row_scale=None, want_err=False):
    """First half of gptq_device: obq.py:198-203 (damp, ordering keys, argsort, permutation)."""
    st = GptqStage()
    st.dampval = ops.damp_value(Hd, damp)                                # obq.py:198
    st.W_in, st.Hd, st.quantizer, st.act_order = Wd, Hd, quantizer, act_order
    st.row_scale, st.nb_ls_moves, st.want_err = row_scale, nb_ls_moves, want_err
    st.leaf, st.num_blocks = _sweep_leaf(min_block_size), num_blocks
    st.chol_form = st.leaf == MAX_LEAF and USE_CHOL_FORM
    st.fuse = row_scale is not None and act_order in ("diag", "none") and not nb_ls_moves
    if row_scale is not None and not st.fuse:
        Wd = ops.scale_rows(Wd, row_scale, 0)                            # scaling.py:73
    st.Wd = Wd
    if act_order == "none":
        order = None
    elif act_order in ("diag", "err", "sqerr"):
        col = None
        if act_order != "diag":
            col = ops.col_resid_sums(Wd, quantizer, squared=(act_order == "sqerr"))
            if colsum_reduce is not None:
                col = colsum_reduce(col)
        order = ops.argsort(ops.order_keys(Hd, st.dampval, col))         # obq.py:199
    else:
        hopt_diag = Hd.diagonal().to(torch.float64) + st.dampval.to(torch.float64)
        hopt = Hd.to(torch.float64)
        hopt.diagonal().add_(st.dampval.to(torch.float64))
        order = _device_order(Wd, hopt, quantizer, act_order, hopt_diag)
    st.order = order
    if st.fuse:
        st.Q = ops.scale_permute_cols(Wd, order, row_scale)               # scaling.py:73 + obq.py:202-203
    else:
        st.Q = ops.permute_cols(Wd, order) if order is not None else Wd.clone()  # obq.py:202-203
    return st


def gptq_finish(st, factor=None, check=False):
    """Second half of gptq_device: fp64 factor (unless `factor` = (r32, rt32, ud32, info) of
    (Hd + damp)[order][:, order] is handed in), sweep, un-permutation, local search, error."""
    Q, Hd, quantizer, order, row_scale = st.Q, st.Hd, st.quantizer, st.order, st.row_scale
    want_err, nb_ls_moves = st.want_err, st.nb_ls_moves
    sums = None
    if st.chol_form:
        # factor only (no triangular inverse): H_opt = R R^T, sweep from R
        r32, rt32, ud32, info = factor if factor is not None else ops.chol_factor(Hd, order, st.dampval)  # obq.py:204
        st.info = info                                   # first non-PD pivot (0 = fine): obq.py:49-50 raises there
        if want_err and not nb_ls_moves and USE_SWEEP_ERROR:
            sums = torch.empty((Q.shape[0], 2), dtype=torch.float32, device=Q.device)
        ops.gptq_sweep_r(Q, r32, rt32, ud32, quantizer, err_sums=sums)    # obq.py:208-209
        if sums is not None:
            err = ops.sweep_error(sums, row_scale, st.dampval, want_rows=True)
    else:
        assert factor is None
        u64, u32, info = ops.hinv(Hd, order, st.dampval)                  # obq.py:204-205
        st.info = info
        ops.gptq_sweep(Q, u64, u32, quantizer, st.leaf, st.num_blocks)    # obq.py:208-209
    if st.fuse:
        Q = ops.scale_permute_cols(Q, order, row_scale, scatter=True)     # obq.py:212-213 + scaling.py:80
        if check:
            _raise_if_not_pd(info)
        if want_err:
            return Q, (err if sums is not None else _k6_error(st.W_in, Q, Hd))
        return Q
    if order is not None:
        Q = ops.permute_cols(Q, order, scatter=True)                      # obq.py:212-213
    if check:
        _raise_if_not_pd(info)
    if nb_ls_moves:
        if want_err and USE_SWEEP_ERROR:
            # the search keeps p = (Q - W) H current: the rows' errors are p . (Q - W), no K6 product afterwards
            sums = torch.empty((Q.shape[0], 2), dtype=torch.float32, device=Q.device)
        ops.local_search(st.Wd, Q, Hd, quantizer, nb_ls_moves, err_sums=sums)   # obq.py:216
        if sums is not None:
            err = ops.sweep_error(sums, row_scale, None, want_rows=True)
    if row_scale is not None:
        Q = ops.scale_rows(Q, row_scale, 1)                               # scaling.py:80
    if want_err:
        return Q, (err if sums is not None else _k6_error(st.W_in, Q, Hd))
    return Q


def gptq_device(Wd, Hd, quantizer, act_order="diag", damp=0.01, nb_ls_moves=0, min_block_size=32, num_blocks=8,
                check=False, colsum_reduce=None, row_scale=None, want_err=False, factor_fn=None):
    """quantize_opt on device tensors (fp32 W [r,n], fp32 H [n,n]); returns quantized values [r,n].
    The whole chain -- damp, keys, argsort, gather, fp64 factor, sweep, scatter, local search --
    is enqueued on the current stream without a host round trip.  colsum_reduce: optional callable
    applied to the column residual sums of the err / sqerr orderings (row-sharded runs all-reduce
    them there, dist.allreduce_column_sums).  row_scale: Wd is then the UNSCALED matrix; the division
    by the row scales and the de-scaling of the result are fused into the two column-permutation
    passes (quantize_with_scaling's scaling.py:73 and :80) and the de-scaled weights are returned.
    want_err: also return (mean layer error [1], row errors [r]) = quantization_error / channelwise_error
    (obq.py:89-103) of the returned weights against Wd under Hd, taken from the sweep's residuals
    (sum E^2 - damp * sum (W-Q)^2, ops.sweep_error) when the factor-form sweep ran and no local-search
    move follows, from the local search's own p = (Q - W) H after its moves (p . (Q - W) per row), from the
    K6 product otherwise.  (The residual form subtracts two fp32 row sums: when
    delta H delta is smaller than ~1e-6 of damp * |delta|^2 -- a Hessian of much lower rank than n under a large
    damp -- it cancels; call channelwise_error for such inputs.  Measured agreement with the product on the
    rank-deficient bench layers: 1.4e-5.)  factor_fn(Hd, order, dampval) -> (r32, rt32, ud32,
    info): an alternative producer of the Cholesky factor (the multi-GPU factorisation)."""
    st = gptq_prepare(Wd, Hd, quantizer, act_order, damp, nb_ls_moves, min_block_size, num_blocks,
                      colsum_reduce, row_scale, want_err)
    factor = factor_fn(Hd, st.order, st.dampval) if (factor_fn is not None and st.chol_form) else None
    return gptq_finish(st, factor, check)


def _k6_error(Wd, Qd, Hd):
    rows = ops.hweighted_error(Wd, Qd, Hd)                                # obq.py:89-95
    return ops.mean(rows), rows                                           # obq.py:98-103


def quantize_opt(W, H, quantizer, act_order="diag", damp=0.01, nb_ls_moves=0, min_block_size=32, num_blocks=8):
    """GPTQ-like error-compensated quantization (obq.py:169-217)."""
    assert W.ndim == 2
    assert H.ndim == 2
    assert H.shape[0] == H.shape[1]
    assert H.shape[0] == W.shape[1]
    assert min_block_size >= 1
    Wd = cv.to_dev_cached(W, torch.float32)   # obq.py:195-196: both cast to fp32
    Hd = cv.to_dev_cached(H, torch.float32)
    Q = gptq_device(Wd, Hd, quantizer, act_order, damp, nb_ls_moves, min_block_size, num_blocks,
                    check=not cv.is_tensor(W))
    return cv.back(Q, W)


def compute_gain(W, Q, H, candidates):
    """Gain of moving each weight to its candidate (obq.py:220-231)."""
    dt = cv.torch_float(np.result_type(*(cv.float_dtype_of(a) for a in (W, Q, H, candidates))))
    out = ops.gain(cv.to_dev_cached(W, dt), cv.to_dev_cached(Q, dt), cv.to_dev_cached(H, dt), cv.to_dev_cached(candidates, dt))
    return cv.back(out, W)


class LocalSearchQuantizer:
    """Best-first single-weight flips (obq.py:234-346).  State lives on the device; the
    attributes the reference exposes (err, Q_up, Q_down, gain_up, gain_down) are derived from
    the current Q on access.  ``quantize_local_search`` runs all moves in one kernel launch.

    Contract of the device path (narrower than the reference outside the tested inputs): Q must lie on the
    codebook (the kernel keeps codes, so an off-grid entry is snapped to its nearest codeword), and H is
    used as its own transpose (a Hessian: symmetric) when (Q - W) H is formed."""

    def __init__(self, W, Q, H, quantizer):
        assert W.ndim == 2
        assert H.ndim == 2
        assert H.shape[0] == H.shape[1]
        assert H.shape[0] == W.shape[1]
        assert Q.shape == W.shape
        self._like = W
        self._W = cv.to_dev_cached(W, torch.float32)
        self._Q = cv.to_dev_cached(Q, torch.float32).clone()
        self._H = cv.to_dev_cached(H, torch.float32)
        self.quantizer = quantizer
        self._state = ops.LocalSearchState(self._W)      # keeps (Q - W) H across do_move calls

    @property
    def nchannels(self):
        return self._W.shape[0]

    @property
    def W(self):
        return cv.back(self._W, self._like)

    @property
    def Q(self):
        return cv.back(self._Q, self._like)

    @property
    def H(self):
        return cv.back(self._H, self._like)

    @property
    def err(self):
        return cv.back(ops.hweighted_error(self._W, self._Q, self._H), self._like)

    def _cand(self, mode):
        return ops.round_to_codebook(self._Q, self.quantizer, mode)[0]

    @property
    def Q_up(self):
        return cv.back(self._cand(ops.UP), self._like)

    @property
    def Q_down(self):
        return cv.back(self._cand(ops.DOWN), self._like)

    @property
    def gain_up(self):
        return cv.back(ops.gain(self._W, self._Q, self._H, self._cand(ops.UP)), self._like)

    @property
    def gain_down(self):
        return cv.back(ops.gain(self._W, self._Q, self._H, self._cand(ops.DOWN)), self._like)

    def do_move(self):
        ops.local_search_step(self._W, self._Q, self._H, self.quantizer, 1, self._state)


def quantize_local_search(W, Q, H, quantizer, nb_moves):
    """nb_moves best-first flips per row (obq.py:349-358)."""
    if nb_moves == 0:
        return Q
    Wd = cv.to_dev_cached(W, torch.float32)
    Qd = cv.to_dev_cached(Q, torch.float32).clone()
    Hd = cv.to_dev_cached(H, torch.float32)
    ops.local_search(Wd, Qd, Hd, quantizer, nb_moves)
    return cv.back(Qd, W)
