"""CPU oracle for the layer-wise quantization hot path of Coloquinte/sleekit.

TEST INFRASTRUCTURE ONLY.  This file is a numpy restatement of the reference's
algorithm, written so that every arithmetic step happens in the same dtype and
the same order as the reference executes it under numpy >= 2 (NEP 50 scalar
promotion).  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of
``bench.py`` (``cpu_baseline`` and ``--impl reference``) may import it.  The
product package ``sleekit_b200`` never does.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the unmodified
reference (imported from /root/reference in the build container) on seeded
inputs and commits inputs+outputs as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against them, and
against the known-answer vectors of the reference's own tests
(tests/test_codebook.py:6-32, tests/test_scaling.py:16-41,56-72).

Citations ``ref:`` are relative to /root/reference/sleekit/.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

# --------------------------------------------------------------------------
# Codebooks (ref: codebook.py:4-95 UniformCodebook, codebook.py:98-188 Codebook)
# --------------------------------------------------------------------------


@dataclass
class UniformGrid:
    """Evenly spaced codewords lo .. hi (ref: codebook.py:4-41)."""

    n: int
    lo: float
    hi: float

    def __post_init__(self):
        self.n = int(self.n)
        assert self.lo < self.hi and self.n >= 2

    def __len__(self):
        return self.n

    def min(self):
        return self.lo

    def max(self):
        return self.hi

    @property
    def step(self):
        # ref: codebook.py:35-37 -- a Python float; it is rounded to the data's
        # dtype when it meets the array (NEP 50 weak scalar).
        return (self.hi - self.lo) / (self.n - 1)

    def _slots(self, x, shift, lo_slot, hi_slot):
        # ref: codebook.py:47-49, 60-62, 71-74, 83-86
        t = x - self.lo
        t /= self.step
        if shift:
            t += shift
        return t.round().clip(lo_slot, hi_slot)

    def index(self, x):
        # ref: codebook.py:43-54
        k = self._slots(x, 0, 0, self.n - 1)
        return k.astype(_index_dtype(self.n))

    def _to_value(self, k):
        # ref: codebook.py:63-65 -- multiply, then add: two roundings, no FMA
        k *= self.step
        k += self.lo
        return k

    def value(self, x):
        return self._to_value(self._slots(x, 0, 0, self.n - 1))

    def up(self, x):
        # ref: codebook.py:67-77
        return self._to_value(self._slots(x, 1, 1, self.n - 1))

    def down(self, x):
        # ref: codebook.py:79-89
        return self._to_value(self._slots(x, -1, 0, self.n - 2))

    __call__ = value
    quantize_up = up
    quantize_down = down
    quantize_index = index


def _index_dtype(n):
    # ref: codebook.py:50-54, 156-160
    if n <= 2**8:
        return np.uint8
    if n <= 2**16:
        return np.uint16
    return np.uint32


class TableGrid:
    """Sorted fp32 codewords with bin limits (ref: codebook.py:103-113)."""

    def __init__(self, values, limits=None):
        self.values = np.array(values, dtype=np.float32)
        if limits is None:
            self.values.sort()
            self.limits = (self.values[:-1] + self.values[1:]) / 2
        else:
            self.limits = np.array(limits, dtype=np.float32)

    def __len__(self):
        return len(self.values)

    def min(self):
        return self.values[0]

    def max(self):
        return self.values[-1]

    def _bin(self, x):
        # ref: codebook.py:155, 172, 180 -- x == limit goes to the upper bin
        return np.digitize(x, self.limits)

    def index(self, x):
        return self._bin(x).astype(_index_dtype(len(self)))

    def value(self, x):
        # ref: codebook.py:162-166
        return self.values[self.index(x)]

    def up(self, x):
        # ref: codebook.py:168-174
        nxt = np.concatenate((self.values[1:], self.values[-1:]))
        return nxt[self._bin(x)]

    def down(self, x):
        # ref: codebook.py:176-182
        prv = np.concatenate((self.values[:1], self.values[:-1]))
        return prv[self._bin(x)]

    __call__ = value
    quantize_up = up
    quantize_down = down
    quantize_index = index


# --------------------------------------------------------------------------
# Statistics (ref: statistics.py:76-87)
# --------------------------------------------------------------------------


class RunningStats:
    """Running mean of x and of x x^T over calibration rows, fp32."""

    def __init__(self, n):
        self.mean = np.zeros(n, dtype=np.float32)
        self.hessian = np.zeros((n, n), dtype=np.float32)
        self.count = 0

    def add_rows(self, X):
        """X is [S, n] (rows are samples) -- the reference sees its transpose."""
        Xt = np.ascontiguousarray(X.reshape(-1, X.shape[-1]).T, dtype=np.float32)
        added = Xt.shape[1]
        keep = self.count / (self.count + added)  # ref: statistics.py:82
        self.count += added
        self.mean *= keep
        self.hessian *= keep
        self.mean += Xt.sum(axis=1) / self.count  # ref: statistics.py:86
        self.hessian += Xt @ Xt.T / self.count  # ref: statistics.py:87


def strip_input_bias(H, m):
    # ref: obq.py:14-25
    assert H.ndim == 2 and m.ndim == 1 and H.shape == (m.shape[0], m.shape[0])
    return H - np.outer(m, m)


def patch_dead_inputs(H, W):
    # ref: obq.py:28-35 (in place)
    d = H.diagonal()
    fill = d.mean()
    dead = d == 0
    H[dead, dead] = fill
    W[:, dead] = 0


# --------------------------------------------------------------------------
# Inverse factor, ordering, errors (ref: obq.py:38-103)
# --------------------------------------------------------------------------


def inverse_upper_factor(H):
    """Upper U with inv(H) = U^T U (ref: obq.py:38-55)."""
    back = H[::-1, ::-1]
    low = np.linalg.cholesky(back)
    low_inv = np.linalg.inv(low)
    return np.ascontiguousarray(low_inv[::-1, ::-1])


def greedy_pivot_order(H):
    """Greedy pivoted-Cholesky ordering: at every step the remaining variable with the largest
    conditional variance comes next (ref: obq.py:140-166)."""
    n = H.shape[0]
    M = np.array(H, dtype=np.float64, copy=True)
    perm = np.arange(n)
    for step in range(n):
        best = step + int(np.argmax(np.abs(M.diagonal()[step:])))
        M[[step, best], :] = M[[best, step], :]
        M[:, [step, best]] = M[:, [best, step]]
        perm[[step, best]] = perm[[best, step]]
        row = M[step, step + 1:]
        M[step + 1:, step + 1:] -= np.outer(row, row) / M[step, step]
    return perm


def column_order(W, H, grid, rule):
    """ref: obq.py:58-86 (the four rules the hot path uses)."""
    d = H.diagonal()
    if rule == "diag":
        return (-d).argsort()
    if rule == "none":
        return np.arange(W.shape[1])
    if rule in ("err", "sqerr"):
        resid = grid(W) - W
        col = np.abs(resid).sum(axis=0) if rule == "err" else np.square(resid).sum(axis=0)
        return (-d * col).argsort()
    if rule == "inv_diag":
        return np.linalg.inv(H).diagonal().argsort()
    if rule == "combined_diag":
        return (-d / np.linalg.inv(H).diagonal()).argsort()
    if rule == "pivot":
        return greedy_pivot_order(H)
    raise RuntimeError(f"Invalid act_order value {rule}")


def rowwise_error(W, Q, H):
    # ref: obq.py:89-95
    D = W - Q
    return ((D @ H) * D).sum(axis=-1)


def mean_error(W, Q, H):
    # ref: obq.py:98-103
    return rowwise_error(W, Q, H).mean()


# --------------------------------------------------------------------------
# GPTQ / OBQ sweep (ref: obq.py:106-137, 169-217)
# --------------------------------------------------------------------------


def sweep_plan(n, leaf=32, fanout=8):
    """Flatten the reference's recursion (ref: obq.py:121-137) into a list of
    ("leaf", a, b) and ("gemm", a, b, end) steps over absolute column ranges.
    "gemm" means Q[:, b:end] -= E[:, a:b] @ U[a:b, b:end]."""
    steps = []

    def walk(a, b):
        size = b - a
        if size <= leaf:
            steps.append(("leaf", a, b))
            return
        width = max((size + fanout - 1) // fanout, leaf)
        for s in range(a, b, width):
            e = min(s + width, b)
            walk(s, e)
            steps.append(("gemm", s, e, b))

    walk(0, n)
    return steps


def sweep_in_place(Q, E, U, grid, leaf=32, fanout=8):
    """Run the blocked sweep on Q (fp32), filling E (fp32).  U is the fp64
    factor.  Arithmetic as executed by the reference under numpy >= 2:
    the scaled residual and every propagation are fp64, each store rounds to
    fp32 (ref: obq.py:110-118, 137)."""
    for st in sweep_plan(Q.shape[1], leaf, fanout):
        if st[0] == "leaf":
            _, a, b = st
            for i in range(a, b):
                w = Q[:, i]
                q = grid(w)
                r = (w - q) / U[i, i]
                E[:, i] = r
                Q[:, i] = q
                Q[:, i + 1 : b] -= np.outer(r, U[i, i + 1 : b])
        else:
            _, a, b, end = st
            # an empty trailing range is a no-op, exactly as in the reference
            Q[:, b:end] -= E[:, a:b] @ U[a:b, b:end]


def gptq(W, H, grid, rule="diag", damp=0.01, ls_moves=0, leaf=32, fanout=8):
    """ref: obq.py:169-217."""
    assert W.ndim == 2 and H.ndim == 2 and H.shape == (W.shape[1], W.shape[1])
    W = W.astype(np.float32)
    H = H.astype(np.float32)
    Hd = H + damp * H.diagonal().mean() * np.eye(H.shape[0])  # fp64 via np.eye
    perm = column_order(W, Hd, grid, rule)
    Wp = W[:, perm]
    Q = Wp.copy()
    U = inverse_upper_factor(Hd[perm][:, perm])
    E = np.zeros_like(Wp)
    sweep_in_place(Q, E, U, grid, leaf, fanout)
    back = np.argsort(perm)
    Q = Q[:, back]
    return local_search(Wp[:, back], Q, H, grid, ls_moves)


# --------------------------------------------------------------------------
# Best-first local search (ref: obq.py:220-358)
# --------------------------------------------------------------------------


def flip_gain(W, Q, H, cand):
    # ref: obq.py:220-231
    resid = Q - W
    step = cand - Q
    return -np.square(step) * H.diagonal() - 2 * (resid @ H) * step


class LocalSearch:
    """State and update rules of the reference (ref: obq.py:234-346): gains are
    patched incrementally after every accepted flip, never recomputed."""

    def __init__(self, W, Q, H, grid):
        assert W.ndim == 2 and H.ndim == 2
        assert H.shape == (W.shape[1], W.shape[1]) and Q.shape == W.shape
        self.W, self.Q, self.H, self.grid = W, Q.copy(), H, grid
        self.err = rowwise_error(W, self.Q, H)
        self.Q_up = grid.quantize_up(self.Q)
        self.Q_down = grid.quantize_down(self.Q)
        self.gain_up = flip_gain(W, self.Q, H, self.Q_up)
        self.gain_down = flip_gain(W, self.Q, H, self.Q_down)

    def _patch(self, gains, rows, cols, q_old, c_old, cands):
        # ref: obq.py:299-336
        k = np.arange(len(rows))
        H = self.H
        Wr = self.W[rows].copy()
        q_new_full = self.Q[rows].copy()
        q_old_full = q_new_full.copy()
        q_old_full[k, cols] = q_old
        c_new_full = cands[rows].copy()
        d_new_full = c_new_full - q_new_full
        hrows = H[cols].copy()
        c_new, q_new = c_new_full[k, cols], q_new_full[k, cols]
        d_old, d_new = c_old - q_old, c_new - q_new
        hd = H.diagonal()[cols]
        gains[rows, cols] += hd * (np.square(d_old) - np.square(d_new))
        gains[rows, cols] += 2 * ((q_old_full - Wr) * hrows).sum(axis=-1) * (d_old - d_new)
        gains[rows] += 2 * np.expand_dims(q_old - q_new, 1) * hrows * d_new_full

    def _apply(self, gains, pick, cands):
        # ref: obq.py:264-297
        rows = np.arange(self.W.shape[0])[pick]
        top = gains.max(axis=1)[pick]
        cols = gains.argmax(axis=1)[pick]
        new = cands[rows, cols]
        old = self.Q[rows, cols].copy()
        self.Q[rows, cols] = new
        old_up = self.Q_up[rows, cols].copy()
        self.Q_up[rows, cols] = self.grid.quantize_up(new)
        old_down = self.Q_down[rows, cols].copy()
        self.Q_down[rows, cols] = self.grid.quantize_down(new)
        self.err[rows] -= top
        self._patch(self.gain_up, rows, cols, old, old_up, self.Q_up)
        self._patch(self.gain_down, rows, cols, old, old_down, self.Q_down)

    def move(self):
        # ref: obq.py:338-346
        best_up = self.gain_up.max(axis=1)
        best_down = self.gain_down.max(axis=1)
        go_up = (best_up > best_down) & (best_up > 0)
        go_down = ~go_up & (best_down > 0)
        self._apply(self.gain_up, go_up, self.Q_up)
        self._apply(self.gain_down, go_down, self.Q_down)


def local_search(W, Q, H, grid, moves):
    # ref: obq.py:349-358
    if moves == 0:
        return Q
    ls = LocalSearch(W, Q, H, grid)
    for _ in range(moves):
        ls.move()
    return ls.Q


# --------------------------------------------------------------------------
# Scaling (ref: scaling.py:11-238)
# --------------------------------------------------------------------------


def _along(data, s, axis):
    # ref: scaling.py:11-18
    assert s.ndim == 1
    shape = [1] * data.ndim
    shape[axis] = -1
    return s.reshape(shape)


def divide_rows(data, s, axis=0):
    # ref: scaling.py:21-25
    return data / _along(data, s, axis)


def rms_scale(data, axis=0):
    # ref: scaling.py:35-41
    rest = tuple(i for i in range(data.ndim) if i != axis)
    return np.sqrt(np.maximum(np.square(data).mean(axis=rest), 1.0e-16))


def no_clip_scale(data, grid, axis=0):
    # ref: scaling.py:44-55
    if grid.min() >= 0 or grid.max() <= 0:
        raise RuntimeError("Codebook should have both negative and positive values.")
    rest = tuple(i for i in range(data.ndim) if i != axis)
    lo, hi = data.min(axis=rest), data.max(axis=rest)
    s = np.maximum(hi / grid.max(), lo / grid.min())
    return np.maximum(s, np.float32(1.0e-16))


def quantize_scaled(data, s, grid, H=None, rule="diag", damp=0.01, ls_moves=0):
    # ref: scaling.py:58-81 -- returns de-scaled weights, never codes
    assert data.ndim == 2 and s.ndim == 1 and data.shape[0] == s.size
    x = divide_rows(data, s, 0)
    x = gptq(x, H, grid, rule=rule, damp=damp, ls_moves=ls_moves) if H is not None else grid(x)
    return divide_rows(x, 1 / s, 0)


def weighted_sq_error(H, D):
    # ref: scaling.py:84-95
    if H is None:
        return np.square(D).sum(axis=1)
    if H.ndim == 1:
        assert D.shape[1] == H.shape[0]
        return (np.expand_dims(H, 0) * np.square(D)).sum(axis=1)
    assert H.ndim == 2 and H.shape == (D.shape[1], D.shape[1])
    return ((D @ H) * D).sum(axis=-1)


def _grid_argmin(base, factors, evaluate):
    # ref: scaling.py:125-134, 178-190 -- strict '<', first factor wins ties
    pick = np.full(base.size, np.inf, dtype=np.float32)
    best = np.full(base.size, np.inf, dtype=np.float32)
    for f in factors:
        e = evaluate(f * base)
        win = e < best
        best[win] = e[win]
        pick[win] = f
    return base * pick


def search_scale(data, grid, axis=0, H=None, lo=0.05, hi=1.0, points=100):
    # ref: scaling.py:98-134
    rest = tuple(i for i in range(data.ndim) if i != axis)
    flat = np.transpose(data, [axis, *rest])
    base = no_clip_scale(flat, grid, 0)
    factors = np.linspace(lo, hi, points, dtype=np.float32)
    return _grid_argmin(
        base, factors, lambda s: weighted_sq_error(H, quantize_scaled(flat, s, grid) - flat)
    )


def gptq_scale_evaluator(data, grid, axis, H, damp=0.01, rule="diag"):
    # ref: scaling.py:137-177 -- (base scales, evaluate) where evaluate(scales) is the per-row error of
    # one sweep at those scales, with the ordering and the factor shared by all grid points
    rest = tuple(i for i in range(data.ndim) if i != axis)
    W = np.transpose(data, [axis, *rest])
    base = no_clip_scale(W, grid, 0)
    Hd = H + damp * H.diagonal().mean() * np.eye(H.shape[0])
    perm = column_order(divide_rows(W, base, 0), Hd, grid, rule)
    W = W[:, perm]
    H = H[perm][:, perm]
    U = inverse_upper_factor(Hd[perm][:, perm])

    def evaluate(s):
        Q = divide_rows(W, s, 0)
        E = np.zeros_like(W)
        sweep_in_place(Q, E, U, grid, 32, 8)
        Q = divide_rows(Q, 1 / s, 0)
        return weighted_sq_error(H, Q - W)

    return base, evaluate


def search_scale_gptq(data, grid, axis, H, damp=0.01, rule="diag", lo=0.05, hi=1.0, points=100):
    # ref: scaling.py:137-190
    base, evaluate = gptq_scale_evaluator(data, grid, axis, H, damp, rule)
    factors = np.linspace(lo, hi, points, dtype=np.float32)
    return _grid_argmin(base, factors, evaluate)


def choose_scale(data, grid, H, mode="mse", axis=0, lo=0.05, hi=1.0, points=100):
    # ref: scaling.py:193-238
    if mode == "max":
        return no_clip_scale(data, grid, axis)
    if mode == "norm":
        return rms_scale(data, axis)
    if mode == "obq":
        return search_scale_gptq(data, grid, axis, H=H, lo=lo, hi=hi, points=points)
    if mode == "mse":
        H = None
    elif mode.startswith("hessian"):
        if len(mode) > 7:
            H = H + 0.01 * float(mode[7:]) * H.diagonal().mean() * np.eye(H.shape[0])
    elif mode.startswith("diag"):
        H = H.diagonal()
        if len(mode) > 4:
            H = H + 0.01 * float(mode[4:]) * H.mean()
    else:
        raise RuntimeError(f"Unknown scaling mode {mode}")
    return search_scale(data, grid, axis, H=H, lo=lo, hi=hi, points=points)


# --------------------------------------------------------------------------
# Per-layer preset (ref: statistics.py:146-190), numpy only
# --------------------------------------------------------------------------


def quantize_layer(W, bias, H, mean, nbits, scaling_mode="mse", order_mode="diag",
                   bias_correction=False, damp=0.01, ls_moves=0, points=100, lo=0.05, hi=1.0):
    """Returns (quantized weight, corrected bias or None)."""
    grid = UniformGrid(2**nbits, -1, 1)
    if bias_correction:
        H = strip_input_bias(H, mean)
    s = choose_scale(W, grid, H=H, mode=scaling_mode, points=points, lo=lo, hi=hi)
    Wq = quantize_scaled(W, s, grid, H=H, rule=order_mode, damp=damp, ls_moves=ls_moves)
    if bias_correction and bias is not None:
        bias = bias + ((W - Wq) * mean).sum(axis=1)
    return Wq, bias
