"""Drop-in for ``sleekit.codebook`` (reference: sleekit/codebook.py).

Rounding (``quantize_index / quantize_value / quantize_up / quantize_down``,
codebook.py:43-95 and 155-188) runs in the K4 CUDA kernel and is bit-exact with
the reference for fp32 and fp64 inputs.  The codebook *design* helpers
(Lloyd-Max etc., codebook.py:190-367) are offline tooling outside the hot
path: they keep the reference's semantics, do their bookkeeping on the host and
call the device rounding for every assignment of data to bins.
"""

import numpy as np

from . import _convert as cv
from . import ops

__all__ = ["UniformCodebook", "Codebook", "lloyd_max", "np"]


def _round(cb, data, mode, index=False):
    """Shared host wrapper around ops.round_to_codebook."""
    if cv.is_tensor(data):
        x = cv.to_dev(data)
        if x.dtype not in (cv.torch.float32, cv.torch.float64):
            x = x.float()
        val, idx = ops.round_to_codebook(x, cb, mode, want_val=not index, want_idx=index)
        return idx if index else val
    a = np.asarray(data)
    x = cv.to_dev(a)
    val, idx = ops.round_to_codebook(x, cb, mode, want_val=not index, want_idx=index)
    out = cv.to_host(idx if index else val)
    return out.reshape(a.shape)


class UniformCodebook:
    """Evenly spaced codewords between ``min_val`` and ``max_val`` (codebook.py:4-95)."""

    def __init__(self, codebook_size, min_val, max_val):
        self.codebook_size = int(codebook_size)
        self.min_val = min_val
        self.max_val = max_val
        assert self.min_val < self.max_val
        assert self.codebook_size >= 2

    def __len__(self):
        return self.codebook_size

    @property
    def values(self):
        return np.linspace(self.min_val, self.max_val, self.codebook_size)

    def min(self):
        return self.min_val

    def max(self):
        return self.max_val

    @property
    def scale(self):
        return (self.max_val - self.min_val) / (self.codebook_size - 1)

    @property
    def zero(self):
        return self.min_val

    def quantize_index(self, data):
        return _round(self, data, ops.NEAREST, index=True)

    def quantize_value(self, data):
        return _round(self, data, ops.NEAREST)

    def quantize_up(self, data):
        return _round(self, data, ops.UP)

    def quantize_down(self, data):
        return _round(self, data, ops.DOWN)

    def __call__(self, data):
        return self.quantize_value(data)


class Codebook:
    """Sorted fp32 codewords plus the limits between their bins (codebook.py:98-188)."""

    def __init__(self, values, limits=None):
        self.values = np.array(values, dtype=np.float32)
        if limits is None:
            self.values.sort()
            self.thresholds = (self.values[:-1] + self.values[1:]) / 2
        else:
            self.thresholds = np.array(limits, dtype=np.float32)
        self.check()

    def clone(self):
        return Codebook(self.values.copy(), self.thresholds.copy())

    def check(self):
        v, t = self.values, self.thresholds
        assert v.ndim == 1 and v.size > 0 and np.isfinite(v).all()
        assert (np.diff(v) > 0).all()
        assert t.ndim == 1 and t.size == v.size - 1 and np.isfinite(t).all()
        assert (np.diff(t) > 0).all()
        assert (t >= v[:-1]).all() and (t <= v[1:]).all()

    def __len__(self):
        return len(self.values)

    def min(self):
        return self.values[0]

    def max(self):
        return self.values[-1]

    def quantize_index(self, data):
        return _round(self, data, ops.NEAREST, index=True)

    def quantize_value(self, data):
        return _round(self, data, ops.NEAREST)

    def quantize_up(self, data):
        return _round(self, data, ops.UP)

    def quantize_down(self, data):
        return _round(self, data, ops.DOWN)

    def __call__(self, data):
        return self.quantize_value(data)

    # ---- offline codebook design (host bookkeeping around the device rounding) ----

    def _counts(self, data):
        return np.bincount(self.quantize_index(data), minlength=len(self.values))

    def probabilities(self, data):
        return self._counts(data) / len(data)

    def entropy(self, data):
        p = self.probabilities(data)
        p = p[p > 0]
        return -(p * np.log2(p)).sum()

    def mse(self, data):
        return np.square(data - self.quantize_value(data)).mean()

    def centroids(self, data):
        """Mean of the data in each bin; empty bins fall back to a point derived from the limits
        (codebook.py:212-231)."""
        labels = self.quantize_index(data)
        t = self.thresholds
        last = len(self.values) - 1
        out = []
        for k in range(last + 1):
            members = data[labels == k]
            if len(members):
                out.append(members.mean())
            elif k == 0:
                out.append(t[0] - 1.0e-6)
            elif k == last:
                out.append(t[-1] + 1.0e-6)
            else:
                out.append((t[k - 1] + t[k]) / 2)
        return np.array(out)

    def remove_unused(self, data):
        used = self._counts(data) != 0
        if used.all():
            return
        self.values = self.values[used]
        keep = used[:-1].copy()  # a limit survives if the bin on its left does
        self.thresholds = self.thresholds[keep]
        if not used[-1]:
            self.thresholds = self.thresholds[:-1]
        self.check()

    def improve(self, data, lagrange_mult=0.0):
        """One Lloyd-Max round, optionally entropy-penalised (codebook.py:248-267)."""
        if lagrange_mult != 0.0:
            self.remove_unused(data)
            v = self.values
            bits = -np.log2(self.probabilities(data))
            slope = np.diff(bits) / np.diff(v)
            self.thresholds = (v[:-1] + v[1:]) / 2 + lagrange_mult * slope / 2
            self.thresholds.sort()
        else:
            v = self.values
            self.thresholds = (v[:-1] + v[1:]) / 2
        self.values = self.centroids(data)
        self.check()

    def close_to(self, other, tol=1.0e-6):
        if len(self) != len(other):
            return False
        span = max(self.values.max() - self.values.min(), 1.0e-10)
        return np.allclose(self.values, other.values, atol=tol * span)

    @staticmethod
    def random(data, codebook_size):
        pool = np.unique(data)
        return Codebook(np.random.choice(pool, min(codebook_size, pool.size), replace=False))

    @staticmethod
    def uniform(codebook_size, min_val, max_val):
        assert min_val <= max_val
        return Codebook(np.linspace(min_val, max_val, codebook_size))

    @staticmethod
    def nf4():
        """NormalFloat4 table (codebook.py:296-320)."""
        return Codebook([
            -1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453,
            -0.28444138169288635, -0.18477343022823334, -0.09105003625154495, 0.0,
            0.07958029955625534, 0.16093020141124725, 0.24611230194568634, 0.33791524171829224,
            0.44070982933044434, 0.5626170039176941, 0.7229568362236023, 1.0,
        ])

    @staticmethod
    def equiprobable(data, codebook_size):
        chunks = [c for c in np.array_split(np.sort(data), codebook_size) if len(c) > 0]
        limits = [(chunks[k][-1] + chunks[k + 1][0]) / 2 for k in range(len(chunks) - 1)]
        cb = Codebook([c.mean() for c in chunks], limits)
        cb.values = cb.centroids(data)
        return cb


def lloyd_max(data, codebook_size, lagrange_mult=0.0, max_iter=100, tol=1e-6, random_init=False,
              sample_count=None):
    """Scalar Lloyd-Max quantizer design (codebook.py:338-367)."""
    data = data.reshape((-1,))
    if sample_count is not None:
        wanted = codebook_size * sample_count
        if wanted < len(data):
            data = np.random.choice(data, wanted, replace=False)
    data = np.sort(data)
    cur = Codebook.random(data, codebook_size) if random_init else Codebook.equiprobable(data, codebook_size)
    for _ in range(max_iter):
        nxt = cur.clone()
        nxt.improve(data, lagrange_mult)
        if nxt.close_to(cur, tol):
            break
        cur = nxt
    return cur
