"""Device-tensor front end of the C ABI.

Every function here takes and returns ``torch`` CUDA tensors (contiguous, row
major), enqueues kernels on the current CUDA stream and never synchronises.
torch is used for device memory and streams only; all arithmetic of the hot
path runs in ``libsleekit_b200.so``.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import SlkCodebook

NEAREST, UP, DOWN = 0, 1, 2

# Optional per-call device timing (bench.py): set PROFILE to a dict and every wrapped entry
# point records a CUDA event pair on the stream it launches on.  None = no overhead.
PROFILE = None
PROFILE_ONLY = None  # optional set of names: time only these


def _timed(name):
    def deco(fn):
        import functools

        @functools.wraps(fn)
        def wrapper(*a, **k):
            if PROFILE is None or (PROFILE_ONLY is not None and name not in PROFILE_ONLY):
                return fn(*a, **k)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = fn(*a, **k)
            e.record()
            PROFILE.setdefault(name, []).append((s, e))
            return out

        return wrapper

    return deco


def profile_totals_ms(profile):
    """{name: (total ms, calls)} after a synchronize."""
    return {k: (sum(s.elapsed_time(e) for s, e in v), len(v)) for k, v in profile.items()}


def launch_count():
    return int(_lib.load().slk_launch_count())


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "sleekit_b200 runs its hot path on a CUDA device (B200, sm_100a) only; "
            "no CUDA device is visible and there is no CPU fallback"
        )


def device():
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class _Target:
    """Device of the tensors of the call being assembled: _chk() notes it, _stream() takes that
    device's current stream and _call() makes it the current device around the C call, so a layer
    that lives on cuda:1 launches there even while cuda:0 is the process's current device."""
    dev = None


def _stream():
    return C.c_void_p(torch.cuda.current_stream(_Target.dev).cuda_stream)


# development aid (tools/timeline.py --ops): {"buf": int64 CUDA tensor, "names": [], "tag": any}; every C call is then
# followed by a stream-ordered %globaltimer stamp so that a replayed graph yields a per-call Gantt chart
TRACE = None


def _call(name, *args):
    if TRACE is not None and name != "slk_debug_timestamp":
        out = _call_inner(name, *args)
        k = len(TRACE["names"])
        if k < TRACE["buf"].numel():
            TRACE["names"].append((name, TRACE.get("tag"), int(torch.cuda.current_stream(_Target.dev).cuda_stream)))
            _lib.call("slk_debug_timestamp", C.c_void_p(TRACE["buf"].data_ptr() + 8 * k), _stream())
        return out
    return _call_inner(name, *args)


def _call_inner(name, *args):
    dev = _Target.dev
    if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
        with torch.cuda.device(dev):
            return _lib.call(name, *args)
    return _lib.call(name, *args)


def _ws(nbytes, dev):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=dev)


def _chk(t, dtype=None):
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    _Target.dev = t.device
    if dtype is not None:
        assert t.dtype == dtype, f"expected {dtype}, got {t.dtype}"
    return t


# ---------------------------------------------------------------------------
# codebooks
# ---------------------------------------------------------------------------


class DeviceCodebook:
    """Host struct + the device tables it points to (kept alive together)."""

    def __init__(self, kind, size, lo, hi, step, values=None, limits=None):
        self.kind, self.size = int(kind), int(size)
        self.lo, self.hi, self.step = float(lo), float(hi), float(step)
        self.values, self.limits = values, limits
        self.struct = SlkCodebook(self.kind, self.size, self.lo, self.hi, self.step,
                                  values.data_ptr() if values is not None else None,
                                  limits.data_ptr() if limits is not None else None)

    @property
    def ref(self):
        return C.byref(self.struct)

    @property
    def index_dtype(self):
        return torch.uint8 if self.size <= 2**8 else (torch.uint16 if self.size <= 2**16 else torch.uint32)


def device_codebook(cb):
    """Build (and cache on the object) the device view of a codebook-like object.

    Accepted: anything exposing (codebook_size, min_val, max_val) -- UniformCodebook,
    codebook.py:4-41 -- or (values, thresholds) -- Codebook, codebook.py:98-113.
    Arbitrary Python callables cannot run on the device and are rejected."""
    if isinstance(cb, DeviceCodebook):
        return cb
    dev = device()
    cached = getattr(cb, "_slk_dev", None)
    key = None
    if hasattr(cb, "codebook_size") and hasattr(cb, "min_val"):
        key = ("u", int(cb.codebook_size), float(cb.min_val), float(cb.max_val), dev.index)
        if cached is not None and cached[0] == key:
            return cached[1]
        size = int(cb.codebook_size)
        lo, hi = float(cb.min_val), float(cb.max_val)
        d = DeviceCodebook(0, size, lo, hi, (hi - lo) / (size - 1))
    elif hasattr(cb, "values") and hasattr(cb, "thresholds"):
        vals = np.ascontiguousarray(cb.values, dtype=np.float32)
        lims = np.ascontiguousarray(cb.thresholds, dtype=np.float32)
        key = ("t", vals.tobytes(), lims.tobytes(), dev.index)
        if cached is not None and cached[0] == key:
            return cached[1]
        tv = torch.from_numpy(vals.copy()).to(dev)
        tl = torch.from_numpy(lims.copy()).to(dev) if lims.size else torch.zeros(1, dtype=torch.float32, device=dev)
        d = DeviceCodebook(1, vals.size, float(vals[0]), float(vals[-1]), 0.0, tv, tl)
    else:
        raise TypeError(
            "quantizer must be a UniformCodebook or Codebook (or expose their attributes); "
            f"{type(cb).__name__} cannot be evaluated on the device and there is no host fallback"
        )
    try:
        cb._slk_dev = (key, d)
    except Exception:
        pass
    return d


# ---------------------------------------------------------------------------
# K4 rounding and the scaling primitives
# ---------------------------------------------------------------------------


@_timed("round_to_codebook")
def round_to_codebook(x, cb, mode=NEAREST, want_val=True, want_idx=False):
    cb = device_codebook(cb)
    _chk(x)
    assert x.dtype in (torch.float32, torch.float64)
    val = idx = None
    if want_val:
        vdt = torch.float32 if (cb.kind == 1 or x.dtype == torch.float32) else torch.float64
        val = torch.empty(x.shape, dtype=vdt, device=x.device)
    if want_idx:
        idx = torch.empty(x.shape, dtype=cb.index_dtype, device=x.device)
    fn = "slk_round_f32" if x.dtype == torch.float32 else "slk_round_f64"
    _call(fn, _ptr(x), x.numel(), cb.ref, mode, _ptr(val), _ptr(idx), _stream())
    return val, idx


@_timed("scale_axis")
def scale_axis(x, s, outer, length, inner, mode=0):
    """out[o, a, i] = x[o, a, i] / s[a]  (mode 1: / (1 / s[a]))."""
    _chk(x)
    _chk(s, x.dtype)
    assert s.numel() == length and x.numel() == outer * length * inner
    out = torch.empty_like(x)
    fn = "slk_scale_axis_f32" if x.dtype == torch.float32 else "slk_scale_axis_f64"
    _call(fn, _ptr(x), outer, length, inner, _ptr(s), mode, _ptr(out), _stream())
    return out


def scale_rows(w, s, mode=0):
    return scale_axis(w, s, 1, w.shape[0], w[0].numel(), mode)


@_timed("row_noclip_scale")
def row_noclip_scale(w2d, cb_min, cb_max):
    _chk(w2d)
    out = torch.empty(w2d.shape[0], dtype=w2d.dtype, device=w2d.device)
    fn = "slk_row_noclip_scale_f32" if w2d.dtype == torch.float32 else "slk_row_noclip_scale_f64"
    _call(fn, _ptr(w2d), w2d.shape[0], w2d.shape[1], float(cb_min), float(cb_max), _ptr(out), _stream())
    return out


def row_rms_scale(w2d):
    _chk(w2d)
    out = torch.empty(w2d.shape[0], dtype=w2d.dtype, device=w2d.device)
    fn = "slk_row_rms_scale_f32" if w2d.dtype == torch.float32 else "slk_row_rms_scale_f64"
    _call(fn, _ptr(w2d), w2d.shape[0], w2d.shape[1], _ptr(out), _stream())
    return out


# ---------------------------------------------------------------------------
# K5 / full-H scale search
# ---------------------------------------------------------------------------


@_timed("scale_search")
def scale_search(w, cb, factors, hdiag=None, want_err=False, want_init=False):
    cb = device_codebook(cb)
    _chk(w, torch.float32)
    _chk(factors, torch.float32)
    r, n = w.shape
    h_dtype = 0
    if hdiag is not None:
        _chk(hdiag)
        assert hdiag.numel() == n
        h_dtype = 1 if hdiag.dtype == torch.float32 else 2
        assert hdiag.dtype in (torch.float32, torch.float64)
    out = torch.empty(r, dtype=torch.float32, device=w.device)
    err = torch.empty(r, dtype=torch.float32, device=w.device) if want_err else None
    init = torch.empty(r, dtype=torch.float32, device=w.device) if want_init else None
    _call("slk_scale_search_f32", _ptr(w), r, n, cb.ref, _ptr(factors), factors.numel(), _ptr(hdiag), h_dtype,
              _ptr(out), _ptr(err), _ptr(init), _stream())
    return out, err, init


@_timed("scale_search_fullh")
def scale_search_fullh(w, cb, factors, h, want_err=False, want_check=False):
    """compute_min_mse_scaling with a 2-D H (scaling.py:98-134): (scales, errors or None[, uncertified]).
    want_check: also an int32 [1] tensor = rows whose minimum the screened search could not certify as the
    minimum over all grid points (slk_scale_search_fullh_checked_f32; 0 expected, -1 = path without a check)."""
    cb = device_codebook(cb)
    _chk(w, torch.float32)
    _chk(factors, torch.float32)
    _chk(h)
    r, n = w.shape
    assert h.shape == (n, n)
    h_dtype = 1 if h.dtype == torch.float32 else 2
    lib = _lib.load()
    nbytes = lib.slk_scale_search_fullh_ws_bytes(r, n, factors.numel(), h_dtype)
    ws = _ws(nbytes, w.device)
    out = torch.empty(r, dtype=torch.float32, device=w.device)
    err = torch.empty(r, dtype=torch.float32, device=w.device) if want_err else None
    if want_check:
        bad = torch.zeros(1, dtype=torch.int32, device=w.device)
        _call("slk_scale_search_fullh_checked_f32", _ptr(w), r, n, cb.ref, _ptr(factors), factors.numel(), _ptr(h), h_dtype,
              _ptr(ws), nbytes, _ptr(out), _ptr(err), _ptr(bad), _stream())
        return out, err, bad
    _call("slk_scale_search_fullh_f32", _ptr(w), r, n, cb.ref, _ptr(factors), factors.numel(), _ptr(h), h_dtype,
              _ptr(ws), nbytes, _ptr(out), _ptr(err), _stream())
    return out, err


# ---------------------------------------------------------------------------
# K6 errors, gains
# ---------------------------------------------------------------------------


@_timed("hweighted_error")
def hweighted_error(w, q, h):
    """((w - q) @ h * (w - q)).sum(-1); q may be None (w is the residual)."""
    _chk(w)
    _chk(h, w.dtype)
    r, n = w.shape
    if q is not None:
        _chk(q, w.dtype)
    lib = _lib.load()
    esz = 4 if w.dtype == torch.float32 else 8
    nbytes = lib.slk_hweighted_error_ws_bytes(r, n, esz)
    ws = _ws(nbytes, w.device)
    out = torch.empty(r, dtype=w.dtype, device=w.device)
    fn = "slk_hweighted_error_f32" if esz == 4 else "slk_hweighted_error_f64"
    _call(fn, _ptr(w), _ptr(q), _ptr(h), r, n, _ptr(ws), nbytes, _ptr(out), _stream())
    return out


def mean(v):
    _chk(v)
    out = torch.empty((), dtype=v.dtype, device=v.device)
    fn = "slk_mean_f32" if v.dtype == torch.float32 else "slk_mean_f64"
    _call(fn, _ptr(v), v.numel(), _ptr(out), _stream())
    return out


def gain(w, q, h, cand):
    for t in (w, q, h, cand):
        _chk(t, w.dtype)
    r, n = w.shape
    out = torch.empty_like(w)
    fn = "slk_gain_f32" if w.dtype == torch.float32 else "slk_gain_f64"
    _call(fn, _ptr(w), _ptr(q), _ptr(h), _ptr(cand), r, n, _ptr(out), _stream())
    return out


# ---------------------------------------------------------------------------
# K1 statistics
# ---------------------------------------------------------------------------


@_timed("hessian_accum")
def hessian_accum(x, hess, mean_vec, keep, new_count):
    """In place: mean = mean*keep + colsum(x)/new_count; hess = hess*keep + x^T x/new_count."""
    _chk(x, torch.float32)
    _chk(hess, torch.float32)
    _chk(mean_vec, torch.float32)
    S, n = x.shape
    assert hess.shape == (n, n) and mean_vec.shape == (n,)
    nbytes = _lib.load().slk_hessian_accum_ws_bytes(S, n)
    ws = _ws(nbytes, x.device)
    _call("slk_hessian_accum_f32", _ptr(x), S, n, n, _ptr(hess), _ptr(mean_vec), float(keep), float(new_count),
              _ptr(ws), nbytes, _stream())


@_timed("remove_input_bias")
def remove_input_bias(h, m):
    _chk(h)
    _chk(m, h.dtype)
    out = torch.empty_like(h)
    fn = "slk_remove_input_bias_f32" if h.dtype == torch.float32 else "slk_remove_input_bias_f64"
    _call(fn, _ptr(h), _ptr(m), h.shape[0], _ptr(out), _stream())
    return out


# ---------------------------------------------------------------------------
# ordering
# ---------------------------------------------------------------------------


def damp_value(h, damp):
    _chk(h, torch.float32)
    out = torch.empty(1, dtype=torch.float32, device=h.device)
    _call("slk_damp_value_f32", _ptr(h), h.shape[0], float(damp), _ptr(out), _stream())
    return out


@_timed("col_resid_sums")
def col_resid_sums(w, cb, squared):
    cb = device_codebook(cb)
    _chk(w, torch.float32)
    out = torch.empty(w.shape[1], dtype=torch.float32, device=w.device)
    _call("slk_col_resid_sums_f32", _ptr(w), w.shape[0], w.shape[1], cb.ref, 1 if squared else 0, _ptr(out),
              _stream())
    return out


def order_keys(h, dampval=None, colsum=None):
    _chk(h, torch.float32)
    keys = torch.empty(h.shape[0], dtype=torch.float64, device=h.device)
    _call("slk_order_keys", _ptr(h), h.shape[0], _ptr(dampval), _ptr(colsum), _ptr(keys), _stream())
    return keys


@_timed("argsort")
def argsort(keys):
    _chk(keys, torch.float64)
    order = torch.empty(keys.numel(), dtype=torch.int64, device=keys.device)
    _call("slk_argsort_f64", _ptr(keys), keys.numel(), _ptr(order), _stream())
    return order


def pivot_order(h64):
    """Greedy pivoted-Cholesky ordering of an fp64 matrix (obq.py:140-166); int64 order."""
    _chk(h64, torch.float64)
    n = h64.shape[0]
    nbytes = _lib.load().slk_pivot_order_ws_bytes(n)
    ws = _ws(nbytes, h64.device)
    order = torch.empty(n, dtype=torch.int64, device=h64.device)
    _call("slk_pivot_order_f64", _ptr(h64), n, _ptr(ws), nbytes, _ptr(order), _stream())
    return order


@_timed("permute_cols")
def permute_cols(src, idx, scatter=False):
    _chk(src, torch.float32)
    _chk(idx, torch.int64)
    dst = torch.empty_like(src)
    _call("slk_permute_cols_f32", _ptr(src), src.shape[0], src.shape[1], _ptr(idx), 1 if scatter else 0,
              _ptr(dst), _stream())
    return dst


@_timed("permute_cols")
def scale_permute_cols(src, idx, s, scatter=False):
    """Row scaling fused with the column permutation: gather -> src[:, idx] / s[:, None];
    scatter -> dst[:, idx] = src / (1 / s)[:, None].  idx may be None."""
    _chk(src, torch.float32)
    _chk(s, torch.float32)
    if idx is not None:
        _chk(idx, torch.int64)
    dst = torch.empty_like(src)
    _call("slk_scale_permute_cols_f32", _ptr(src), src.shape[0], src.shape[1], _ptr(idx), _ptr(s),
              1 if scatter else 0, _ptr(dst), _stream())
    return dst


# ---------------------------------------------------------------------------
# K2 / K3 / K7
# ---------------------------------------------------------------------------


def symmetric_block_rows(n):
    """Rows per block row of the symmetric upload: 4 block rows up to n = 1024, 16 above (multiples of 32)."""
    nb = 4 if n <= 1024 else 16
    return max(32, ((n + nb - 1) // nb + 31) // 32 * 32)


def upload_symmetric(h_pinned, h_dev):
    """Enqueue the upload of a symmetric page-locked host matrix into h_dev moving only its block upper
    triangle over PCIe (slk_upload_symmetric_f32); returns the bytes sent."""
    assert h_pinned.is_pinned() and h_pinned.dtype == torch.float32 and h_pinned.is_contiguous()
    _chk(h_dev, torch.float32)
    n = h_dev.shape[0]
    assert h_pinned.shape == (n, n) and h_dev.shape == (n, n)
    bs = symmetric_block_rows(n)
    _call("slk_upload_symmetric_f32", C.c_void_p(h_pinned.data_ptr()), _ptr(h_dev), n, bs, _stream())
    return int(_lib.load().slk_upload_symmetric_bytes(n, bs))


def upload_symmetric_copy(h_pinned, h_dev):
    """DMA half of upload_symmetric (current stream)."""
    assert h_pinned.is_pinned() and h_pinned.dtype == torch.float32 and h_pinned.is_contiguous()
    _chk(h_dev, torch.float32)
    n = h_dev.shape[0]
    _call("slk_upload_symmetric_copy_f32", C.c_void_p(h_pinned.data_ptr()), _ptr(h_dev), n, symmetric_block_rows(n), _stream())


def mirror_symmetric(h_dev):
    """Device half of upload_symmetric (current stream): mirrors the block upper triangle."""
    _chk(h_dev, torch.float32)
    n = h_dev.shape[0]
    _call("slk_mirror_symmetric_f32", _ptr(h_dev), n, symmetric_block_rows(n), _stream())


@_timed("hinv")
def hinv(h, order=None, dampval=None, want64=False, want32=True):
    """Upper factor U of the inverse of (h + dampval*I)[order][:, order]; returns (u64, u32, info)."""
    _chk(h)
    n = h.shape[0]
    lib = _lib.load()
    nbytes = lib.slk_hinv_ws_bytes(n)
    ws = _ws(nbytes, h.device)
    u64 = torch.empty((n, n), dtype=torch.float64, device=h.device) if want64 else None
    u32 = torch.empty((n, n), dtype=torch.float32, device=h.device) if want32 else None
    info = torch.empty(1, dtype=torch.int32, device=h.device)
    if h.dtype == torch.float32:
        if order is not None:
            _chk(order, torch.int64)
        _call("slk_hinv_from_f32", _ptr(h), n, _ptr(order), _ptr(dampval), _ptr(ws), nbytes, _ptr(u64), _ptr(u32),
                  _ptr(info), _stream())
    else:
        assert h.dtype == torch.float64 and order is None and dampval is None
        _call("slk_hinv_from_f64", _ptr(h), n, _ptr(ws), nbytes, _ptr(u64), _ptr(u32), _ptr(info), _stream())
    return u64, u32, info


@_timed("chol_factor")
def chol_factor(h, order=None, dampval=None, want_rt=True):
    """fp64 Cholesky of (h + dampval*I)[order][:, order] in the sweep's orientation: returns
    (r32 [n,n] upper with H_opt = R R^T, rt = (hi, lo) TF32 parts of its transpose (lower part
    only) or None, ud32 [ceil(n/32),32,32] diagonal-block inverses, info)."""
    _chk(h, torch.float32)
    n = h.shape[0]
    lib = _lib.load()
    nbytes = lib.slk_chol_factor_ws_bytes(n)
    ws = _ws(nbytes, h.device)
    r32 = torch.empty((n, n), dtype=torch.float32, device=h.device)
    rt = torch.empty((2, n, n), dtype=torch.float32, device=h.device) if want_rt else None
    ud32 = torch.empty(((n + 31) // 32, 32, 32), dtype=torch.float32, device=h.device)
    info = torch.empty(1, dtype=torch.int32, device=h.device)
    if order is not None:
        _chk(order, torch.int64)
    _call("slk_chol_factor_f32", _ptr(h), n, _ptr(order), _ptr(dampval), _ptr(ws), nbytes, _ptr(r32),
              _ptr(rt[0]) if want_rt else None, _ptr(rt[1]) if want_rt else None, _ptr(ud32), _ptr(info), _stream())
    return r32, rt, ud32, info


def create_streams(count, priority=0):
    """`count` real CUDA streams of the given priority on the current device (slk_stream_create), as
    torch ExternalStreams.  torch.cuda.Stream() draws from a pool of 32 streams per priority, which
    would make 72 "independent" layers share 32 queues."""
    require_cuda()
    out = []
    for _ in range(int(count)):
        p = C.c_void_p()
        _lib.call("slk_stream_create", int(priority), C.byref(p))
        out.append(torch.cuda.ExternalStream(p.value))
    return out


def set_option(name, value):
    _lib.call("slk_set_option", name.encode(), int(value))


def timestamp(trace, row, col):
    """Development aid: %globaltimer into trace[row, col] (int64 CUDA tensor) when the current stream gets there."""
    _Target.dev = trace.device
    _call("slk_debug_timestamp", C.c_void_p(trace.data_ptr() + 8 * (row * trace.shape[1] + col)), _stream())


def _ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[(t.data_ptr() if t is not None else None) for t in tensors])


@_timed("chol_factor")
def chol_factor_batched(hs, orders=None, dampvals=None, want_rt=True):
    """chol_factor for a list of equal-sized matrices in ONE launch sequence whose tile tasks share a
    ticket queue (slk_chol_factor_batched_f32).  Returns a list of (r32, rt, ud32, info) per matrix."""
    B = len(hs)
    n = hs[0].shape[0]
    dev = hs[0].device
    for h in hs:
        _chk(h, torch.float32)
        assert h.shape == (n, n)
    lib = _lib.load()
    nbytes = int(lib.slk_chol_factor_ws_bytes(n))
    nbytes = (nbytes + 255) // 256 * 256
    ws = torch.empty((B, nbytes), dtype=torch.uint8, device=dev)
    r32 = torch.empty((B, n, n), dtype=torch.float32, device=dev)
    rt = torch.empty((B, 2, n, n), dtype=torch.float32, device=dev) if want_rt else None
    ud32 = torch.empty((B, (n + 31) // 32, 32, 32), dtype=torch.float32, device=dev)
    info = torch.empty((B, 1), dtype=torch.int32, device=dev)
    if orders is not None:
        for o in orders:
            if o is not None:
                _chk(o, torch.int64)
    _call("slk_chol_factor_batched_f32", B, _ptr_array(hs), n,
          _ptr_array(orders) if orders is not None else None,
          _ptr_array(dampvals) if dampvals is not None else None,
          _ptr_array([ws[k] for k in range(B)]), _ptr_array([r32[k] for k in range(B)]),
          _ptr_array([rt[k, 0] for k in range(B)]) if want_rt else None,
          _ptr_array([rt[k, 1] for k in range(B)]) if want_rt else None,
          _ptr_array([ud32[k] for k in range(B)]), _ptr_array([info[k] for k in range(B)]), _stream())
    return [(r32[k], rt[k] if want_rt else None, ud32[k], info[k]) for k in range(B)]


class PeerWorkspace:
    """This rank's workspace of the multi-GPU factorisation in peer-visible device memory
    (slk_peer_alloc), plus the peers' workspaces mapped into this process (slk_peer_open).  The CUDA
    IPC handles travel through torch.distributed (all_gather_object)."""

    def __init__(self, nbytes, group=None):
        import torch.distributed as dist

        require_cuda()
        self.nbytes = int(nbytes)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        _lib.call("slk_peer_alloc", self.nbytes, C.byref(ptr), handle)
        self.ptr = ptr.value
        self._opened = []
        self.peers = [None] * self.world
        self.peers[self.rank] = self.ptr
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, handle.raw, group=group)
            for q in range(self.world):
                if q == self.rank:
                    continue
                p = C.c_void_p()
                _lib.call("slk_peer_open", C.create_string_buffer(handles[q], 64), C.byref(p))
                self.peers[q] = p.value
                self._opened.append(p.value)
            dist.barrier(group=group)

    def peer_array(self):
        return (C.c_void_p * self.world)(*self.peers)

    def close(self):
        import torch.distributed as dist

        if self.ptr is None:
            return
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)
        for p in self._opened:
            _lib.call("slk_peer_close", C.c_void_p(p))
        self._opened = []
        if self.world > 1:
            dist.barrier(group=self.group)
        _lib.call("slk_peer_free", C.c_void_p(self.ptr))
        self.ptr = None

    def __del__(self):
        try:
            if self.ptr is not None and not self._opened:
                _lib.call("slk_peer_free", C.c_void_p(self.ptr))
                self.ptr = None
        except Exception:
            pass


class PeerBuffer(PeerWorkspace):
    """A peer-visible fp32 buffer (one per rank, mapped everywhere) + the in-place NVLink all-reduce over it
    (slk_peer_allreduce_f32)."""

    def __init__(self, count, group=None):
        self.count = (int(count) + 3) // 4 * 4
        super().__init__(4 * self.count, group)

    def tensor(self):
        """This rank's buffer as a float32 CUDA tensor view (no copy)."""
        iface = {"shape": (self.count,), "typestr": "<f4", "data": (self.ptr, False), "version": 3, "strides": None}

        class _Holder:
            __cuda_array_interface__ = iface

        return torch.as_tensor(_Holder(), device=device())

    def allreduce(self, barrier):
        """Sum over the ranks, in place, every copy bit-identical afterwards.  barrier(): stream-ordered."""
        barrier()
        _Target.dev = device()
        _call("slk_peer_allreduce_f32", self.peer_array(), self.world, self.rank, self.count, _stream())
        barrier()


def chol_dist_ws_bytes(n):
    return int(_lib.load().slk_chol_factor_ws_bytes(n))


@_timed("chol_factor")
def chol_factor_dist(h, order, dampval, pws, barrier, want_rt=True):
    """chol_factor with the tile rows of the factorisation spread over the ranks of pws (PeerWorkspace):
    finished tiles are pushed to every peer through NVLink (slk_chol_dist_*), so every rank ends with
    the complete factor.  barrier(): a stream-ordered barrier over the ranks (an all-reduce of one
    element on the current stream).  Returns (r32, rt, ud32, info) like chol_factor; info is
    max-reduced over the ranks by the caller if it is going to be read."""
    _chk(h, torch.float32)
    n = h.shape[0]
    assert pws.nbytes >= chol_dist_ws_bytes(n)
    r32 = torch.empty((n, n), dtype=torch.float32, device=h.device)
    rt = torch.empty((2, n, n), dtype=torch.float32, device=h.device) if want_rt else None
    ud32 = torch.empty(((n + 31) // 32, 32, 32), dtype=torch.float32, device=h.device)
    info = torch.empty(1, dtype=torch.int32, device=h.device)
    if order is not None:
        _chk(order, torch.int64)
    ws = C.c_void_p(pws.ptr)
    _call("slk_chol_dist_gather_f32", _ptr(h), n, _ptr(order), _ptr(dampval), ws, pws.nbytes, _ptr(info), _stream())
    barrier()
    _call("slk_chol_dist_factor", n, ws, pws.world, pws.rank, pws.peer_array(), _ptr(info), _stream())
    barrier()
    _call("slk_chol_dist_export_f32", n, ws, _ptr(r32), _ptr(rt[0]) if want_rt else None,
          _ptr(rt[1]) if want_rt else None, _ptr(ud32), _stream())
    return r32, rt, ud32, info


def sym_pack(h, packed, scale=1.0):
    """packed[:L] = scale * (block upper triangle of the symmetric h); returns L (floats)."""
    _chk(h, torch.float32)
    _chk(packed, torch.float32)
    n = h.shape[0]
    bs = symmetric_block_rows(n)
    L = int(_lib.load().slk_upload_symmetric_bytes(n, bs)) // 4
    assert packed.numel() >= L
    _call("slk_sym_pack_f32", _ptr(h), n, bs, float(scale), _ptr(packed), _stream())
    return L


def sym_packed_len(n):
    return int(_lib.load().slk_upload_symmetric_bytes(n, symmetric_block_rows(n))) // 4


def sym_unpack(packed, h, scale=1.0):
    """h = scale * unpack(packed), mirrored to the full symmetric matrix."""
    _chk(h, torch.float32)
    _chk(packed, torch.float32)
    n = h.shape[0]
    _call("slk_sym_unpack_f32", _ptr(packed), n, symmetric_block_rows(n), float(scale), _ptr(h), _stream())
    return h


@_timed("gptq_sweep")
def gptq_sweep_r(q, r32, rt, ud32, cb, d=None, err_sums=None):
    """In place on q; the sweep from the Cholesky factor (chol_factor); returns (q, d = W - Q).
    err_sums: optional [rows, 2] fp32 out -- per row (sum E^2, sum (W-Q)^2), see sweep_error."""
    cb = device_codebook(cb)
    _chk(q, torch.float32)
    _chk(r32, torch.float32)
    _chk(ud32, torch.float32)
    if d is None:
        d = torch.empty_like(q)
    if err_sums is not None:
        _chk(err_sums, torch.float32)
        assert err_sums.shape == (q.shape[0], 2)
    nbytes = _lib.load().slk_gptq_sweep_r_ws_bytes(q.shape[0], q.shape[1]) if rt is not None else 0
    ws = _ws(nbytes, q.device) if nbytes else None
    _call("slk_gptq_sweep_r_err_f32", _ptr(q), _ptr(d), q.shape[0], q.shape[1], _ptr(r32),
              _ptr(rt[0]) if rt is not None else None, _ptr(rt[1]) if rt is not None else None, _ptr(ud32),
              cb.ref, _ptr(ws), nbytes, _ptr(err_sums), _stream())
    return q, d


@_timed("sweep_error")
def sweep_error(err_sums, row_scale=None, dampval=None, want_rows=False):
    """Layer error from the sweep's row sums (slk_sweep_error_f32): mean over rows of
    scale^2 * (sum E^2 - damp * sum (W-Q)^2) = quantization_error(W, Q_descaled, H) (obq.py:89-103)
    for the H the sweep's factor was formed from.  Returns (mean [1], rows or None)."""
    _chk(err_sums, torch.float32)
    r = err_sums.shape[0]
    out = torch.empty(1, dtype=torch.float32, device=err_sums.device)
    rows = torch.empty(r, dtype=torch.float32, device=err_sums.device) if want_rows else None
    _call("slk_sweep_error_f32", _ptr(err_sums), _ptr(row_scale), _ptr(dampval), r, _ptr(rows), _ptr(out), _stream())
    return out, rows


@_timed("gptq_sweep")
def gptq_sweep(q, u64, u32, cb, leaf=32, fanout=8, e=None, exact_leaf=False):
    """In place on q ([rows, n] scaled, permuted weights -> quantized values); returns (q, e).
    exact_leaf=True reproduces the reference's fp64/fp32 leaf arithmetic bit for bit from u64."""
    cb = device_codebook(cb)
    _chk(q, torch.float32)
    _chk(u32, torch.float32)
    if exact_leaf:
        _chk(u64, torch.float64)
    if e is None:
        e = torch.empty_like(q)
    _call("slk_gptq_sweep_f32", _ptr(q), _ptr(e), q.shape[0], q.shape[1], _ptr(u64) if exact_leaf else None,
              _ptr(u32), cb.ref, int(leaf), int(fanout), 1 if exact_leaf else 0, _stream())
    return q, e


def row_wsq(e, h=None):
    """sum_j h[j] e[row, j]^2 per row (h None: sum of squares): _compute_mse's None / 1-D branches."""
    _chk(e)
    assert e.dtype in (torch.float32, torch.float64) and e.ndim == 2
    odt = e.dtype
    fn = "slk_row_wsq_f32" if e.dtype == torch.float32 else "slk_row_wsq_f64"
    if h is not None:
        _chk(h)
        assert h.numel() == e.shape[1] and h.dtype in (torch.float32, torch.float64)
        if h.dtype != e.dtype:
            if e.dtype == torch.float64:
                h = h.to(torch.float64)
            else:                                    # fp32 residuals, fp64 diagonal: result fp64
                fn, odt = "slk_row_wsq_f32_h64", torch.float64
    out = torch.empty(e.shape[0], dtype=odt, device=e.device)
    _call(fn, _ptr(e), _ptr(h), e.shape[0], e.shape[1], _ptr(out), _stream())
    return out


class LocalSearchState:
    """Workspace of an instalment-wise local search (slk_local_search_step_f32): keeps P = (Q - W) H
    between calls so that LocalSearchQuantizer.do_move does not redo the 2 r n^2 product per move."""

    def __init__(self, w):
        r, n = w.shape
        self.nbytes = int(_lib.load().slk_local_search_ws_bytes(r, n))
        self.ws = _ws(self.nbytes, w.device)
        self.valid = False


@_timed("local_search")
def local_search_step(w, q, h, cb, moves, state):
    """In place on q; continues from the state a previous call left (same w, q, h)."""
    cb = device_codebook(cb)
    _chk(w, torch.float32)
    _chk(q, torch.float32)
    _chk(h, torch.float32)
    if moves <= 0:
        return q
    r, n = w.shape
    _call("slk_local_search_step_f32", _ptr(w), _ptr(q), _ptr(h), r, n, cb.ref, int(moves), _ptr(state.ws), state.nbytes,
          1 if state.valid else 0, _stream())
    state.valid = True
    return q


@_timed("local_search")
def local_search(w, q, h, cb, moves, err_sums=None):
    """In place on q (values on the codebook).  err_sums: optional [rows, 2] fp32 tensor that receives
    (channelwise_error of the row after its moves, 0), the layout sweep_error() takes (moves >= 1)."""
    cb = device_codebook(cb)
    _chk(w, torch.float32)
    _chk(q, torch.float32)
    _chk(h, torch.float32)
    if moves <= 0:
        assert err_sums is None, "the row errors come out of the moves"
        return q
    r, n = w.shape
    lib = _lib.load()
    nbytes = lib.slk_local_search_ws_bytes(r, n)
    ws = _ws(nbytes, w.device)
    if err_sums is not None:
        _chk(err_sums, torch.float32)
        assert err_sums.shape == (r, 2)
        _call("slk_local_search_err_f32", _ptr(w), _ptr(q), _ptr(h), r, n, cb.ref, int(moves), _ptr(ws), nbytes,
              _ptr(err_sums), _stream())
    else:
        _call("slk_local_search_f32", _ptr(w), _ptr(q), _ptr(h), r, n, cb.ref, int(moves), _ptr(ws), nbytes, _stream())
    return q


@_timed("bias_delta")
def bias_delta(w, wq, mean_vec):
    _chk(w, torch.float32)
    _chk(wq, torch.float32)
    _chk(mean_vec, torch.float32)
    out = torch.empty(w.shape[0], dtype=torch.float32, device=w.device)
    _call("slk_bias_delta_f32", _ptr(w), _ptr(wq), _ptr(mean_vec), w.shape[0], w.shape[1], _ptr(out), _stream())
    return out
