"""CPU statements of the two identities the CUDA path relies on beyond the reference's own op
sequence, checked with the oracle's arithmetic (numpy, no GPU):

* layer error from the sweep's residuals (sweep.cu, slk_gptq_sweep_r_err_f32):
      (W-Q) H_opt (W-Q)^T = sum_i E_i^2,   hence   error under H = sum E^2 - damp * sum (W-Q)^2
* the scale search's threshold tables, their mirror image for negative weights and the walk along
  an ascending grid (scale_search.cu): the index every weight gets equals the reference's chain.
"""

import numpy as np

from oracle import sleekit_oracle as orc
from sleekit_b200 import workloads as wl


def test_sweep_residuals_give_the_layer_error():
    r, n = 24, 160
    W, H, _ = wl.synthetic_layer(r, n, 21, samples=96)          # samples < n: rank deficient, damping matters
    grid = orc.UniformGrid(8, -1, 1)
    sc = orc.search_scale(W, grid, 0, H=H.diagonal())
    Ws = orc.divide_rows(W, sc, 0)
    damp = 0.01 * H.diagonal().mean()                              # obq.py:198
    Hd = H + damp * np.eye(n)
    perm = orc.column_order(Ws, Hd, grid, "diag")
    Wp = Ws[:, perm].astype(np.float32)
    Q = Wp.copy()
    E = np.zeros_like(Wp)
    U = orc.inverse_upper_factor(Hd[perm][:, perm])
    orc.sweep_in_place(Q, E, U, grid)
    D = (Wp - Q).astype(np.float64)
    from_sweep = (E.astype(np.float64) ** 2).sum(1) - damp * (D ** 2).sum(1)
    direct = orc.rowwise_error(Wp.astype(np.float64), Q.astype(np.float64), H[perm][:, perm].astype(np.float64))
    assert np.all(direct > 0)
    np.testing.assert_allclose(from_sweep, direct, rtol=2e-5)
    # the damping term is not negligible here: the identity really is exercised
    assert (damp * (D ** 2).sum(1) / direct).max() > 0.05


def _ord(x):
    i = np.asarray(x, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.where(i >= 0, i, -(i & 0x7FFFFFFF))      # order-preserving integer key (-0 and +0 coincide)


def _unord(o):
    o = np.asarray(o, dtype=np.int64)
    bits = np.where(o >= 0, o, (-o) | 0x80000000).astype(np.uint32)
    return bits.view(np.float32)


def _thresholds(grid, scale):
    """T[k] = min{w : index(w / scale) >= k}, k = 1..C-1, by bisection over the ordered fp32 values with
    the oracle's own chain (scaling.py:73 then codebook.py:43-54)."""
    C = len(grid)
    scale = np.float32(scale)
    out = []
    for k in range(1, C):
        lo, hi = int(_ord(np.float32(-3.0e38))), int(_ord(np.float32(3.0e38)))
        while hi - lo > 1:
            mid = (lo + hi) // 2
            w = _unord(np.array([mid]))
            if int(grid.index(w / scale)[0]) >= k:
                hi = mid
            else:
                lo = mid
        out.append(_unord(np.array([hi]))[0])
    return np.array(out, dtype=np.float32)


def test_threshold_tables_mirror_and_walk_reproduce_the_reference_index():
    rng = np.random.default_rng(5)
    for C in (3, 4, 8):
        grid = orc.UniformGrid(C, -1, 1)
        w = (rng.standard_normal(4000) * 0.04).astype(np.float32)
        w[:8] = [0.0, -0.0, 1e-30, -1e-30, 0.5, -0.5, 3.0, -3.0]
        init = np.float32(max(w.max() / grid.max(), w.min() / grid.min()))
        factors = np.linspace(0.05, 1.0, 40, dtype=np.float32)
        tables = [_thresholds(grid, f * init) for f in factors]
        ninf, pinf = np.float32(-np.inf), np.float32(np.inf)
        T = np.array([np.concatenate(([ninf], t, [pinf])) for t in tables], dtype=np.float32)      # [G, C+1]
        # mirrored tables: T'[j] = nextup(-T[C-j])
        Tm = np.empty_like(T)
        Tm[:, 0], Tm[:, C] = ninf, pinf
        for j in range(1, C):
            Tm[:, j] = np.nextafter(-T[:, C - j], pinf)
        # walk conditions of the kernel on both table sets
        for tab in (T, Tm):
            tg, tn = tab[:-1, 1:C], tab[1:, 1:C]
            assert np.all((tg <= 0) | (tn >= tg))
            assert np.all(tab[1:, 0:C - 1] <= np.maximum(tg, 0))
        neg = w < 0
        wm = np.where(neg, -w, w).astype(np.float32)
        k = None
        for g, f in enumerate(factors):
            ref = grid.index(w / (f * init)).astype(np.int64)                                  # the reference's chain
            by_table = (w[:, None] >= T[g][None, 1:C]).sum(1)
            np.testing.assert_array_equal(by_table, ref)
            by_mirror = (wm[:, None] >= Tm[g][None, 1:C]).sum(1)
            np.testing.assert_array_equal(np.where(neg, C - 1 - by_mirror, by_table), ref)
            # the walk: index in the weight's own class, one conditional step down per grid point
            tab_w = np.where(neg[:, None], Tm[g][None, :], T[g][None, :])
            if k is None:
                k = np.where(neg, C - 1 - ref, ref)
            else:
                k = k - (wm < tab_w[np.arange(w.size), k])
            np.testing.assert_array_equal(np.where(neg, C - 1 - k, k), ref)
