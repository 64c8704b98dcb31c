"""Pin oracle/ against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py) and against the known-answer vectors in the
reference's own tests.  CPU only."""

import numpy as np
import pytest

from oracle import sleekit_oracle as orc
from sleekit_b200 import workloads as wl
from tests.conftest import load_golden

UNI = [(2, 1), (3, 1), (4, 1), (8, 1), (16, 1), (9, 2), (9, 3)]


def _same_bits(a, b):
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    assert a.shape == b.shape
    # compare numerically and in sign-of-zero-insensitive bit pattern
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("c,hi", UNI)
def test_uniform_rounding_matches_reference(c, hi):
    g = load_golden("rounding")
    tag = f"u{c}_{hi}"
    cfg = g[tag + "_cfg"]
    lo_v, hi_v = (int(cfg[1]), int(cfg[2])) if hi != 3 else (float(cfg[1]), float(cfg[2]))
    grid = orc.UniformGrid(int(cfg[0]), lo_v, hi_v)
    x = g[tag + "_x"]
    _same_bits(grid.index(x), g[tag + "_idx"])
    _same_bits(grid.value(x), g[tag + "_val"])
    _same_bits(grid.up(x), g[tag + "_up"])
    _same_bits(grid.down(x), g[tag + "_down"])
    xd = x.astype(np.float64)
    _same_bits(grid.index(xd), g[tag + "_idx64"])
    _same_bits(grid.value(xd), g[tag + "_val64"])
    _same_bits(grid.up(xd), g[tag + "_up64"])
    _same_bits(grid.down(xd), g[tag + "_down64"])


def test_table_rounding_matches_reference():
    g = load_golden("rounding")
    grid = orc.TableGrid(g["nf4_values"])
    np.testing.assert_array_equal(grid.limits, g["nf4_limits"])
    x = g["nf4_x"]
    _same_bits(grid.index(x), g["nf4_idx"])
    _same_bits(grid.value(x), g["nf4_val"])
    _same_bits(grid.up(x), g["nf4_up"])
    _same_bits(grid.down(x), g["nf4_down"])


def test_table_known_answers():
    # reference tests/test_codebook.py:6-32
    grid = orc.TableGrid([-1.0, 2.0, 4.0, 8.0])
    x = [-2.0, -1.0, 0.0, 0.9, 1.9, 2.9, 3.1, 5.9, 6.1, 9.0]
    np.testing.assert_array_equal(grid.index(x), [0, 0, 0, 1, 1, 1, 2, 2, 3, 3])
    np.testing.assert_array_equal(grid.value(x), [-1, -1, -1, 2, 2, 2, 4, 4, 8, 8])
    np.testing.assert_array_equal(grid.up(x), [2, 2, 2, 4, 4, 4, 8, 8, 8, 8])
    np.testing.assert_array_equal(grid.down(x), [-1, -1, -1, -1, -1, -1, 2, 2, 4, 4])


def test_scale_known_answers():
    # reference tests/test_scaling.py:16-41, 56-72
    data = np.array([[0.0, 10.0], [5.0, 5.0]], dtype=np.float32)
    np.testing.assert_allclose(orc.rms_scale(data, 0), [10.0 / np.sqrt(2), 5.0], rtol=1e-6)
    np.testing.assert_allclose(orc.rms_scale(data, 1), [5.0 / np.sqrt(2), np.sqrt(125 / 2)], rtol=1e-6)
    data = np.array(
        [[0.0, 10.0, -20.0, 15.0], [5.0, 5.0, 10.0, -10.0], [1.0, 2.0, -4.0, 3.0], [0.0, 0.0, 0.0, 0.0],
         [1.0, 10.0, 100.0, 1000.0], [-1.0, 10.0, 100.0, 1000.0]], dtype=np.float32)
    grid = orc.TableGrid([-1.0, 0.0, 10.0, 20.0])
    np.testing.assert_allclose(orc.no_clip_scale(data, grid, 0), [20, 10, 4, 1e-16, 50, 50])
    np.testing.assert_allclose(orc.no_clip_scale(data, grid, 1), [1, 0.5, 20, 50])


def test_scales_match_reference():
    g = load_golden("scales")
    W, H = g["W"], g["H"]
    for c in (3, 8):
        grid = orc.UniformGrid(c, -1, 1)
        _same_bits(orc.no_clip_scale(W, grid, 0), g[f"max_c{c}"])
        _same_bits(orc.search_scale(W, grid, 0), g[f"mse_c{c}"])
        _same_bits(orc.search_scale(W, grid, 0, H=H.diagonal()), g[f"diag_c{c}"])
        # the full-H error goes through sgemm: allow the BLAS kernel of this host to differ
        np.testing.assert_allclose(orc.search_scale(W, grid, 0, H=H), g[f"full_c{c}"], rtol=0.03)
        _same_bits(orc.choose_scale(W, grid, H, mode="diag5"), g[f"diag5_c{c}"])
        np.testing.assert_allclose(orc.choose_scale(W, grid, H, mode="hessian2"), g[f"hess2_c{c}"], rtol=0.03)
        _same_bits(orc.search_scale(W, grid, 1, points=17, lo=0.2), g[f"axis1_c{c}"])
    _same_bits(orc.rms_scale(W, 0), g["norm0"])
    _same_bits(orc.rms_scale(W, 1), g["norm1"])
    _same_bits(orc.no_clip_scale(W, orc.TableGrid([-1.0, 0.0, 10.0, 20.0]), 0), g["max_tab0"])
    grid = orc.UniformGrid(8, -1, 1)
    _same_bits(orc.quantize_scaled(W, g["diag_c8"], grid), g["qws_plain"])
    _same_bits(orc.divide_rows(W, g["diag_c8"], 0), g["apply"])


def _agree(a, b):
    return float((a == b).mean())


def test_factor_order_sweep_match_reference():
    g = load_golden("sweep")
    W, H, m, Hd, Ws = g["W"], g["H"], g["mean"], g["Hd"], g["Ws"]
    np.testing.assert_allclose(orc.inverse_upper_factor(Hd), g["U"], rtol=1e-9, atol=1e-12)
    _same_bits(orc.strip_input_bias(H, m), g["Hc"])
    grid = orc.UniformGrid(8, -1, 1)
    for rule in ("diag", "none", "err", "sqerr"):
        np.testing.assert_array_equal(orc.column_order(Ws, Hd, grid, rule), g[f"order_{rule}"])
        assert _agree(orc.gptq(Ws, H, grid, rule=rule, damp=0.01), g[f"gptq_{rule}"]) >= 0.999
    assert _agree(orc.gptq(Ws, H, grid, rule="sqerr", damp=0.03), g["gptq_damp3"]) >= 0.999
    assert _agree(orc.gptq(Ws, H, grid, rule="diag", damp=0.01, ls_moves=20), g["gptq_ls20"]) >= 0.999
    Q = Ws.copy()
    E = np.zeros_like(Ws)
    orc.sweep_in_place(Q, E, g["U"], grid)
    assert _agree(Q, g["sweep_Q"]) >= 0.999
    np.testing.assert_allclose(E, g["sweep_E"], rtol=1e-4, atol=1e-6)
    out = orc.quantize_scaled(W, g["scale"], grid, H=H, rule="diag", damp=0.01)
    assert _agree(out, g["qws_gptq"]) >= 0.999
    np.testing.assert_allclose(orc.rowwise_error(W, g["qws_gptq"], H), g["err_rows"], rtol=1e-4)
    np.testing.assert_allclose(orc.mean_error(W, g["qws_gptq"], H), g["err_mean"], rtol=1e-4)
    grid4 = orc.UniformGrid(4, -1, 1)
    assert _agree(orc.quantize_scaled(g["W2"], g["scale2"], grid4, H=g["H2"]), g["qws2"]) >= 0.999
    assert _agree(orc.quantize_scaled(g["W2"], g["scale2"], grid4, H=g["H2"], ls_moves=15), g["qws2_ls"]) >= 0.999


def test_sweep_plan_shapes():
    # n=768 -> 24 leaves + 23 non-empty GEMMs
    plan = orc.sweep_plan(768)
    leaves = [p for p in plan if p[0] == "leaf"]
    gemms = [p for p in plan if p[0] == "gemm" and p[2] < p[3]]
    assert len(leaves) == 24 and all(b - a == 32 for _, a, b in leaves)
    assert len(gemms) == 8 * 2 + 7
    cover = sorted((a, b) for _, a, b in leaves)
    assert cover[0][0] == 0 and cover[-1][1] == 768
    assert all(cover[i][1] == cover[i + 1][0] for i in range(len(cover) - 1))


def test_local_search_matches_reference():
    g = load_golden("local_search")
    Ws, H, Q0 = g["Ws"], g["H"], g["Q0"]
    grid = orc.UniformGrid(4, -1, 1)
    np.testing.assert_allclose(orc.flip_gain(Ws, Q0, H, grid.up(Q0)), g["gain_up"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(orc.flip_gain(Ws, Q0, H, grid.down(Q0)), g["gain_down"], rtol=1e-4, atol=1e-7)
    for k in (1, 5, 30):
        assert _agree(orc.local_search(Ws, Q0, H, grid, k), g[f"ls_{k}"]) >= 0.999


def test_obq_scaling_matches_reference():
    g = load_golden("obq_scaling")
    grid = orc.UniformGrid(8, -1, 1)
    a = orc.search_scale_gptq(g["W"], grid, 0, H=g["H"], points=12, lo=0.3)
    assert _agree(a, g["sc_obq"]) >= 0.85
    b = orc.search_scale_gptq(g["W"], grid, 0, H=g["H"], points=12, lo=0.3, rule="sqerr", damp=0.03)
    assert _agree(b, g["sc_obq_sqerr"]) >= 0.85


def test_statistics_match_reference():
    g = load_golden("statistics")
    st = orc.RunningStats(48)
    st.add_rows(g["X1"])
    assert st.count == int(g["count1"])
    np.testing.assert_allclose(st.mean, g["mean1"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(st.hessian, g["hess1"], rtol=1e-5, atol=1e-5)
    st.add_rows(g["X2"])
    assert st.count == int(g["count2"])
    np.testing.assert_allclose(st.mean, g["mean2"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(st.hessian, g["hess2"], rtol=1e-5, atol=1e-5)
    # presets, starting from the reference's own statistics so the stage is isolated
    for name, kw in (
        ("basic", dict(scaling_mode="mse", order_mode="diag", bias_correction=False, damp=0.01, ls_moves=0)),
        ("light", dict(scaling_mode="diag", order_mode="sqerr", bias_correction=True, damp=0.03, ls_moves=0)),
        ("heavy", dict(scaling_mode="hessian", order_mode="sqerr", bias_correction=True, damp=0.03, ls_moves=100)),
    ):
        Wq, b = orc.quantize_layer(g["W"], g["b"], g["hess2"], g["mean2"], 3, **kw)
        assert _agree(Wq, g[f"{name}_W"]) >= 0.995, name
        if kw["bias_correction"]:
            np.testing.assert_allclose(b, g[f"{name}_b"], rtol=1e-3, atol=1e-4)


def test_synthetic_layer_is_deterministic():
    a = wl.synthetic_layer(8, 64, 3, samples=128)
    b = wl.synthetic_layer(8, 64, 3, samples=128)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)
    assert len(wl.layer_shapes("opt-125m")) == 72
    assert sum(r * n for r, n in wl.layer_shapes("opt-125m")) == 84934656


def test_pivot_ordering_vs_golden():
    """act_order = "pivot" (obq.py:140-166): the oracle's greedy pivoted-Cholesky order and GPTQ with it."""
    g = load_golden("pivot")
    grid = orc.UniformGrid(8, -1, 1)
    np.testing.assert_array_equal(orc.column_order(g["Ws"], g["Hd"], grid, "pivot"), g["order_pivot"])
    np.testing.assert_array_equal(orc.gptq(g["Ws"], g["H"], grid, rule="pivot", damp=0.01), g["gptq_pivot"])


# ---------------------------------------------------------------------------
# The reference's own property tests (tests/test_obq.py), restated on the oracle: they pin the
# factor, the blocked sweep, the bias removal and the gain formula independently of the golden
# vectors.
# ---------------------------------------------------------------------------


def _wishart(size, rank, damp=0.0, seed=0):
    # ref: obq.py:4-11 (random_psd_matrix), seeded here
    X = np.random.default_rng(seed).standard_normal((size, rank))
    return X @ X.T / rank + damp * np.eye(size)


def test_factor_is_the_cholesky_of_the_inverse():
    # reference tests/test_obq.py:21-32
    H = _wishart(4, 6, 1.0e-6, seed=1)
    gptq_way = np.linalg.cholesky(np.linalg.inv(H)).T
    U = orc.inverse_upper_factor(H)
    assert np.allclose(U, gptq_way)
    assert np.allclose(np.linalg.inv(U.T @ U), H)
    assert np.allclose(U, np.triu(U)) and U.flags["C_CONTIGUOUS"]


def test_blocked_sweep_equals_unblocked_sweep():
    # reference tests/test_obq.py:57-70
    size = 64
    H = _wishart(size, 2, 1.0e-6, seed=2)
    W = 10.0 * np.random.default_rng(3).standard_normal((1, size))
    round_to_int = lambda x: np.round(x)  # noqa: E731  (the reference's test quantizer)
    Q = orc.gptq(W, H, round_to_int, rule="none", leaf=1)
    for leaf in (3, 4, 7, 8, 63, 64):
        for fanout in (2, 4):
            assert np.allclose(Q, orc.gptq(W, H, round_to_int, rule="none", leaf=leaf, fanout=fanout))


def test_orderings_improve_the_error():
    # reference tests/test_obq.py:35-54 (smaller, seeded)
    size = 200
    H = _wishart(size, 2, 1.0e-6, seed=4)
    W = 10.0 * np.random.default_rng(5).standard_normal((10, size))
    round_to_int = lambda x: np.round(x)  # noqa: E731
    errs = {rule: orc.mean_error(W.astype(np.float32), orc.gptq(W, H, round_to_int, rule=rule, leaf=size), H.astype(np.float32))
            for rule in ("diag", "none", "pivot")}
    direct = orc.mean_error(W.astype(np.float32), np.round(W).astype(np.float32), H.astype(np.float32))
    assert errs["none"] <= direct
    assert errs["diag"] <= errs["none"]
    assert errs["pivot"] <= errs["none"]


def test_input_bias_removal_equals_centering_the_samples():
    # reference tests/test_obq.py:73-109 (the averaged form is the library's)
    g = np.random.default_rng(6)
    X = g.standard_normal((16, 32))
    H = X @ X.T / 32
    m = X.mean(axis=1)
    centred = (X - m[:, None]) @ (X - m[:, None]).T / 32
    assert np.allclose(orc.strip_input_bias(H, m), centred)
    assert (np.linalg.eigvalsh(orc.strip_input_bias(H, m)) >= -1e-12).all()


def test_gain_formula_equals_exhaustive_evaluation():
    # reference tests/test_obq.py:112-140
    g = np.random.default_rng(7)
    W = g.standard_normal((10, 16))
    H = _wishart(16, 10, seed=8)
    Q = np.round(W)
    cand = Q + np.square(g.standard_normal((10, 16)))
    base = orc.rowwise_error(W, Q, H)
    exhaustive = np.zeros_like(Q)
    for i in range(16):
        cur = Q.copy()
        cur[:, i] = cand[:, i]
        exhaustive[:, i] = base - orc.rowwise_error(W, cur, H)
    assert np.allclose(exhaustive, orc.flip_gain(W, Q, H, cand))


def test_scaling_quality_and_modes():
    # reference tests/test_scaling.py:130-163 (seeded): the Hessian-aware searches may not be worse
    # under the Hessian-weighted error, and every mode string of the dispatcher is accepted
    g = np.random.default_rng(9)
    size = 100
    data = g.standard_normal((20, size)).astype(np.float32)
    grid = orc.UniformGrid(9, -3, 3)
    H = _wishart(size, 10, 1.0e-6, seed=10).astype(np.float32)
    sc = {"base": orc.search_scale(data, grid, 0), "diag": orc.search_scale(data, grid, 0, H=H.diagonal()),
          "hessian": orc.search_scale(data, grid, 0, H=H), "obq": orc.search_scale_gptq(data, grid, 0, H)}
    err = {k: orc.mean_error(data, orc.quantize_scaled(data, s, grid, H=H if k == "obq" else None), H)
           for k, s in sc.items()}
    assert err["hessian"] <= err["base"] and err["hessian"] <= err["diag"] and err["obq"] <= err["hessian"]
    for mode in ("norm", "max", "mse", "diag", "hessian", "diag1", "hessian1", "diag1.8", "hessian1.8"):
        s = orc.choose_scale(data[:, :20], grid, H[:20, :20], mode=mode)
        assert s.shape == (20,) and np.all(s > 0)
    with pytest.raises(RuntimeError):
        orc.choose_scale(data[:, :20], grid, H[:20, :20], mode="bogus")
