"""GPU parity tests: the CUDA path (through the public API, hence through the C ABI) against
the committed golden vectors of the reference and against the numpy oracle on seeded inputs.

Bars (BASELINE.json north_star): rounding indices and values bit-exact; chosen scale-grid point
equal; GPTQ layer error within 1e-3 relative of the reference with the code agreement reported;
fp64 factor to fp64 round-off; local search identical from identical inputs."""

import numpy as np
import pytest
import torch

from oracle import sleekit_oracle as orc
from sleekit_b200 import workloads as wl
from tests.conftest import load_golden, record_parity, scales_equivalent

pytestmark = pytest.mark.gpu

UNI = [(2, 1), (3, 1), (4, 1), (8, 1), (16, 1), (9, 2), (9, 3)]


@pytest.fixture(scope="module")
def slk():
    import sleekit_b200
    from sleekit_b200 import codebook, obq, scaling

    class NS:
        pass

    ns = NS()
    ns.codebook, ns.obq, ns.scaling, ns.Sleekit = codebook, obq, scaling, sleekit_b200.Sleekit
    return ns


def agree(a, b):
    return float((np.asarray(a) == np.asarray(b)).mean())


def rel(a, b):
    return abs(float(a) - float(b)) / abs(float(b))


def row_error_fn(W, grid, H):
    """scales -> the oracle's per-row search error at those scales (scaling.py:84-95, :127-130)."""
    return lambda sc: orc.weighted_sq_error(H, orc.quantize_scaled(W, np.asarray(sc, dtype=np.float32), grid) - W)


# ---------------------------------------------------------------------------
# K4 rounding
# ---------------------------------------------------------------------------


@pytest.mark.parametrize("c,hi", UNI)
def test_uniform_rounding_bit_exact_vs_golden(slk, c, hi):
    g = load_golden("rounding")
    tag = f"u{c}_{hi}"
    cfg = g[tag + "_cfg"]
    lo_v, hi_v = (int(cfg[1]), int(cfg[2])) if hi != 3 else (float(cfg[1]), float(cfg[2]))
    cb = slk.codebook.UniformCodebook(int(cfg[0]), lo_v, hi_v)
    for suffix, x in (("", g[tag + "_x"]), ("64", g[tag + "_x"].astype(np.float64))):
        for name, fn in (("idx", cb.quantize_index), ("val", cb.quantize_value), ("up", cb.quantize_up),
                         ("down", cb.quantize_down)):
            got, want = fn(x), g[f"{tag}_{name}{suffix}"]
            assert got.dtype == want.dtype and got.shape == want.shape
            np.testing.assert_array_equal(got, want, err_msg=f"{tag} {name}{suffix}")


def test_table_rounding_bit_exact_vs_golden(slk):
    g = load_golden("rounding")
    cb = slk.codebook.Codebook(g["nf4_values"])
    x = g["nf4_x"]
    for name, fn in (("idx", cb.quantize_index), ("val", cb.quantize_value), ("up", cb.quantize_up),
                     ("down", cb.quantize_down)):
        got = fn(x)
        assert got.dtype == g["nf4_" + name].dtype
        np.testing.assert_array_equal(got, g["nf4_" + name])
    # reference tests/test_codebook.py:6-32 (Python lists, saturation, N-D)
    cb = slk.codebook.Codebook([-1.0, 2.0, 4.0, 8.0])
    x = [-2.0, -1.0, 0.0, 0.9, 1.9, 2.9, 3.1, 5.9, 6.1, 9.0]
    np.testing.assert_array_equal(cb.quantize_index(x), [0, 0, 0, 1, 1, 1, 2, 2, 3, 3])
    np.testing.assert_array_equal(cb(x), [-1, -1, -1, 2, 2, 2, 4, 4, 8, 8])
    np.testing.assert_array_equal(cb.quantize_up(x), [2, 2, 2, 4, 4, 4, 8, 8, 8, 8])
    np.testing.assert_array_equal(cb.quantize_down(x), [-1, -1, -1, -1, -1, -1, 2, 2, 4, 4])
    nd = np.random.default_rng(0).standard_normal((3, 4, 5))
    assert cb(nd).shape == (3, 4, 5)


def test_rounding_large_random_vs_oracle_and_idempotent(slk):
    rng = np.random.default_rng(7)
    x = (rng.standard_normal(2_000_003) * 0.7).astype(np.float32)  # odd length: vector body + scalar tail
    for c in (3, 8, 16, 300):
        cb = slk.codebook.UniformCodebook(c, -1, 1)
        ref = orc.UniformGrid(c, -1, 1)
        np.testing.assert_array_equal(cb.quantize_index(x), ref.index(x))
        v = cb(x)
        np.testing.assert_array_equal(v, ref.value(x))
        np.testing.assert_array_equal(cb(v), v)  # q(q(x)) == q(x), tests/test_codebook.py:35-40
        np.testing.assert_array_equal(cb.quantize_up(x), ref.up(x))
        np.testing.assert_array_equal(cb.quantize_down(x), ref.down(x))
    # unaligned views and empty input
    cb = slk.codebook.UniformCodebook(8, -1, 1)
    np.testing.assert_array_equal(cb(x[1:1001]), orc.UniformGrid(8, -1, 1).value(x[1:1001].copy()))
    assert cb(np.zeros((0, 5), np.float32)).shape == (0, 5)


def test_uniform_equals_table_on_fp64_3d(slk):
    # reference tests/test_codebook.py:43-57
    data = np.random.default_rng(3).standard_normal((10, 20, 30))
    ucb = slk.codebook.UniformCodebook(9, -2, 2)
    tcb = slk.codebook.Codebook.uniform(9, -2, 2)
    np.testing.assert_array_equal(ucb.quantize_index(data), tcb.quantize_index(data))
    np.testing.assert_allclose(ucb(data), tcb(data))
    np.testing.assert_allclose(ucb.quantize_up(data), tcb.quantize_up(data))
    np.testing.assert_allclose(ucb.quantize_down(data), tcb.quantize_down(data))


# ---------------------------------------------------------------------------
# scales and the scale-grid search
# ---------------------------------------------------------------------------


def test_scales_vs_golden(slk):
    g = load_golden("scales")
    W, H = g["W"], g["H"]
    S = slk.scaling
    for c in (3, 8):
        cb = slk.codebook.UniformCodebook(c, -1, 1)
        np.testing.assert_array_equal(S.compute_non_saturating_scaling(W, cb, 0), g[f"max_c{c}"])
        np.testing.assert_array_equal(S.compute_min_mse_scaling(W, cb, 0), g[f"mse_c{c}"])
        np.testing.assert_array_equal(S.compute_min_mse_scaling(W, cb, 0, H=H.diagonal()), g[f"diag_c{c}"])
        grid = orc.UniformGrid(c, -1, 1)
        scales_equivalent(S.compute_min_mse_scaling(W, cb, 0, H=H), g[f"full_c{c}"], row_error_fn(W, grid, H),
                          "test_scales_vs_golden", f"full-H search c={c}")
        np.testing.assert_array_equal(S.compute_scaling(W, cb, H, mode="diag5"), g[f"diag5_c{c}"])
        H2 = H + np.float64(0.02) * H.diagonal().mean() * np.eye(H.shape[0])          # scaling.py:222
        scales_equivalent(S.compute_scaling(W, cb, H, mode="hessian2"), g[f"hess2_c{c}"], row_error_fn(W, grid, H2),
                          "test_scales_vs_golden", f"hessian2 search c={c}")
        np.testing.assert_array_equal(S.compute_min_mse_scaling(W, cb, 1, grid_size=17, min_factor=0.2),
                                      g[f"axis1_c{c}"])
    np.testing.assert_allclose(S.compute_norm_scaling(W, 0), g["norm0"], rtol=2e-7)
    np.testing.assert_allclose(S.compute_norm_scaling(W, 1), g["norm1"], rtol=2e-7)
    np.testing.assert_array_equal(
        S.compute_non_saturating_scaling(W, slk.codebook.Codebook([-1.0, 0.0, 10.0, 20.0]), 0), g["max_tab0"])
    cb = slk.codebook.UniformCodebook(8, -1, 1)
    np.testing.assert_array_equal(S.quantize_with_scaling(W, g["diag_c8"], cb), g["qws_plain"])
    np.testing.assert_array_equal(S.apply_scaling(W, g["diag_c8"], 0), g["apply"])


def test_scale_known_answers(slk):
    # reference tests/test_scaling.py:16-41, 56-72
    S = slk.scaling
    data = np.array([[0.0, 10.0], [5.0, 5.0]], dtype=np.float32)
    sc = S.compute_norm_scaling(data, 0)
    np.testing.assert_allclose(sc, [10.0 / np.sqrt(2), 5.0], rtol=1e-6)
    np.testing.assert_allclose(S.apply_scaling(data, sc, 0), [[0.0, np.sqrt(2)], [1.0, 1.0]], rtol=1e-6)
    np.testing.assert_allclose(S.apply_scaling(S.apply_scaling(data, sc, 0), 1 / sc, 0), data, rtol=1e-6)
    sc = S.compute_norm_scaling(data, 1)
    np.testing.assert_allclose(sc, [5.0 / np.sqrt(2), np.sqrt(125 / 2)], rtol=1e-6)
    np.testing.assert_allclose(S.apply_scaling(data, sc, 1),
                               [[0.0, 10.0 / np.sqrt(125 / 2)], [np.sqrt(2), 5.0 / np.sqrt(125 / 2)]], rtol=1e-6)
    data = np.array(
        [[0.0, 10.0, -20.0, 15.0], [5.0, 5.0, 10.0, -10.0], [1.0, 2.0, -4.0, 3.0], [0.0, 0.0, 0.0, 0.0],
         [1.0, 10.0, 100.0, 1000.0], [-1.0, 10.0, 100.0, 1000.0]], dtype=np.float32)
    cb = slk.codebook.Codebook([-1.0, 0.0, 10.0, 20.0])
    np.testing.assert_allclose(S.compute_non_saturating_scaling(data, cb, 0), [20, 10, 4, 1e-16, 50, 50])
    np.testing.assert_allclose(S.compute_non_saturating_scaling(data, cb, 1), [1, 0.5, 20, 50])
    nd = np.random.default_rng(1).standard_normal((10, 20, 30, 40)).astype(np.float32)
    ucb = slk.codebook.UniformCodebook(9, -2, 2)
    for ax, ln in enumerate((10, 20, 30, 40)):
        assert len(S.compute_norm_scaling(nd, ax)) == ln
        assert len(S.compute_non_saturating_scaling(nd, ucb, ax)) == ln
        np.testing.assert_array_equal(S.compute_non_saturating_scaling(nd, ucb, ax),
                                      orc.no_clip_scale(nd, orc.UniformGrid(9, -2, 2), ax))


@pytest.mark.parametrize("r,n,c", [(96, 768, 8), (64, 3072, 3), (33, 1000, 4)])
def test_scale_search_selects_reference_grid_point(slk, r, n, c):
    W, H, m = wl.synthetic_layer(r, n, 5, samples=256)
    cb = slk.codebook.UniformCodebook(c, -1, 1)
    grid = orc.UniformGrid(c, -1, 1)
    for Harg in (None, H.diagonal().copy()):
        got = slk.scaling.compute_min_mse_scaling(W, cb, 0, H=Harg)
        want = orc.search_scale(W, grid, 0, H=Harg)
        assert got.dtype == np.float32
        scales_equivalent(got, want, row_error_fn(W, grid, Harg), "test_scale_search_selects_reference_grid_point",
                          f"[{r},{n}] c={c} {'mse' if Harg is None else 'diag'}")
    # fp64 diagonal (reference tests/test_scaling.py:97-105): errors accumulate in fp64
    h64 = np.random.default_rng(2).random(n)
    scales_equivalent(slk.scaling.compute_min_mse_scaling(W, cb, 0, H=h64), orc.search_scale(W, grid, 0, H=h64),
                      row_error_fn(W, grid, h64), "test_scale_search_selects_reference_grid_point",
                      f"[{r},{n}] c={c} fp64 diag")


@pytest.mark.parametrize("c", [2, 3, 4, 8, 16])
@pytest.mark.parametrize("diag", [False, True])
def test_scale_search_table_form_equals_direct_form(slk, c, diag):
    """The threshold-table kernel assigns every weight the code the reference's op chain assigns
    (exact breakpoints by bisection), so it must pick the same grid point as the direct kernel;
    the reported best errors differ only by summation order."""
    from sleekit_b200 import _lib, ops

    r, n = 300, 1111
    W, H, m = wl.synthetic_layer(r, n, 3, samples=256)
    W[5, :] = 0                        # all-zero row: init = 1e-16
    W[6, :17] *= 40                    # outliers
    W[7] = np.abs(W[7])                # one-signed row
    cb = slk.codebook.UniformCodebook(c, -1, 1)
    Wd = torch.from_numpy(W).cuda()
    f = torch.linspace(0.05, 1.0, 100, device="cuda")
    hd = torch.from_numpy(np.ascontiguousarray(H.diagonal())).cuda() if diag else None
    out = {}
    for direct in (1, 2, 0):         # direct op chain / tables + estimate loop / tables + walk (default)
        _lib.call("slk_debug_scale_search_direct", direct)
        try:
            sc, err, init = ops.scale_search(Wd, cb, f, hd, want_err=True, want_init=True)
            out[direct] = (sc.cpu().numpy(), err.cpu().numpy(), init.cpu().numpy())
        finally:
            _lib.call("slk_debug_scale_search_direct", 0)
    for form in (0, 2):
        np.testing.assert_array_equal(out[form][2], out[1][2])
        same = float((out[form][0] == out[1][0]).mean())
        print(f"c={c} diag={diag} form={form}: same grid point in {same:.4f} of rows")
        assert same >= 0.995          # ties between near-equal errors may flip with summation order
        np.testing.assert_allclose(out[form][1], out[1][1], rtol=2e-5)


@pytest.mark.parametrize("c,grid", [(8, "g16"), (8, "g17"), (8, "g15"), (8, "g128"), (16, "coarse"), (4, "descending"),
                                    (8, "wide"), (3, "g33")])
def test_scale_search_walk_form_grids(slk, c, grid):
    """The walk form of the table kernel (index tracked along the grid, conditions checked on the exact
    tables) against the direct op chain on grids that exercise the 16-point blocks, the tail, and the
    rows that must fall back to the estimate loop (steps of more than one code per grid point,
    non-ascending grids)."""
    from sleekit_b200 import _lib, ops

    r, n = 130, 1000
    W, H, m = wl.synthetic_layer(r, n, 11, samples=256)
    W[3, :] = 0
    W[4, :9] *= 60
    W[5] = -np.abs(W[5])
    cb = slk.codebook.UniformCodebook(c, -1, 1)
    f = {"g16": np.linspace(0.05, 1.0, 16), "g17": np.linspace(0.3, 1.0, 17), "g15": np.linspace(0.3, 1.0, 15),
         "g128": np.linspace(0.05, 1.0, 128), "coarse": np.linspace(0.02, 1.0, 9),
         "descending": np.linspace(1.0, 0.05, 40), "wide": np.linspace(0.01, 3.0, 100),
         "g33": np.linspace(0.05, 1.0, 33)}[grid].astype(np.float32)
    Wd, fd = torch.from_numpy(W).cuda(), torch.from_numpy(f).cuda()
    hd = torch.from_numpy(np.ascontiguousarray(H.diagonal())).cuda()
    out = {}
    for direct in (1, 0):
        _lib.call("slk_debug_scale_search_direct", direct)
        try:
            sc, err, init = ops.scale_search(Wd, cb, fd, hd, want_err=True, want_init=True)
            out[direct] = (sc.cpu().numpy(), err.cpu().numpy())
        finally:
            _lib.call("slk_debug_scale_search_direct", 0)
    same = float((out[0][0] == out[1][0]).mean())
    print(f"c={c} grid={grid}: same grid point in {same:.4f} of rows")
    assert same >= 0.99
    np.testing.assert_allclose(out[0][1], out[1][1], rtol=2e-5)


def test_full_h_scale_search_vs_oracle(slk):
    W, H, m = wl.synthetic_layer(48, 256, 9, samples=512)
    cb = slk.codebook.UniformCodebook(3, -1, 1)
    got = slk.scaling.compute_min_mse_scaling(W, cb, 0, H=H)
    grid = orc.UniformGrid(3, -1, 1)
    want = orc.search_scale(W, grid, 0, H=H)
    scales_equivalent(got, want, row_error_fn(W, grid, H), "test_full_h_scale_search_vs_oracle", "[48,256] c=3 fp32 H")
    H64 = orc.strip_input_bias(H, m).astype(np.float64)
    got = slk.scaling.compute_min_mse_scaling(W, cb, 0, H=H64)
    want = orc.search_scale(W, grid, 0, H=H64)
    scales_equivalent(got, want, row_error_fn(W, grid, H64), "test_full_h_scale_search_vs_oracle",
                      "[48,256] c=3 fp64 H - m m^T")


@pytest.mark.parametrize("r,n,c,samples", [(200, 1024, 8, 2048), (96, 4096, 16, 2048), (130, 300, 3, 128),
                                           (64, 2048, 4, 512)])
def test_full_h_search_screening_is_exact(slk, r, n, c, samples):
    """The full-H search ranks all grid points with one low-precision pass (bf16, or one TF32 pass) and
    evaluates only the best-ranked candidates exactly (dense.cu): scales AND errors must be bit-identical to
    evaluating every grid point (fullh_topk = 0), for 4 / 8 / 16 candidates (fixed per row, or compacted to
    those within 2^-5 of the best-ranked one), both operand types and both tile widths of the screening product; the
    all-points result is itself compared with the oracle (scaling.py:98-134) on a slice of the rows.
    Rows of zeros and a row of tiny weights (every grid point gives the same error: first one wins)."""
    from sleekit_b200 import ops, _convert as cv
    from sleekit_b200.scaling import _factors

    W, H, m = wl.synthetic_layer(r, n, 23, samples=samples)
    W[3] = 0.0
    W[5] *= 1e-30
    cb, grid = slk.codebook.UniformCodebook(c, -1, 1), orc.UniformGrid(c, -1, 1)
    Wd, Hd = cv.to_dev(W, torch.float32), cv.to_dev(H, torch.float32)
    f = _factors(0.05, 1.0, 100, Wd.device)
    res = {}
    try:
        for topk, bn, bf16, compact in [(0, 256, 1, 1), (8, 256, 1, 1), (8, 128, 1, 0), (4, 256, 1, 1), (16, 128, 1, 1),
                                        (8, 256, 0, 1), (4, 128, 0, 0), (16, 256, 1, 0)]:
            ops.set_option("fullh_compact", compact)
            ops.set_option("fullh_ctas", 1 if (topk, bn) == (8, 256) else 2)
            ops.set_option("fullh_topk", topk)
            ops.set_option("fullh_bn", bn)
            ops.set_option("fullh_bf16", bf16)
            sc, err = ops.scale_search_fullh(Wd, cb, f, Hd, want_err=True)
            res[(topk, bn, bf16, compact)] = (sc.cpu().numpy(), err.cpu().numpy())
    finally:
        ops.set_option("fullh_topk", 8)
        ops.set_option("fullh_bn", 256)
        ops.set_option("fullh_bf16", 1)
        ops.set_option("fullh_ctas", 2)
        ops.set_option("fullh_compact", 1)
    sc_c, err_c, bad = ops.scale_search_fullh(Wd, cb, f, Hd, want_err=True, want_check=True)
    assert int(bad.item()) == 0, "rows without a certificate"          # and the certificate itself (fullh_certify_kernel)
    res[("default", "checked")] = (sc_c.cpu().numpy(), err_c.cpu().numpy())
    base = res[(0, 256, 1, 1)]
    for key, (sc, err) in res.items():
        same = float((sc == base[0]).mean())
        print(f"[{r}x{n} c={c}] topk, tile, bf16, compacted = {key}: scales identical in {same:.6f} of rows")
        np.testing.assert_array_equal(sc, base[0], err_msg=str(key))
        np.testing.assert_array_equal(err, base[1], err_msg=str(key))
    rows = slice(0, 24)
    want = orc.search_scale(W[rows], grid, 0, H=H)
    scales_equivalent(base[0][rows], want, row_error_fn(W[rows], grid, H), "test_full_h_search_screening_is_exact",
                      f"[{r},{n}] c={c}")


def test_full_h_search_screening_many_flat_rows(slk):
    """Rows whose grid points all tie (all-zero rows of a pruned layer) ask for the maximum number of
    candidates each; when that exceeds the candidate capacity every row is cut to its best-ranked ones
    (screen_pairs_kernel) -- no row may be left without candidates, and the result stays bit-identical to
    evaluating every grid point (the first grid point wins the ties, scaling.py:131)."""
    from sleekit_b200 import ops, _convert as cv
    from sleekit_b200.scaling import _factors

    r, n = 300, 1024
    W, H, m = wl.synthetic_layer(r, n, 29, samples=1024)
    W[20:260] = 0.0                                   # 80 % of the rows
    W[5, ::2] = 0.0
    for c in (3, 8):
        cb = slk.codebook.UniformCodebook(c, -1, 1)
        Wd, Hd = cv.to_dev(W, torch.float32), cv.to_dev(H, torch.float32)
        f = _factors(0.05, 1.0, 100, Wd.device)
        res = {}
        try:
            for topk in (0, 4, 8, 16):
                ops.set_option("fullh_topk", topk)
                sc, err = ops.scale_search_fullh(Wd, cb, f, Hd, want_err=True)
                res[topk] = (sc.cpu().numpy(), err.cpu().numpy())
        finally:
            ops.set_option("fullh_topk", 8)
        assert np.all(np.isfinite(res[0][0]))
        assert int(ops.scale_search_fullh(Wd, cb, f, Hd, want_check=True)[2].item()) == 0
        for topk in (4, 8, 16):
            np.testing.assert_array_equal(res[topk][0], res[0][0], err_msg=f"c={c} topk={topk}")
            np.testing.assert_array_equal(res[topk][1], res[0][1], err_msg=f"c={c} topk={topk}")
        want = orc.search_scale(W[:40], orc.UniformGrid(c, -1, 1), 0, H=H)
        scales_equivalent(res[0][0][:40], want, row_error_fn(W[:40], orc.UniformGrid(c, -1, 1), H),
                          "test_full_h_search_screening_many_flat_rows", f"[{r},{n}] c={c}, 80 % zero rows")


# ---------------------------------------------------------------------------
# K2 factor, ordering
# ---------------------------------------------------------------------------


def test_factor_vs_golden_and_identities(slk):
    g = load_golden("sweep")
    U = slk.obq.compute_hessian_chol(g["Hd"])
    assert U.dtype == np.float64 and U.flags["C_CONTIGUOUS"]
    np.testing.assert_allclose(U, g["U"], rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(np.tril(U, -1), 0)
    np.testing.assert_array_equal(slk.obq.remove_input_bias(g["H"], g["mean"]), g["Hc"])
    # reference tests/test_obq.py:21-32
    rng = np.random.default_rng(4)
    A = rng.standard_normal((4, 6)).astype(np.float32)
    H = A @ A.T + 1e-6 * np.eye(4)
    U = slk.obq.compute_hessian_chol(H)
    np.testing.assert_allclose(U, np.linalg.cholesky(np.linalg.inv(H)).T)
    np.testing.assert_allclose(np.linalg.inv(U.T @ U), H)


@pytest.mark.parametrize("n", [50, 64, 200, 768, 1000])
def test_factor_vs_oracle(slk, n):
    _, H, _ = wl.synthetic_layer(4, n, 11, samples=max(2 * n, 256))
    Hd = H + 0.01 * H.diagonal().mean() * np.eye(n)
    U = slk.obq.compute_hessian_chol(Hd)
    want = orc.inverse_upper_factor(Hd)
    scale = np.abs(want).max()
    assert np.abs(U - want).max() <= 1e-9 * scale
    np.testing.assert_allclose(U.T @ U @ Hd, np.eye(n), atol=1e-7)


@pytest.mark.parametrize("n,perm,damp", [(50, False, 0.01), (64, True, 0.01), (200, True, 0.03), (768, True, 0.01),
                                         (1000, True, 0.01), (3072, True, 0.01)])
def test_cholesky_form_vs_oracle(slk, n, perm, damp):
    """K2 without the inverse (tile-task Cholesky): R = flip(chol(flip(H_opt))) = inv(U) and the
    32x32 diagonal blocks of U, against the oracle's factor (obq.py:38-55) -- fp64 work, fp32 out."""
    from sleekit_b200 import ops

    _, H, _ = wl.synthetic_layer(4, n, 13, samples=max(min(2 * n, 4096), 256))
    dampval = np.float32(damp) * H.diagonal().mean()
    order = np.argsort(-H.diagonal().astype(np.float64), kind="stable") if perm else np.arange(n)
    Hd = (H.astype(np.float64) + np.float64(dampval) * np.eye(n))[order][:, order]
    U = orc.inverse_upper_factor(Hd)
    L = np.linalg.cholesky(Hd[::-1, ::-1])
    R = np.ascontiguousarray(L[::-1, ::-1])           # upper, Hd = R R^T, U = inv(R)
    Hdev = torch.from_numpy(H).cuda()
    r32, rt32, ud32, info = ops.chol_factor(Hdev, torch.from_numpy(order).cuda() if perm else None,
                                            torch.tensor([dampval], dtype=torch.float32, device="cuda"))
    assert int(info.item()) == 0
    r32, rt32, ud32 = r32.cpu().numpy(), rt32.cpu().numpy(), ud32.cpu().numpy()
    np.testing.assert_array_equal(np.tril(r32, -1), 0)
    # TF32 hi + lo parts; the part above the diagonal is never written (nor read by the sweep): mask it first
    np.testing.assert_allclose(np.tril(rt32[0]) + np.tril(rt32[1]), r32.T, rtol=3e-7, atol=0)
    assert np.abs(r32 - R).max() <= 2e-7 * np.abs(R).max()
    for b in range((n + 31) // 32):
        w = min(32, n - 32 * b)
        want = np.eye(32)
        want[:w, :w] = U[32 * b:32 * b + w, 32 * b:32 * b + w]
        assert np.abs(ud32[b] - want).max() <= 2e-7 * np.abs(want).max(), b
    # the factor reproduces H_opt to fp32 round-off of R
    Rd = r32.astype(np.float64)
    assert np.abs(Rd @ Rd.T - Hd).max() <= 1e-6 * np.abs(Hd).max()


def test_batched_factor_equals_single_factor(slk):
    """slk_chol_factor_batched_f32 (one ticket queue over the tile tasks of several matrices) gives, bit
    for bit, what slk_chol_factor_f32 gives per matrix: same tasks, same arithmetic, another schedule."""
    from sleekit_b200 import ops

    for n, B in ((200, 3), (768, 5), (1100, 2)):
        hs, orders, damps = [], [], []
        for k in range(B):
            _, H, _ = wl.synthetic_layer(4, n, 20 + k, samples=max(min(2 * n, 2048), 256))
            Hd = torch.from_numpy(H).cuda()
            dv = ops.damp_value(Hd, 0.01)
            hs.append(Hd)
            damps.append(dv)
            orders.append(ops.argsort(ops.order_keys(Hd, dv, None)) if k % 2 == 0 else None)
        got = ops.chol_factor_batched(hs, orders, damps)
        for k in range(B):
            r32, rt, ud32, info = ops.chol_factor(hs[k], orders[k], damps[k])
            assert int(info.item()) == 0 and int(got[k][3].item()) == 0
            assert torch.equal(got[k][0], r32) and torch.equal(got[k][2], ud32)
            assert torch.equal(torch.tril(got[k][1][0]), torch.tril(rt[0]))
            assert torch.equal(torch.tril(got[k][1][1]), torch.tril(rt[1]))
    # a non-PD member is reported for that matrix only
    bad = hs[0].clone()
    bad[5, 5] = -1.0
    out = ops.chol_factor_batched([hs[0], bad], None, None)
    assert int(out[0][3].item()) == 0 and int(out[1][3].item()) != 0


def test_symmetric_pack_roundtrip(slk):
    """slk_sym_pack_f32 / slk_sym_unpack_f32: the block-upper-triangle exchange format of the
    sample-sharded statistics reproduces a symmetric matrix bit for bit (scale 1) and scales exactly."""
    from sleekit_b200 import ops

    for n in (32, 100, 1000, 1100, 3072):
        _, H, _ = wl.synthetic_layer(4, n, 3, samples=256)
        Hd = torch.from_numpy(H).cuda()
        Hd = ((Hd + Hd.T) * 0.5).contiguous()
        L = ops.sym_packed_len(n)
        buf = torch.empty(L + 7, dtype=torch.float32, device="cuda")
        assert ops.sym_pack(Hd, buf, 1.0) == L
        back = torch.full_like(Hd, float("nan"))
        ops.sym_unpack(buf, back, 1.0)
        assert torch.equal(back, Hd)
        ops.sym_pack(Hd, buf, 0.25)
        ops.sym_unpack(buf, back, 4.0)
        assert torch.equal(back, Hd)


def test_sweep_rows_do_not_depend_on_cta_tiling(slk):
    """The macro-block sweep sums a row's products in an order that does not depend on how many rows a
    CTA takes (8 / 16 / 32): the same layer quantized with different tile heights -- what the layer-set
    driver, a row slice or a row-sharded run use -- gives the same bits."""
    from sleekit_b200 import ops

    r, n = 768, 1024
    W, H, _ = wl.synthetic_layer(r, n, 29, samples=1024)
    cb = slk.codebook.UniformCodebook(8, -1, 1)
    Wd, Hd = torch.from_numpy(W).cuda(), torch.from_numpy(H).cuda()
    sc = slk.scaling.compute_min_mse_scaling(Wd, cb, 0, H=Hd.diagonal().contiguous())
    outs = []
    try:
        for want in (0, 24, 96, 100000):
            ops.set_option("sweep_ctas", want)
            outs.append(slk.scaling.quantize_with_scaling(Wd, sc, cb, H=Hd))
    finally:
        ops.set_option("sweep_ctas", 0)
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    part = slk.scaling.quantize_with_scaling(Wd[:40].contiguous(), sc[:40].contiguous(), cb, H=Hd)
    assert torch.equal(part, outs[0][:40])


def test_cholesky_form_sweep_equals_inverse_form(slk):
    """quantize_opt through (Cholesky factor, R-form sweep) and through (inverse factor, U-form
    sweep): the same algebra in different fp32 rounding -- codes agree to the GPTQ noise floor."""
    W, H, m = wl.synthetic_layer(256, 1024, 5)
    cb, grid = slk.codebook.UniformCodebook(8, -1, 1), orc.UniformGrid(8, -1, 1)
    sc = orc.search_scale(W, grid, 0, H=H.diagonal())
    outs = []
    for flag in (True, False):
        old, slk.obq.USE_CHOL_FORM = slk.obq.USE_CHOL_FORM, flag
        try:
            outs.append(slk.scaling.quantize_with_scaling(W, sc, cb, H=H))
        finally:
            slk.obq.USE_CHOL_FORM = old
    a = agree(grid.index(orc.divide_rows(outs[0], sc, 0)), grid.index(orc.divide_rows(outs[1], sc, 0)))
    e0, e1 = orc.mean_error(W, outs[0], H), orc.mean_error(W, outs[1], H)
    print(f"chol form vs inverse form: code agreement {a:.6f}, errors {e0:.6e} {e1:.6e}")
    assert a >= 0.999 and rel(e0, e1) <= 1e-3


@pytest.mark.parametrize("r,n,c,rule,samples", [(768, 768, 8, "diag", 2048), (96, 3072, 8, "diag", 2048),
                                                  (40, 300, 4, "none", 512), (50, 1100, 3, "sqerr", 256),
                                                  (33, 514, 8, "diag", 2048), (3072, 768, 8, "diag", 2048)])
def test_layer_error_from_sweep_residuals_equals_product(slk, r, n, c, rule, samples):
    """gptq_device(want_err=True): the layer error taken from the sweep (sum E^2 - damp * sum D^2, exact
    algebra: W - Q = E U, U H_opt U^T = I) against the explicit product ((W-Q) H (W-Q)^T) of K6 and
    against the oracle's quantization_error (obq.py:89-103) on the same quantized weights -- with
    rank-deficient Hessians (samples < n: the damping term carries a large share) and every sweep
    kernel (macro blocks, single fused launch for n < 512 or n % 4 != 0)."""
    from sleekit_b200 import _convert as cv

    W, H, m = wl.synthetic_layer(r, n, 3, samples=samples)
    cb, grid = slk.codebook.UniformCodebook(c, -1, 1), orc.UniformGrid(c, -1, 1)
    sc = orc.search_scale(W, grid, 0, H=H.diagonal())
    Wd, Hd, sd = cv.to_dev(W, torch.float32), cv.to_dev(H, torch.float32), cv.to_dev(sc, torch.float32)
    for scaled in (True, False):
        q, (err, rows) = slk.obq.gptq_device(Wd, Hd, cb, rule, 0.01, row_scale=sd if scaled else None, want_err=True)
        old, slk.obq.USE_SWEEP_ERROR = slk.obq.USE_SWEEP_ERROR, False
        try:
            q2, (err2, rows2) = slk.obq.gptq_device(Wd, Hd, cb, rule, 0.01, row_scale=sd if scaled else None,
                                                    want_err=True)
        finally:
            slk.obq.USE_SWEEP_ERROR = old
        assert torch.equal(q, q2)
        rows, rows2 = rows.cpu().numpy(), rows2.cpu().numpy()
        worst = float(np.max(np.abs(rows - rows2) / np.abs(rows2)))
        e_ref = orc.mean_error(W, q.cpu().numpy(), H)
        print(f"[{r}x{n} {rule} scaled={scaled}] error {float(err):.6e} product {float(err2):.6e} oracle {e_ref:.6e} "
              f"worst row {worst:.2e}")
        assert rel(err, err2) <= 1e-4 and worst <= 5e-4
        assert rel(err, e_ref) <= 1e-4


@pytest.mark.parametrize("r,n,c,moves", [(64, 1024, 4, 10), (24, 4096, 16, 10), (40, 300, 3, 50)])
def test_layer_error_after_local_search_equals_product(slk, r, n, c, moves):
    """gptq_device(want_err=True, nb_ls_moves > 0): the row errors taken from the local search's own
    p = (Q - W) H after its moves (p . (Q - W), local_search.cu) against the explicit K6 product and the
    oracle's quantization_error (obq.py:89-103) on the same weights, with and without row scales."""
    from sleekit_b200 import _convert as cv

    W, H, m = wl.synthetic_layer(r, n, 17, samples=max(256, n // 2))
    cb, grid = slk.codebook.UniformCodebook(c, -1, 1), orc.UniformGrid(c, -1, 1)
    sc = orc.search_scale(W, grid, 0, H=H.diagonal())
    Wd, Hd, sd = cv.to_dev(W, torch.float32), cv.to_dev(H, torch.float32), cv.to_dev(sc, torch.float32)
    for scaled in (True, False):
        args = dict(row_scale=sd if scaled else None, want_err=True, nb_ls_moves=moves)
        q, (err, rows) = slk.obq.gptq_device(Wd, Hd, cb, "diag", 0.01, **args)
        old, slk.obq.USE_SWEEP_ERROR = slk.obq.USE_SWEEP_ERROR, False
        try:
            q2, (err2, rows2) = slk.obq.gptq_device(Wd, Hd, cb, "diag", 0.01, **args)
        finally:
            slk.obq.USE_SWEEP_ERROR = old
        assert torch.equal(q, q2)
        rows, rows2 = rows.cpu().numpy(), rows2.cpu().numpy()
        worst = float(np.max(np.abs(rows - rows2) / np.abs(rows2)))
        e_ref = orc.mean_error(W, q.cpu().numpy(), H)
        print(f"[{r}x{n} c={c} scaled={scaled}] error {float(err):.6e} product {float(err2):.6e} oracle {e_ref:.6e} "
              f"worst row {worst:.2e}")
        assert rel(err, err2) <= 2e-5 and worst <= 1e-4
        assert rel(err, e_ref) <= 1e-4


def test_factor_not_positive_definite_raises(slk):
    H = np.eye(70)
    H[40, 40] = -1.0
    with pytest.raises(np.linalg.LinAlgError):
        slk.obq.compute_hessian_chol(H)
    with pytest.raises(np.linalg.LinAlgError):
        W = np.ones((4, 70), np.float32)
        slk.obq.quantize_opt(W, H.astype(np.float32), slk.codebook.UniformCodebook(4, -1, 1), damp=0.0)


def test_ordering_vs_golden(slk):
    g = load_golden("sweep")
    cb = slk.codebook.UniformCodebook(8, -1, 1)
    for rule in ("diag", "none", "err", "sqerr"):
        got = slk.obq.compute_hessian_order(g["Ws"], g["Hd"], cb, rule)
        assert got.dtype == np.int64
        np.testing.assert_array_equal(got, g[f"order_{rule}"], err_msg=rule)
    with pytest.raises(RuntimeError):
        slk.obq.compute_hessian_order(g["Ws"], g["Hd"], cb, "bogus")
    for rule in ("inv_diag", "combined_diag"):
        np.testing.assert_array_equal(slk.obq.compute_hessian_order(g["Ws"], g["Hd"], cb, rule),
                                      orc.column_order(g["Ws"], g["Hd"], orc.UniformGrid(8, -1, 1), rule))


def test_pivot_ordering_vs_golden_and_oracle(slk):
    """act_order = "pivot": the device ordering performs the reference's fp64 operations on the same
    operands (obq.py:140-166), so the permutation is the reference's; GPTQ with it agrees as usual."""
    g = load_golden("pivot")
    cb, grid = slk.codebook.UniformCodebook(8, -1, 1), orc.UniformGrid(8, -1, 1)
    np.testing.assert_array_equal(slk.obq.compute_hessian_order(g["Ws"], g["Hd"], cb, "pivot"), g["order_pivot"])
    got = slk.obq.quantize_opt(g["Ws"], g["H"], cb, act_order="pivot", damp=0.01)
    assert agree(got, g["gptq_pivot"]) >= 0.995
    _, H, _ = wl.synthetic_layer(4, 300, 17, samples=1024)
    Hd = H.astype(np.float64) + 0.01 * H.diagonal().mean() * np.eye(300)
    W = np.zeros((2, 300), np.float32)
    np.testing.assert_array_equal(slk.obq.compute_hessian_order(W, Hd, cb, "pivot"), orc.greedy_pivot_order(Hd))


# ---------------------------------------------------------------------------
# K3 sweep / GPTQ
# ---------------------------------------------------------------------------


def test_sweep_with_reference_factor_vs_golden(slk):
    g = load_golden("sweep")
    cb = slk.codebook.UniformCodebook(8, -1, 1)
    Q = g["Ws"].copy()
    E = np.zeros_like(Q)
    slk.obq._quantize_opt_block(Q, E, g["U"], cb, 32, 8)
    assert agree(Q, g["sweep_Q"]) >= 0.999
    np.testing.assert_allclose(E, g["sweep_E"], rtol=2e-3, atol=1e-5)
    # a single 32-wide leaf has no GEMM in it: bit-exact with the reference arithmetic
    Q1 = g["Ws"][:, :32].copy()
    E1 = np.zeros_like(Q1)
    U1 = np.ascontiguousarray(g["U"][:32, :32])
    slk.obq._quantize_opt_core(Q1, E1, U1, cb)
    Qr = g["Ws"][:, :32].copy()
    Er = np.zeros_like(Qr)
    orc.sweep_in_place(Qr, Er, U1, orc.UniformGrid(8, -1, 1))
    np.testing.assert_array_equal(Q1, Qr)
    np.testing.assert_array_equal(E1, Er)


def test_gptq_vs_golden(slk):
    g = load_golden("sweep")
    cb = slk.codebook.UniformCodebook(8, -1, 1)
    grid = orc.UniformGrid(8, -1, 1)
    Ws, H = g["Ws"], g["H"]
    cases = [(dict(act_order=r, damp=0.01), f"gptq_{r}") for r in ("diag", "none", "err", "sqerr")]
    cases += [(dict(act_order="sqerr", damp=0.03), "gptq_damp3"),
              (dict(act_order="diag", damp=0.01, nb_ls_moves=20), "gptq_ls20")]
    for kw, key in cases:
        got = slk.obq.quantize_opt(Ws, H, cb, **kw)
        assert got.dtype == np.float32
        np.testing.assert_array_equal(grid.value(got), got)  # every output is a codeword
        a = agree(got, g[key])
        e_got, e_ref = orc.mean_error(Ws, got, H), orc.mean_error(Ws, g[key], H)
        record_parity("test_gptq_vs_golden", f"{key}: code agreement", a, 0.999)
        record_parity("test_gptq_vs_golden", f"{key}: layer error rel diff", rel(e_got, e_ref), 1e-3)
        assert a >= 0.999 and rel(e_got, e_ref) <= 1e-3, (key, a, e_got, e_ref)
    got = slk.scaling.quantize_with_scaling(g["W"], g["scale"], cb, H=H, act_order="diag", damp=0.01)
    record_parity("test_gptq_vs_golden", "qws_gptq: weights equal", agree(got, g["qws_gptq"]), 0.999)
    assert agree(got, g["qws_gptq"]) >= 0.999
    np.testing.assert_allclose(slk.obq.channelwise_error(g["W"], g["qws_gptq"], H), g["err_rows"], rtol=1e-4)
    e = slk.obq.quantization_error(g["W"], g["qws_gptq"], H)
    assert isinstance(e, np.float32) and rel(e, g["err_mean"]) < 1e-5
    cb4 = slk.codebook.UniformCodebook(4, -1, 1)
    got = slk.scaling.quantize_with_scaling(g["W2"], g["scale2"], cb4, H=g["H2"])  # ragged 200-column recursion
    record_parity("test_gptq_vs_golden", "qws2 (ragged 200 columns): weights equal", agree(got, g["qws2"]), 0.999)
    assert agree(got, g["qws2"]) >= 0.999
    got = slk.scaling.quantize_with_scaling(g["W2"], g["scale2"], cb4, H=g["H2"], nb_ls_moves=15)
    record_parity("test_gptq_vs_golden", "qws2 + 15 moves: weights equal", agree(got, g["qws2_ls"]), 0.999)
    assert agree(got, g["qws2_ls"]) >= 0.999


@pytest.mark.parametrize("r,n,c,rule,damp", [(768, 768, 8, "diag", 0.01), (128, 3072, 8, "diag", 0.01),
                                               (200, 1024, 3, "sqerr", 0.03), (64, 1100, 4, "err", 0.01)])
def test_gptq_stagewise_vs_oracle(slk, r, n, c, rule, damp):
    """Same W, H and scales into both implementations (BASELINE config 1 is the first case)."""
    W, H, m = wl.synthetic_layer(r, n, 0)
    cb, grid = slk.codebook.UniformCodebook(c, -1, 1), orc.UniformGrid(c, -1, 1)
    sc = orc.search_scale(W, grid, 0, H=H.diagonal())
    want = orc.quantize_scaled(W, sc, grid, H=H, rule=rule, damp=damp)
    got = slk.scaling.quantize_with_scaling(W, sc, cb, H=H, act_order=rule, damp=damp)
    codes_w = grid.index(orc.divide_rows(want, sc, 0))
    codes_g = grid.index(orc.divide_rows(got, sc, 0))
    a = agree(codes_g, codes_w)
    e_got, e_ref = orc.mean_error(W, got, H), orc.mean_error(W, want, H)
    print(f"[{r}x{n} c={c} {rule}] code agreement {a:.6f}  layer error {e_got:.6e} vs {e_ref:.6e}")
    assert a >= 0.999
    assert rel(e_got, e_ref) <= 1e-3
    # GPTQ must beat plain rounding at the same scales (reference tests/test_obq.py:35-54)
    assert e_got <= orc.mean_error(W, orc.quantize_scaled(W, sc, grid), H)


# ---------------------------------------------------------------------------
# K7 local search, gains
# ---------------------------------------------------------------------------


def test_local_search_vs_golden(slk):
    g = load_golden("local_search")
    cb = slk.codebook.UniformCodebook(4, -1, 1)
    Ws, H, Q0 = g["Ws"], g["H"], g["Q0"]
    np.testing.assert_allclose(slk.obq.compute_gain(Ws, Q0, H, cb.quantize_up(Q0)), g["gain_up"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(slk.obq.compute_gain(Ws, Q0, H, cb.quantize_down(Q0)), g["gain_down"], rtol=1e-4,
                               atol=1e-6)
    for k in (1, 5, 30):
        got = slk.obq.quantize_local_search(Ws, Q0, H, cb, k)
        np.testing.assert_array_equal(got, g[f"ls_{k}"], err_msg=f"{k} moves")
    assert slk.obq.quantize_local_search(Ws, Q0, H, cb, 0) is Q0
    ls = slk.obq.LocalSearchQuantizer(Ws, Q0, H, cb)
    for _ in range(5):
        ls.do_move()
    np.testing.assert_array_equal(ls.Q, g["ls_5"])
    np.testing.assert_allclose(ls.err, orc.rowwise_error(Ws, g["ls_5"], H), rtol=1e-4)


@pytest.mark.parametrize("r,n,c,moves", [(64, 1024, 4, 40), (16, 4096, 3, 25)])
def test_local_search_vs_oracle(slk, r, n, c, moves):
    W, H, m = wl.synthetic_layer(r, n, 3, samples=512)
    grid, cb = orc.UniformGrid(c, -1, 1), slk.codebook.UniformCodebook(c, -1, 1)
    sc = orc.no_clip_scale(W, grid, 0) * np.float32(0.5)
    Ws = orc.divide_rows(W, sc, 0)
    Q0 = grid.value(Ws)
    want = orc.local_search(Ws, Q0, H, grid, moves)
    got = slk.obq.quantize_local_search(Ws, Q0, H, cb, moves)
    a = agree(got, want)
    e_got, e_ref = orc.mean_error(Ws, got, H), orc.mean_error(Ws, want, H)
    print(f"[local search {r}x{n} c={c} moves={moves}] agreement {a:.6f} error {e_got:.6e} vs {e_ref:.6e}")
    assert a >= 0.9999 and rel(e_got, e_ref) <= 1e-3
    assert e_got < orc.mean_error(Ws, Q0, H)


def test_gain_fp64_vs_exhaustive(slk):
    # reference tests/test_obq.py:112-140
    rng = np.random.default_rng(8)
    W = rng.standard_normal((10, 16))
    A = rng.standard_normal((16, 10)).astype(np.float32)
    H = (A @ A.T).astype(np.float64)
    Q = np.round(W)
    C = Q + np.square(rng.standard_normal((10, 16)))
    base = orc.rowwise_error(W, Q, H)
    want = np.zeros_like(Q)
    for i in range(16):
        cur = Q.copy()
        cur[:, i] = C[:, i]
        want[:, i] = base - orc.rowwise_error(W, cur, H)
    got = slk.obq.compute_gain(W, Q, H, C)
    assert got.dtype == np.float64
    np.testing.assert_allclose(got, want)
    np.testing.assert_allclose(slk.obq.channelwise_error(W, Q, H), base)


# ---------------------------------------------------------------------------
# obq-aware scaling, statistics, presets
# ---------------------------------------------------------------------------


def test_obq_scaling_vs_golden(slk):
    g = load_golden("obq_scaling")
    cb = slk.codebook.UniformCodebook(8, -1, 1)
    grid = orc.UniformGrid(8, -1, 1)
    # A grid point's error here is the error AFTER a GPTQ sweep at that scale, and GPTQ amplifies the
    # last-ulp difference between two fp64 factors (noise floor 1-3e-4 of the error), so
    # two grid points whose post-sweep errors are closer than BASELINE's 1e-3 tolerance are a tie.
    a = slk.scaling.compute_obq_scaling(g["W"], cb, 0, H=g["H"], grid_size=12, min_factor=0.3)
    _, ev = orc.gptq_scale_evaluator(g["W"], grid, 0, g["H"], 0.01, "diag")
    scales_equivalent(a, g["sc_obq"], ev, "test_obq_scaling_vs_golden", "golden [8,80] diag order", min_same=0.85,
                      tie_tol=1e-3)
    b = slk.scaling.compute_obq_scaling(g["W"], cb, 0, H=g["H"], grid_size=12, min_factor=0.3, act_order="sqerr",
                                        damp=0.03)
    _, ev = orc.gptq_scale_evaluator(g["W"], grid, 0, g["H"], 0.03, "sqerr")
    scales_equivalent(b, g["sc_obq_sqerr"], ev, "test_obq_scaling_vs_golden", "golden [8,80] sqerr order, 3 % damp",
                      min_same=0.85, tie_tol=1e-3)


def test_obq_scaling_vs_oracle_more_rows(slk):
    """compute_obq_scaling (scaling.py:137-190) on enough rows for a meaningful fraction: [96, 256],
    20 grid points, against the oracle's own search."""
    W, H, _ = wl.synthetic_layer(96, 256, 17, samples=512)
    cb, grid = slk.codebook.UniformCodebook(8, -1, 1), orc.UniformGrid(8, -1, 1)
    got = slk.scaling.compute_obq_scaling(W, cb, 0, H=H, grid_size=20, min_factor=0.2)
    want = orc.search_scale_gptq(W, grid, 0, H, lo=0.2, points=20)
    _, ev = orc.gptq_scale_evaluator(W, grid, 0, H, 0.01, "diag")
    scales_equivalent(got, want, ev, "test_obq_scaling_vs_oracle_more_rows", "[96,256] c=8, 20 points", min_same=0.97,
                      tie_tol=1e-3)


def test_scaling_quality_ordering(slk):
    # reference tests/test_scaling.py:130-149
    rng = np.random.default_rng(5)
    data = rng.standard_normal((20, 100)).astype(np.float32)
    cb = slk.codebook.UniformCodebook(9, -3, 3)
    A = rng.standard_normal((100, 10)).astype(np.float32)
    H = (A @ A.T).astype(np.float64)
    H = H + 1e-6 * np.linalg.norm(H, ord=2, axis=1) * np.eye(100)
    S, O = slk.scaling, slk.obq
    sc = {k: S.compute_min_mse_scaling(data, cb, 0, H=h) for k, h in
          (("base", None), ("diag", H.diagonal()), ("hess", H))}
    sc["obq"] = S.compute_obq_scaling(data, cb, 0, H=H)
    err = {k: O.quantization_error(S.quantize_with_scaling(data, sc[k], cb), data, H) for k in ("base", "diag", "hess")}
    err["obq"] = O.quantization_error(S.quantize_with_scaling(data, sc["obq"], cb, H=H), data, H)
    assert err["hess"] <= err["base"] and err["hess"] <= err["diag"] and err["obq"] <= err["hess"]
    for mode in ("norm", "max", "mse", "diag", "hessian", "diag1", "hessian1", "diag1.8", "hessian1.8"):
        assert len(S.compute_scaling(data, cb, H, mode=mode)) == 20  # tests/test_scaling.py:152-165
    with pytest.raises(RuntimeError):
        S.compute_scaling(data, cb, H, mode="nope")


def test_statistics_vs_golden(slk):
    g = load_golden("statistics")
    lin = torch.nn.Linear(48, 20)
    with torch.no_grad():
        lin.weight.copy_(torch.from_numpy(g["W"]))
        lin.bias.copy_(torch.from_numpy(g["b"]))
    st = slk.Sleekit(lin)
    st.add_batch(torch.from_numpy(g["X1"]))
    assert st.count == int(g["count1"])
    np.testing.assert_allclose(st.mean.cpu().numpy(), g["mean1"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(st.hessian.cpu().numpy(), g["hess1"], rtol=1e-5, atol=1e-5)
    st.add_batch(torch.from_numpy(g["X2"]))
    assert st.count == int(g["count2"])
    np.testing.assert_allclose(st.mean.cpu().numpy(), g["mean2"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(st.hessian.cpu().numpy(), g["hess2"], rtol=1e-5, atol=1e-5)
    for name in ("basic", "light", "heavy"):
        with torch.no_grad():
            lin.weight.copy_(torch.from_numpy(g["W"]))
            lin.bias.copy_(torch.from_numpy(g["b"]))
        # isolate the stage: start from the reference's own statistics
        st.hessian.copy_(torch.from_numpy(g["hess2"]))
        st.mean.copy_(torch.from_numpy(g["mean2"]))
        getattr(st, {"basic": "quantize_basic", "light": "quantize_sleekit_light", "heavy": "quantize_sleekit_heavy"}[name])(3)
        a = agree(lin.weight.detach().numpy(), g[f"{name}_W"])
        record_parity("test_statistics_vs_golden", f"preset {name}: weights equal", a, 0.999)
        assert a >= 0.999, (name, a)
        np.testing.assert_allclose(lin.bias.detach().numpy(), g[f"{name}_b"], rtol=2e-2, atol=2e-3)
    # sample counting through conv layers (reference tests/test_statistics.py:7-46)
    conv = torch.nn.Conv2d(3, 4, 3, padding=1)
    sc = slk.Sleekit(conv)
    sc.add_batch(torch.randn(2, 3, 8, 8))
    assert sc.count == 2 * 8 * 8 and sc.hessian.shape == (27, 27)
    c1 = torch.nn.Conv1d(3, 4, 3)
    s1 = slk.Sleekit(c1)
    s1.add_batch(torch.randn(2, 3, 10))
    assert s1.count == 2 * 8 and s1.hessian.shape == (9, 9)


def test_hessian_accumulation_vs_numpy(slk):
    from sleekit_b200 import ops

    rng = np.random.default_rng(12)
    for S, n in ((2048, 768), (300, 130), (17, 5)):
        X = (rng.standard_normal((S, n)) * np.exp(rng.standard_normal(n)) + 0.5).astype(np.float32)
        ref = orc.RunningStats(n)
        ref.add_rows(X[: S // 2])
        ref.add_rows(X[S // 2:])
        H = torch.zeros((n, n), dtype=torch.float32, device="cuda")
        mean = torch.zeros(n, dtype=torch.float32, device="cuda")
        cnt = 0
        for part in (X[: S // 2], X[S // 2:]):
            xs = torch.from_numpy(part).cuda()
            keep = cnt / (cnt + part.shape[0])
            cnt += part.shape[0]
            ops.hessian_accum(xs, H, mean, keep, cnt)
        Hn = H.cpu().numpy()
        assert np.abs(Hn - ref.hessian).max() <= 2e-6 * np.abs(ref.hessian).max()
        np.testing.assert_allclose(mean.cpu().numpy(), ref.mean, rtol=1e-5, atol=1e-6)
        np.testing.assert_array_equal(Hn, Hn.T)  # symmetric to the bit: both halves use the same sums


# ---------------------------------------------------------------------------
# device-resident pass-through and edge cases
# ---------------------------------------------------------------------------


def test_device_tensors_pass_through(slk):
    W, H, m = wl.synthetic_layer(64, 256, 2, samples=512)
    cb = slk.codebook.UniformCodebook(8, -1, 1)
    Wd, Hd = torch.from_numpy(W).cuda(), torch.from_numpy(H).cuda()
    sc = slk.scaling.compute_min_mse_scaling(Wd, cb, 0, H=Hd.diagonal().contiguous())
    assert isinstance(sc, torch.Tensor) and sc.is_cuda
    out = slk.scaling.quantize_with_scaling(Wd, sc, cb, H=Hd)
    assert isinstance(out, torch.Tensor) and out.is_cuda
    host = slk.scaling.quantize_with_scaling(W, sc.cpu().numpy(), cb, H=H)
    np.testing.assert_array_equal(out.cpu().numpy(), host)
    e = slk.obq.quantization_error(Wd, out, Hd)
    assert isinstance(e, torch.Tensor) and rel(e.item(), slk.obq.quantization_error(W, host, H)) < 1e-6


def test_host_plan_equals_per_layer_api(slk):
    """The pinned-host layer-set plan (H2D -> hot path -> D2H as graph branches) returns exactly
    what the per-layer numpy API returns, on every replay and after the inputs change."""
    from sleekit_b200.pipeline import LayerSetQuantizer

    shapes = [(64, 256), (96, 128), (32, 320), (64, 256)]
    cb = slk.codebook.UniformCodebook(8, -1, 1)
    lsq = LayerSetQuantizer(cb, scaling_mode="diag", act_order="diag", damp=0.01, streams=3)
    plan = lsq.host_plan(shapes)
    for rep in range(3):
        layers = [wl.synthetic_layer(r, n, 10 * rep + i, samples=512) for i, (r, n) in enumerate(shapes)]
        for i, (W, H, m) in enumerate(layers):
            plan.W[i][...] = W
            plan.H[i][...] = H
        Q, err = plan.run()
        for i, (W, H, m) in enumerate(layers):
            sc = slk.scaling.compute_min_mse_scaling(W, cb, 0, H=H.diagonal())
            want = slk.scaling.quantize_with_scaling(W, sc, cb, H=H)
            np.testing.assert_array_equal(Q[i], want)
            # the plan takes the layer error from the sweep's residuals (test_layer_error_from_sweep_...)
            assert rel(err[i], slk.obq.quantization_error(W, want, H)) < 1e-4


@pytest.mark.parametrize("batched", [False, True])
def test_host_plan_variants(slk, batched):
    """The plan with batched factorisations equals the per-layer plan bit for bit; outputs="codes" returns the
    codebook indices and row scales of the same result; a non-positive-definite Hessian raises LinAlgError
    (obq.py:49-50) naming the layer."""
    from sleekit_b200.pipeline import LayerSetQuantizer

    shapes = [(64, 256), (96, 256), (32, 320), (64, 256), (48, 320)]
    cb = slk.codebook.UniformCodebook(8, -1, 1)
    grid = orc.UniformGrid(8, -1, 1)
    lsq = LayerSetQuantizer(cb, scaling_mode="diag", act_order="diag", damp=0.01, streams=3)
    layers = [wl.synthetic_layer(r, n, 40 + i, samples=512) for i, (r, n) in enumerate(shapes)]
    ref_plan = lsq.host_plan(shapes)
    ref_plan.batch_k2 = False
    plan = lsq.host_plan(shapes, outputs="codes")
    plan.batch_k2 = batched
    for p in (ref_plan, plan):
        for i, (W, H, m) in enumerate(layers):
            p.W[i][...] = W
            p.H[i][...] = H
    Q, err = ref_plan.run()
    (codes, scales), err2 = plan.run()
    np.testing.assert_array_equal(err, err2)
    for i, (W, H, m) in enumerate(layers):
        assert codes[i].dtype == np.uint8 and scales[i].shape == (shapes[i][0],)
        np.testing.assert_array_equal(codes[i], grid.index(orc.divide_rows(Q[i], scales[i], 0)))
        np.testing.assert_array_equal(orc.divide_rows(grid._to_value(codes[i].astype(np.float32)), 1 / scales[i], 0), Q[i])
    plan.H[2][...] = -np.eye(320, dtype=np.float32)
    with pytest.raises(np.linalg.LinAlgError, match="layers \\[2\\]"):
        plan.run()
    plan.H[2][...] = layers[2][1]
    (codes, scales), err3 = plan.run()
    np.testing.assert_array_equal(err3, err)


@pytest.mark.parametrize("n", [32, 100, 768, 1000, 1100, 3072])
def test_symmetric_upload_equals_full_copy(slk, n):
    """slk_upload_symmetric_f32: only the block upper triangle of a symmetric pinned host matrix crosses
    PCIe, the device mirrors the rest -- the device matrix must equal the host matrix bit for bit."""
    from sleekit_b200 import ops

    g = np.random.default_rng(n)
    a = g.standard_normal((n, n)).astype(np.float32)
    h = torch.from_numpy(a + a.T).pin_memory()
    d = torch.full((n, n), float("nan"), device="cuda")
    sent = ops.upload_symmetric(h, d)
    torch.cuda.synchronize()
    assert torch.equal(d.cpu(), h)
    assert sent < 4 * n * n or n <= ops.symmetric_block_rows(n)
    print(f"n={n}: {sent / (4 * n * n):.3f} of the matrix sent")


def test_host_plan_symmetric_upload_equals_full_upload(slk):
    from sleekit_b200.pipeline import LayerSetQuantizer

    shapes = [(64, 1056), (96, 128), (32, 2080)]
    cb = slk.codebook.UniformCodebook(8, -1, 1)
    lsq = LayerSetQuantizer(cb, scaling_mode="diag", act_order="diag", damp=0.01, streams=3)
    outs = []
    for sym in (True, False):
        plan = lsq.host_plan(shapes, symmetric_h=sym)
        for i, (r, n) in enumerate(shapes):
            W, H, m = wl.synthetic_layer(r, n, 40 + i, samples=512)
            plan.W[i][...] = W
            plan.H[i][...] = np.triu(H) + np.triu(H, 1).T          # exactly symmetric input
        Q, err = plan.run()
        outs.append(([q.copy() for q in Q], err.copy(), plan.h2d_bytes))
    for a, b in zip(outs[0][0], outs[1][0]):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(outs[0][1], outs[1][1])
    assert outs[0][2] < outs[1][2]


def test_edge_cases(slk):
    cb = slk.codebook.UniformCodebook(4, -1, 1)
    W = np.zeros((3, 40), np.float32)
    W[1, :] = np.linspace(-1, 1, 40)
    sc = slk.scaling.compute_min_mse_scaling(W, cb)
    assert sc[0] == np.float32(1e-16) * np.float32(0.05) and sc[1] > 0  # all-zero rows hit the floor
    np.testing.assert_array_equal(sc, orc.search_scale(W, orc.UniformGrid(4, -1, 1)))
    # dead input column: remove_dead_values then GPTQ still works
    Wl, H, m = wl.synthetic_layer(8, 48, 1, samples=128)
    H[:, 7] = 0
    H[7, :] = 0
    H2, W2 = H.copy(), Wl.copy()
    slk.obq.remove_dead_values(H2, W2)
    Hr, Wr = H.copy(), Wl.copy()
    orc.patch_dead_inputs(Hr, Wr)
    np.testing.assert_array_equal(H2, Hr)
    np.testing.assert_array_equal(W2, Wr)
    q = slk.obq.quantize_opt(W2, H2, slk.codebook.UniformCodebook(8, -1, 1))
    assert np.isfinite(q).all()
    # single row, single leaf narrower than a warp
    q = slk.obq.quantize_opt(W2[:1, :5].copy(), np.ascontiguousarray(H2[:5, :5]), slk.codebook.UniformCodebook(8, -1, 1))
    want = orc.gptq(W2[:1, :5].copy(), np.ascontiguousarray(H2[:5, :5]), orc.UniformGrid(8, -1, 1))
    np.testing.assert_array_equal(q, want)
    with pytest.raises(TypeError):
        slk.obq.quantize_opt(W2, H2, lambda x: np.round(x))


def test_fast_exact_division_matches_ieee_exhaustively(slk):
    """K5 / K3 divide by per-row and per-codebook constants through a reciprocal + two FMA
    corrections; it must equal the IEEE quotient for every dividend (2^32 of them) per divisor."""
    import ctypes

    from sleekit_b200 import _lib

    rng = np.random.default_rng(99)
    divs = [2.0 / (c - 1) for c in (2, 3, 4, 8, 16, 9, 256)] + [4.0 / 8, 6.0 / 8, 1.0, 3.0, 0.1, 5e-18, 7.0, 1e-3]
    divs += list(np.exp(rng.uniform(-12, 12, 20)))            # scales
    divs += list(1.0 / np.exp(rng.uniform(-12, 12, 12)))      # reciprocals of scales
    divs += [float(np.nextafter(np.float32(2.0), np.float32(0.0)))]  # all-ones significand: falls back
    d = torch.tensor(np.array(divs, dtype=np.float32), device="cuda")
    bad = torch.zeros(d.numel(), dtype=torch.int64, device="cuda")
    _lib.call("slk_selftest_fastdiv_f32", ctypes.c_void_p(d.data_ptr()), d.numel(), ctypes.c_void_p(bad.data_ptr()),
              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert int(bad.sum().item()) == 0, [(divs[i], int(b)) for i, b in enumerate(bad.tolist()) if b]


# ---------------------------------------------------------------------------
# the small public functions (a7, a11, a18, a22, a2/a3 for conv layers) and the upload cache
# ---------------------------------------------------------------------------


def test_compute_mse_all_branches_vs_oracle(slk):
    """_compute_mse (scaling.py:84-95): H None, 1-D (fp32 and fp64) and 2-D, against the oracle."""
    rng = np.random.default_rng(12)
    E = (rng.standard_normal((37, 515)) * 0.1).astype(np.float32)
    _, H, _ = wl.synthetic_layer(4, 515, 2, samples=256)
    f = slk.scaling._compute_mse
    got = f(None, E)
    assert got.dtype == np.float32 and got.shape == (37,)
    np.testing.assert_allclose(got, orc.weighted_sq_error(None, E), rtol=2e-6)
    hd = H.diagonal().copy()
    got = f(hd, E)
    assert got.dtype == np.float32
    np.testing.assert_allclose(got, orc.weighted_sq_error(hd, E), rtol=2e-6)
    h64 = rng.random(515)
    got = f(h64, E)
    assert got.dtype == np.float64
    np.testing.assert_allclose(got, orc.weighted_sq_error(h64, E), rtol=1e-12)
    got = f(H, E)
    assert got.dtype == np.float32
    np.testing.assert_allclose(got, orc.weighted_sq_error(H, E), rtol=2e-5)
    # device tensors pass through
    gd = f(torch.from_numpy(hd).cuda(), torch.from_numpy(E).cuda())
    assert gd.is_cuda
    np.testing.assert_allclose(gd.cpu().numpy(), orc.weighted_sq_error(hd, E), rtol=2e-6)


def test_apply_scaling_in_place(slk):
    """apply_scaling_in_place (scaling.py:28-32) writes data / scale into the caller's array."""
    rng = np.random.default_rng(4)
    for axis, shape in ((0, (13, 70)), (1, (13, 70)), (1, (5, 6, 7))):
        data = rng.standard_normal(shape).astype(np.float32)
        scale = (rng.random(shape[axis]) + 0.5).astype(np.float32)
        want = orc.divide_rows(data, scale, axis)
        keep = data
        slk.scaling.apply_scaling_in_place(data, scale, axis)
        assert data is keep
        np.testing.assert_array_equal(data, want)


def test_conv_layer_hessian_values(slk):
    """Sleekit.add_batch on Conv2d / Conv1d layers (statistics.py:44-87): the Hessian and the mean equal
    numpy's on the unfolded input (the unfold itself stays torch's, as in the reference)."""
    import torch.nn.functional as F

    torch.manual_seed(0)
    conv = torch.nn.Conv2d(3, 4, 3, padding=1, stride=2)
    st = slk.Sleekit(conv)
    x1, x2 = torch.randn(2, 3, 9, 8), torch.randn(3, 3, 9, 8)
    st.add_batch(x1)
    st.add_batch(x2)
    cols = [F.unfold(x, conv.kernel_size, conv.dilation, conv.padding, conv.stride).permute(1, 0, 2).flatten(1)
            for x in (x1, x2)]
    X = torch.cat(cols, dim=1).numpy().astype(np.float64)          # [features, samples]
    assert st.count == X.shape[1]
    np.testing.assert_allclose(st.hessian.cpu().numpy(), X @ X.T / X.shape[1], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(st.mean.cpu().numpy(), X.mean(axis=1), rtol=1e-5, atol=1e-6)
    c1 = torch.nn.Conv1d(3, 4, 3, dilation=2)
    s1 = slk.Sleekit(c1)
    y = torch.randn(4, 3, 20)
    s1.add_batch(y)
    u = F.unfold(y.unsqueeze(-1), (3, 1), (2, 1), (0, 0), (1, 1)).permute(1, 0, 2).flatten(1).numpy().astype(np.float64)
    assert s1.count == u.shape[1]
    np.testing.assert_allclose(s1.hessian.cpu().numpy(), u @ u.T / u.shape[1], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("block_size", [3, 4, 7, 8, 63, 64])
@pytest.mark.parametrize("nb_blocks", [2, 8])
def test_blocked_sweep_equals_unblocked_for_reference_leaf_widths(slk, block_size, nb_blocks):
    """The reference's tests/test_obq.py:57-70 (test_blockobq) on the device: _quantize_opt_block with
    min_block_size in {3,4,7,8,63,64} equals _quantize_opt_core."""
    rng = np.random.default_rng(7)
    size = 65
    A = rng.standard_normal((size, 2 * size)).astype(np.float32)
    H = (A @ A.T).astype(np.float64) + 0.5 * np.eye(size)
    W = rng.standard_normal((10, size)).astype(np.float32)
    cb = slk.codebook.UniformCodebook(9, -4, 4)       # step 1: rounds like the reference test's np.round
    Hinv = slk.obq.compute_hessian_chol(H)
    Q1, E1 = W.copy(), np.zeros_like(W)
    Q2, E2 = W.copy(), np.zeros_like(W)
    slk.obq._quantize_opt_core(Q1, E1, Hinv, cb)
    slk.obq._quantize_opt_block(Q2, E2, Hinv, cb, block_size, nb_blocks)
    np.testing.assert_allclose(Q1, Q2, atol=1e-5)
    np.testing.assert_allclose(E1, E2, rtol=1e-4, atol=1e-5)
    Qo, Eo = W.copy(), np.zeros_like(W)
    orc.sweep_in_place(Qo, Eo, orc.inverse_upper_factor(H), orc.UniformGrid(9, -4, 4), min(block_size, 32), nb_blocks)
    np.testing.assert_allclose(Q2, Qo, atol=1e-5)


def test_local_search_quantizer_moves_one_by_one(slk):
    """LocalSearchQuantizer.do_move (obq.py:338-346) keeps (Q - W) H between calls
    (slk_local_search_step_f32): k single moves equal quantize_local_search(k), and the attributes the
    reference exposes follow the current Q."""
    g = load_golden("local_search")
    cb = slk.codebook.UniformCodebook(4, -1, 1)
    Ws, H, Q0 = g["Ws"], g["H"], g["Q0"]
    ls = slk.obq.LocalSearchQuantizer(Ws, Q0, H, cb)
    for k in range(1, 31):
        ls.do_move()
        if k in (1, 5, 30):
            np.testing.assert_array_equal(ls.Q, g[f"ls_{k}"], err_msg=f"{k} single moves")
    grid = orc.UniformGrid(4, -1, 1)
    np.testing.assert_array_equal(ls.Q_up, grid.up(g["ls_30"]))
    np.testing.assert_array_equal(ls.Q_down, grid.down(g["ls_30"]))
    np.testing.assert_allclose(ls.gain_up, orc.flip_gain(Ws, g["ls_30"], H, grid.up(g["ls_30"])), rtol=1e-4, atol=1e-6)
    assert ls.nchannels == Ws.shape[0]


def test_upload_cache_reuses_and_invalidates(slk):
    """_convert: a large read-only host operand is uploaded once and found again by identity; sleekit's own
    in-place functions and invalidate() drop the stale copy; a changed array is detected by its fingerprint."""
    from sleekit_b200 import _convert as cv

    cv.cache_clear()
    W, H, m = wl.synthetic_layer(96, 768, 8, samples=512)
    cb = slk.codebook.UniformCodebook(8, -1, 1)
    h0, miss0 = cv.CACHE_HITS, cv.CACHE_MISSES
    sc = slk.scaling.compute_min_mse_scaling(W, cb, 0, H=H.diagonal().copy())
    q = slk.scaling.quantize_with_scaling(W, sc, cb, H=H)
    e1 = slk.obq.quantization_error(W, q, H)
    assert cv.CACHE_MISSES - miss0 == 2                      # W and H, once each
    assert cv.CACHE_HITS - h0 >= 3                           # W twice more, H once more, q from its own download
    before = cv.H2D_BYTES
    e2 = slk.obq.quantization_error(W, q, H)
    assert cv.H2D_BYTES == before and e1 == e2               # nothing crosses PCIe the second time
    # in-place mutation through sleekit's own function: the copy is dropped and the new content is used
    H2, W2 = H.copy(), W.copy()
    H2[5, :] = 0
    H2[:, 5] = 0
    slk.obq.quantization_error(W2, q, H2)                    # caches H2 (with its dead column)
    slk.obq.remove_dead_values(H2, W2)
    assert H2[5, 5] != 0 and np.all(W2[:, 5] == 0)
    e3 = slk.obq.quantization_error(W2, q, H2)
    np.testing.assert_allclose(e3, orc.mean_error(W2, q, H2), rtol=1e-4)
    # a diagonal overwrite (dampening in place) changes the fingerprint
    H2[np.arange(768), np.arange(768)] += np.float32(1.0)
    e4 = slk.obq.quantization_error(W2, q, H2)
    np.testing.assert_allclose(e4, orc.mean_error(W2, q, H2), rtol=1e-4)
    cv.cache_clear()


def test_layer_on_a_second_device(slk):
    """kernel attributes and SM counts are per device, and ops launch on the device (and stream)
    of their tensors -- a layer that lives on cuda:1 while cuda:0 is the current device gives the same bits."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r, n = 768, 1024                                       # macro sweep (> 48 KB of shared memory), batched-size factor
    W, H, _ = wl.synthetic_layer(r, n, 33, samples=1024)
    cb = slk.codebook.UniformCodebook(8, -1, 1)
    outs = []
    torch.cuda.set_device(0)
    for dev in ("cuda:0", "cuda:1"):
        Wd, Hd = torch.from_numpy(W).to(dev), torch.from_numpy(H).to(dev)
        sc = slk.scaling.compute_min_mse_scaling(Wd, cb, 0, H=Hd.diagonal().contiguous())
        q = slk.scaling.quantize_with_scaling(Wd, sc, cb, H=Hd, nb_ls_moves=3)
        e = slk.obq.quantization_error(Wd, q, Hd)
        assert q.device == Wd.device and sc.device == Wd.device
        outs.append((sc.cpu(), q.cpu(), float(e)))
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and outs[0][2] == outs[1][2]
