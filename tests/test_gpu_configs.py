"""GPU parity at the shapes of BASELINE.json's configs 3-5: the CUDA path at
full layer size, the oracle on a row subset of the same layer (rows never interact once H, the
ordering and the factor are fixed: obq.py:106-137, scaling.py:127-133), plus the size-independent
properties of the domain (GPTQ beats plain rounding, local search never increases the error,
values lie on the codebook, a row slice quantized alone equals the same rows of the full run)."""

import numpy as np
import pytest
import torch

from oracle import sleekit_oracle as orc
from sleekit_b200 import workloads as wl
from tests.conftest import record_parity, scales_equivalent

pytestmark = pytest.mark.gpu


def agree(a, b):
    return float((np.asarray(a) == np.asarray(b)).mean())


def rel(a, b):
    return abs(float(a) - float(b)) / abs(float(b))


@pytest.fixture(scope="module")
def api():
    from sleekit_b200 import codebook, obq, ops, scaling

    class NS:
        pass

    ns = NS()
    ns.codebook, ns.obq, ns.ops, ns.scaling = codebook, obq, ops, scaling
    return ns


def device_layer(r, n, lid, samples):
    """W on the host, H = X^T X / S and the mean built on the device by K1 (fp32-faithful)."""
    from sleekit_b200 import ops

    W = wl.synthetic_weight(r, n, lid)
    X = torch.from_numpy(wl.synthetic_calibration(n, lid, samples)).cuda()
    H = torch.zeros((n, n), dtype=torch.float32, device="cuda")
    m = torch.zeros(n, dtype=torch.float32, device="cuda")
    ops.hessian_accum(X, H, m, 0.0, samples)
    return W, H, m


def test_config3_full_h_scaling_bias_corrected_1p5_bit(api):
    """OPT-350M / BLOOM-560M-shaped [1024, 4096] layer, 3-entry codebook, H - m m^T, full-H scale
    search, GPTQ: first rows against the oracle given the same H."""
    r, n, rows = 1024, 4096, 24
    W, Hd, md = device_layer(r, n, 31, 4096)
    cb, grid = api.codebook.UniformCodebook(3, -1, 1), orc.UniformGrid(3, -1, 1)
    Hc = api.obq.remove_input_bias(Hd, md)
    Wd = torch.from_numpy(W).cuda()
    sc = api.scaling.compute_min_mse_scaling(Wd, cb, 0, H=Hc)
    q = api.scaling.quantize_with_scaling(Wd, sc, cb, H=Hc)
    err = float(api.obq.quantization_error(Wd, q, Hc))
    # oracle on the first rows, same H (host copy)
    Hh = Hc.cpu().numpy()
    np.testing.assert_array_equal(Hh, orc.strip_input_bias(Hd.cpu().numpy(), md.cpu().numpy()))
    sc_ref = orc.search_scale(W[:rows], grid, 0, H=Hh)
    sc_got = sc[:rows].cpu().numpy()
    Wr = W[:rows]
    scales_equivalent(sc_got, sc_ref, lambda s_: orc.weighted_sq_error(Hh, orc.quantize_scaled(Wr, s_, grid) - Wr),
                      "test_config3", "[1024,4096] c=3 full-H search, first 24 rows")
    q_ref = orc.quantize_scaled(W[:rows], sc_got, grid, H=Hh, rule="diag", damp=0.01)
    q_got = q[:rows].cpu().numpy()
    a = agree(grid.index(orc.divide_rows(q_got, sc_got, 0)), grid.index(orc.divide_rows(q_ref, sc_got, 0)))
    eg, er = orc.mean_error(W[:rows], q_got, Hh), orc.mean_error(W[:rows], q_ref, Hh)
    print(f"config 3: code agreement {a:.6f}, layer error {eg:.6e} vs {er:.6e}; full layer error {err:.6e}")
    record_parity("test_config3", "GPTQ code agreement (24 rows)", a, 0.999)
    record_parity("test_config3", "GPTQ layer error rel diff", rel(eg, er), 1e-3)
    assert a >= 0.999 and rel(eg, er) <= 1e-3
    rtn = api.scaling.quantize_with_scaling(Wd, sc, cb)
    assert err < float(api.obq.quantization_error(Wd, rtn, Hc))


def test_config4_llama7b_mlp_gptq_local_search(api):
    """Llama-2-7B gate/up-shaped [11008, 4096] layer, 4-entry codebook, GPTQ + 10 best-first moves."""
    r, n, rows, moves = 11008, 4096, 16, 10
    W, Hd, md = device_layer(r, n, 41, 4096)
    cb, grid = api.codebook.UniformCodebook(4, -1, 1), orc.UniformGrid(4, -1, 1)
    Wd = torch.from_numpy(W).cuda()
    sc = api.scaling.compute_min_mse_scaling(Wd, cb, 0, H=Hd.diagonal().contiguous())
    q0 = api.scaling.quantize_with_scaling(Wd, sc, cb, H=Hd)
    q1 = api.scaling.quantize_with_scaling(Wd, sc, cb, H=Hd, nb_ls_moves=moves)
    e0 = api.obq.channelwise_error(Wd, q0, Hd).cpu().numpy()
    e1 = api.obq.channelwise_error(Wd, q1, Hd).cpu().numpy()
    assert np.all(e1 <= e0 * (1 + 1e-5) + 1e-12), "local search increased a row error"
    assert e1.mean() < e0.mean()
    codes = api.codebook.UniformCodebook(4, -1, 1).quantize_value(api.scaling.apply_scaling(q1, sc, 0))
    np.testing.assert_allclose(codes.cpu().numpy(), api.scaling.apply_scaling(q1, sc, 0).cpu().numpy(), atol=2e-6)
    # a row slice alone gives the same rows (rows never interact)
    part = api.scaling.quantize_with_scaling(Wd[:rows].contiguous(), sc[:rows].contiguous(), cb, H=Hd, nb_ls_moves=moves)
    assert torch.equal(part, q1[:rows])
    # oracle on those rows
    Hh, sch = Hd.cpu().numpy(), sc[:rows].cpu().numpy()
    np.testing.assert_array_equal(sch, orc.search_scale(W[:rows], grid, 0, H=Hh.diagonal()))
    ref0 = orc.quantize_scaled(W[:rows], sch, grid, H=Hh, rule="diag", damp=0.01)
    a0 = agree(grid.index(orc.divide_rows(q0[:rows].cpu().numpy(), sch, 0)), grid.index(orc.divide_rows(ref0, sch, 0)))
    eg, er = orc.mean_error(W[:rows], q0[:rows].cpu().numpy(), Hh), orc.mean_error(W[:rows], ref0, Hh)
    print(f"config 4: GPTQ code agreement {a0:.6f}, error {eg:.6e} vs {er:.6e}")
    record_parity("test_config4", "[11008,4096] c=4 GPTQ code agreement (16 rows)", a0, 0.999)
    record_parity("test_config4", "GPTQ layer error rel diff", rel(eg, er), 1e-3)
    assert a0 >= 0.999 and rel(eg, er) <= 1e-3
    # local search from identical (W, Q, H): identical moves
    Ws = orc.divide_rows(W[:rows], sch, 0)
    Qs = grid(orc.divide_rows(q0[:rows].cpu().numpy(), sch, 0))      # exactly on the codebook
    ls_ref = orc.local_search(Ws, Qs, Hh, grid, moves)
    ls_got = api.obq.quantize_local_search(Ws, Qs, Hh, cb, moves)
    a1 = agree(ls_got, ls_ref)
    print(f"config 4: local search ({moves} moves) agreement {a1:.6f}")
    record_parity("test_config4", "local search 10 moves from identical (W,Q,H): weights equal", a1, 0.9999)
    assert a1 >= 0.9999


def test_config5_llama70b_down_proj_rank_share(api):
    """Llama-3-70B down-proj shape, n = 28672, with the 1024 rows one of 8 GPUs owns: the fp64
    factor at full size (R R^T = H_opt on sampled rows), GPTQ beats rounding, values on the
    codebook, a row slice alone equals the same rows of the run."""
    from sleekit_b200 import ops

    r, n = 1024, 28672
    W, Hd, md = device_layer(r, n, 51, 2048)
    cb = api.codebook.UniformCodebook(8, -1, 1)
    Wd = torch.from_numpy(W).cuda()
    dampval = ops.damp_value(Hd, 0.01)
    order = ops.argsort(ops.order_keys(Hd, dampval, None))
    r32, rt, ud32, info = ops.chol_factor(Hd, order, dampval)
    assert int(info.item()) == 0
    pick = torch.tensor([0, 1, 31, 32, 4095, 14336, 28000, 28671], device="cuda")
    Rrows = r32[pick].double()
    got = Rrows @ r32.double().T                                     # rows of R R^T
    want = Hd[order][:, order][pick].double()
    want[torch.arange(len(pick)), pick] += dampval.double()
    scale = float(Hd.diagonal().max())
    assert float((got - want).abs().max()) <= 2e-5 * scale
    del r32, rt, ud32, got, want, Rrows
    sc = api.scaling.compute_min_mse_scaling(Wd, cb, 0, H=Hd.diagonal().contiguous())
    q = api.scaling.quantize_with_scaling(Wd, sc, cb, H=Hd)
    rtn = api.scaling.quantize_with_scaling(Wd, sc, cb)
    eq, er = float(api.obq.quantization_error(Wd, q, Hd)), float(api.obq.quantization_error(Wd, rtn, Hd))
    print(f"config 5 (1024-row share): GPTQ layer error {eq:.6e}, plain rounding {er:.6e}")
    assert np.isfinite(eq) and eq < er
    scaled = api.scaling.apply_scaling(q, sc, 0)
    assert float((cb.quantize_value(scaled) - scaled).abs().max()) <= 2e-6
    part = api.scaling.quantize_with_scaling(Wd[:32].contiguous(), sc[:32].contiguous(), cb, H=Hd)
    assert torch.equal(part, q[:32])


# ---------------------------------------------------------------------------
# n > 4096: the two-level lazy batching of the sweep (2048-column super blocks + one K = 2048 push per
# super block, sweep.cu slk_gptq_sweep_r_err_f32) against the oracle's recursion (obq.py:121-137)
# ---------------------------------------------------------------------------


def _wide_layer_vs_oracle(api, r, n, c, lid, test, moves=0):
    W, Hd, md = device_layer(r, n, lid, 2048)           # S = 2048 < n: rank-deficient H, PD through damping
    cb, grid = api.codebook.UniformCodebook(c, -1, 1), orc.UniformGrid(c, -1, 1)
    Wd = torch.from_numpy(W).cuda()
    Hh = Hd.cpu().numpy()
    sc = api.scaling.compute_min_mse_scaling(Wd, cb, 0, H=Hd.diagonal().contiguous())
    sch = orc.search_scale(W, grid, 0, H=Hh.diagonal())
    hdiag = Hh.diagonal().copy()
    scales_equivalent(sc.cpu().numpy(), sch, lambda s_: orc.weighted_sq_error(hdiag, orc.quantize_scaled(W, s_, grid) - W),
                      test, f"[{r},{n}] c={c} diag-H search")
    sc = torch.from_numpy(sch).cuda()                    # identical scales on both sides from here on
    q = api.scaling.quantize_with_scaling(Wd, sc, cb, H=Hd)
    err_dev = float(api.obq.quantization_error(Wd, q, Hd))
    got = q.cpu().numpy()
    ref = orc.quantize_scaled(W, sch, grid, H=Hh, rule="diag", damp=0.01)   # fp64 LAPACK factor + reference recursion
    a = agree(grid.index(orc.divide_rows(got, sch, 0)), grid.index(orc.divide_rows(ref, sch, 0)))
    eg, er = orc.mean_error(W, got, Hh), orc.mean_error(W, ref, Hh)
    print(f"{test}: [{r},{n}] c={c}: code agreement {a:.6f}, layer error {eg:.6e} vs oracle {er:.6e}")
    record_parity(test, f"[{r},{n}] c={c} two-level sweep: code agreement", a, 0.999)
    record_parity(test, f"[{r},{n}] c={c} two-level sweep: layer error rel diff", rel(eg, er), 1e-3)
    assert a >= 0.999 and rel(eg, er) <= 1e-3
    assert rel(err_dev, er) <= 1e-3                      # K6 on the device agrees with the oracle's error too
    # the fused path's own error (from the sweep residuals) as well
    from sleekit_b200.scaling import quantize_scaled_device

    q2, (e2, _) = quantize_scaled_device(Wd, sc, cb, Hd, "diag", 0.01, 0, want_err=True)
    assert torch.equal(q2, q)
    record_parity(test, f"[{r},{n}] layer error from sweep residuals rel diff vs oracle", rel(float(e2), er), 1e-3)
    assert rel(float(e2), er) <= 1e-3
    if moves:
        Ws = orc.divide_rows(W, sch, 0)
        Qs = grid(orc.divide_rows(got, sch, 0))
        ls_ref = orc.local_search(Ws, Qs, Hh, grid, moves)
        ls_got = api.obq.quantize_local_search(Ws, Qs, Hh, cb, moves)
        a1 = agree(ls_got, ls_ref)
        record_parity(test, f"[{r},{n}] local search {moves} moves: weights equal", a1, 0.9999)
        assert a1 >= 0.9999


@pytest.mark.parametrize("r,n,c", [(16, 4352, 8), (16, 6144, 8), (24, 6144, 4)])
def test_two_level_sweep_vs_oracle(api, r, n, c):
    """[16,4352]: super blocks 2048 + 2048 + 256 (a ragged last one); [.,6144]: three full super blocks."""
    _wide_layer_vs_oracle(api, r, n, c, 61, "test_two_level_sweep_vs_oracle")


def test_config4_down_proj_vs_oracle(api):
    """Llama-2-7B down-proj shape: 32 of its 4096 rows over the full n = 11008 (six super blocks, the last
    one ragged: 11008 = 5 * 2048 + 768), 4-entry codebook, + 10 local-search moves; the oracle runs the
    reference's whole path once (one fp64 LAPACK factor of an 11008 x 11008 matrix, ~1 min of CPU)."""
    _wide_layer_vs_oracle(api, 32, 11008, 4, 43, "test_config4_down_proj_vs_oracle", moves=10)
