"""Generate golden input/output vectors by running the UNMODIFIED reference.

Run in the build container only (the reference is not present on the GPU box):

    PYTHONPATH=/root/reference python tests/golden/make_golden.py

It imports ``sleekit`` from /root/reference, feeds it small seeded inputs and
stores inputs and outputs in ``tests/golden/*.npz``.  Nothing here is product
code; the fixtures pin ``oracle/`` (CPU tests) and the CUDA path (GPU tests).
"""

import os
import sys

import numpy as np

REF = os.environ.get("SLEEKIT_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
# make sure we are not picking up this repo's drop-in shim
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))]

import sleekit  # noqa: E402

assert os.path.abspath(sleekit.__file__).startswith(os.path.abspath(REF)), sleekit.__file__

import torch  # noqa: E402
from sleekit.codebook import Codebook, UniformCodebook  # noqa: E402
from sleekit import obq as robq  # noqa: E402
from sleekit import scaling as rsc  # noqa: E402
from sleekit.statistics import Sleekit  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def calib(rng, S, n):
    base = rng.standard_normal((S, 16)).astype(np.float32) @ rng.standard_normal((16, n)).astype(np.float32)
    X = base + np.float32(0.3) * rng.standard_normal((S, n)).astype(np.float32)
    X = X * np.exp(0.5 * rng.standard_normal(n)).astype(np.float32) + np.float32(0.5)
    return np.ascontiguousarray(X, dtype=np.float32)


def layer(seed, r, n, S=256):
    rng = np.random.default_rng(seed)
    W = (0.02 * rng.standard_normal((r, n))).astype(np.float32)
    X = calib(rng, S, n)
    H = (X.T @ X / np.float32(S)).astype(np.float32)
    m = X.mean(axis=0).astype(np.float32)
    return W, H, m, X


def rounding():
    rng = np.random.default_rng(1)
    out = {}
    for c, lo, hi in [(2, -1, 1), (3, -1, 1), (4, -1, 1), (8, -1, 1), (16, -1, 1), (9, -2, 2), (9, -3.0, 3.0)]:
        cb = UniformCodebook(c, lo, hi)
        x = (rng.standard_normal((6, 40)) * 0.8 * hi).astype(np.float32)
        # exact midpoints and codewords as fp32 see them, plus their neighbours
        step = np.float32(cb.scale)
        mids = (np.arange(c - 1, dtype=np.float32) + np.float32(0.5)) * step + np.float32(lo)
        pts = np.concatenate([mids, np.nextafter(mids, np.float32(9)), np.nextafter(mids, np.float32(-9)),
                              cb.values.astype(np.float32), np.float32([lo - 1, hi + 1, 0.0, -0.0])])
        x = np.concatenate([x.ravel(), pts]).astype(np.float32)
        tag = f"u{c}_{int(hi)}"
        out[tag + "_cfg"] = np.array([c, lo, hi], dtype=np.float64)
        out[tag + "_x"] = x
        out[tag + "_idx"] = cb.quantize_index(x)
        out[tag + "_val"] = cb.quantize_value(x)
        out[tag + "_up"] = cb.quantize_up(x)
        out[tag + "_down"] = cb.quantize_down(x)
        xd = x.astype(np.float64)
        out[tag + "_idx64"] = cb.quantize_index(xd)
        out[tag + "_val64"] = cb.quantize_value(xd)
        out[tag + "_up64"] = cb.quantize_up(xd)
        out[tag + "_down64"] = cb.quantize_down(xd)
    tb = Codebook.nf4()
    x = rng.standard_normal(300).astype(np.float32) * np.float32(0.6)
    x = np.concatenate([x, tb.thresholds, tb.values, np.float32([-3, 3])]).astype(np.float32)
    out["nf4_values"] = tb.values
    out["nf4_limits"] = tb.thresholds
    out["nf4_x"] = x
    out["nf4_idx"] = tb.quantize_index(x)
    out["nf4_val"] = tb.quantize_value(x)
    out["nf4_up"] = tb.quantize_up(x)
    out["nf4_down"] = tb.quantize_down(x)
    np.savez_compressed(os.path.join(OUT, "rounding.npz"), **out)


def scales():
    out = {}
    W, H, m, X = layer(11, 24, 96)
    W[5] = 0.0  # an all-zero row hits the 1e-16 floor (ref: scaling.py:54)
    out["W"], out["H"], out["mean"] = W, H, m
    for c in (3, 8):
        cb = UniformCodebook(c, -1, 1)
        out[f"max_c{c}"] = rsc.compute_non_saturating_scaling(W, cb, 0)
        out[f"mse_c{c}"] = rsc.compute_min_mse_scaling(W, cb, 0)
        out[f"diag_c{c}"] = rsc.compute_min_mse_scaling(W, cb, 0, H=H.diagonal())
        out[f"full_c{c}"] = rsc.compute_min_mse_scaling(W, cb, 0, H=H)
        out[f"diag5_c{c}"] = rsc.compute_scaling(W, cb, H, mode="diag5")
        out[f"hess2_c{c}"] = rsc.compute_scaling(W, cb, H, mode="hessian2")
        out[f"axis1_c{c}"] = rsc.compute_min_mse_scaling(W, cb, 1, grid_size=17, min_factor=0.2)
    out["norm0"] = rsc.compute_norm_scaling(W, 0)
    out["norm1"] = rsc.compute_norm_scaling(W, 1)
    tb = Codebook([-1.0, 0.0, 10.0, 20.0])
    out["max_tab0"] = rsc.compute_non_saturating_scaling(W, tb, 0)
    cb = UniformCodebook(8, -1, 1)
    sc = out["diag_c8"]
    out["qws_plain"] = rsc.quantize_with_scaling(W, sc, cb)
    out["apply"] = rsc.apply_scaling(W, sc, 0)
    np.savez_compressed(os.path.join(OUT, "scales.npz"), **out)


def factor_and_sweep():
    out = {}
    W, H, m, X = layer(21, 24, 96)
    out["W"], out["H"], out["mean"] = W, H, m
    Hd = H + 0.01 * H.diagonal().mean() * np.eye(96)
    out["Hd"] = Hd
    out["U"] = robq.compute_hessian_chol(Hd)
    out["Hc"] = robq.remove_input_bias(H, m)
    cb = UniformCodebook(8, -1, 1)
    sc = rsc.compute_min_mse_scaling(W, cb, 0, H=H.diagonal())
    Ws = rsc.apply_scaling(W, sc, 0)
    out["scale"], out["Ws"] = sc, Ws
    for rule in ("diag", "none", "err", "sqerr"):
        out[f"order_{rule}"] = robq.compute_hessian_order(Ws, Hd, cb, rule)
        out[f"gptq_{rule}"] = robq.quantize_opt(Ws, H, cb, act_order=rule, damp=0.01)
    out["gptq_damp3"] = robq.quantize_opt(Ws, H, cb, act_order="sqerr", damp=0.03)
    out["gptq_ls20"] = robq.quantize_opt(Ws, H, cb, act_order="diag", damp=0.01, nb_ls_moves=20)
    out["qws_gptq"] = rsc.quantize_with_scaling(W, sc, cb, H=H, act_order="diag", damp=0.01)
    out["err_rows"] = robq.channelwise_error(W, out["qws_gptq"], H)
    out["err_mean"] = np.array(robq.quantization_error(W, out["qws_gptq"], H))
    # raw sweep with a given factor: pins the leaf + trailing-update arithmetic
    Q = Ws.copy()
    E = np.zeros_like(Ws)
    U = robq.compute_hessian_chol(Hd)
    robq._quantize_opt_block(Q, E, U, cb, 32, 8)
    out["sweep_Q"], out["sweep_E"] = Q, E
    # a wider layer with a ragged recursion (200 -> 32-wide blocks, last one 8)
    W2, H2, m2, _ = layer(22, 16, 200)
    cb4 = UniformCodebook(4, -1, 1)
    sc2 = rsc.compute_min_mse_scaling(W2, cb4, 0)
    out["W2"], out["H2"], out["scale2"] = W2, H2, sc2
    out["qws2"] = rsc.quantize_with_scaling(W2, sc2, cb4, H=H2, act_order="diag", damp=0.01)
    out["qws2_ls"] = rsc.quantize_with_scaling(W2, sc2, cb4, H=H2, act_order="diag", damp=0.01, nb_ls_moves=15)
    np.savez_compressed(os.path.join(OUT, "sweep.npz"), **out)


def pivot_ordering():
    """act_order = "pivot" (obq.py:140-166): the order itself and GPTQ with it."""
    out = {}
    W, H, m, _ = layer(31, 24, 96)
    cb = UniformCodebook(8, -1, 1)
    sc = rsc.compute_non_saturating_scaling(W, cb, 0)
    Ws = rsc.apply_scaling(W, sc, 0).astype(np.float32)
    Hd = H + 0.01 * H.diagonal().mean() * np.eye(H.shape[0])
    out.update(W=W, H=H, Ws=Ws, Hd=Hd)
    out["order_pivot"] = robq.compute_hessian_order(Ws, Hd, cb, "pivot")
    out["gptq_pivot"] = robq.quantize_opt(Ws, H, cb, act_order="pivot", damp=0.01)
    np.savez_compressed(os.path.join(OUT, "pivot.npz"), **out)


def local_search():
    out = {}
    W, H, m, X = layer(31, 12, 64)
    cb = UniformCodebook(4, -1, 1)
    sc = rsc.compute_non_saturating_scaling(W, cb, 0) * np.float32(0.6)
    Ws = rsc.apply_scaling(W, sc, 0)
    Q0 = cb(Ws)
    out["Ws"], out["H"], out["Q0"] = Ws, H, Q0
    out["gain_up"] = robq.compute_gain(Ws, Q0, H, cb.quantize_up(Q0))
    out["gain_down"] = robq.compute_gain(Ws, Q0, H, cb.quantize_down(Q0))
    for k in (1, 5, 30):
        out[f"ls_{k}"] = robq.quantize_local_search(Ws, Q0, H, cb, k)
    np.savez_compressed(os.path.join(OUT, "local_search.npz"), **out)


def obq_scaling():
    out = {}
    W, H, m, X = layer(41, 8, 80)
    cb = UniformCodebook(8, -1, 1)
    out["W"], out["H"] = W, H
    out["sc_obq"] = rsc.compute_obq_scaling(W, cb, 0, H=H, grid_size=12, min_factor=0.3)
    out["sc_obq_sqerr"] = rsc.compute_obq_scaling(W, cb, 0, H=H, grid_size=12, min_factor=0.3, act_order="sqerr", damp=0.03)
    np.savez_compressed(os.path.join(OUT, "obq_scaling.npz"), **out)


def statistics():
    out = {}
    torch.manual_seed(0)
    rng = np.random.default_rng(51)
    lin = torch.nn.Linear(48, 20)
    Wt = (0.05 * rng.standard_normal((20, 48))).astype(np.float32)
    bt = (0.1 * rng.standard_normal(20)).astype(np.float32)
    X1 = calib(rng, 40, 48).reshape(5, 8, 48)
    X2 = calib(rng, 24, 48)
    out["W"], out["b"], out["X1"], out["X2"] = Wt, bt, X1, X2
    with torch.no_grad():
        lin.weight.copy_(torch.from_numpy(Wt))
        lin.bias.copy_(torch.from_numpy(bt))
    st = Sleekit(lin)
    st.add_batch(torch.from_numpy(X1))
    out["count1"] = np.array(st.count)
    out["mean1"], out["hess1"] = st.mean.numpy().copy(), st.hessian.numpy().copy()
    st.add_batch(torch.from_numpy(X2))
    out["count2"] = np.array(st.count)
    out["mean2"], out["hess2"] = st.mean.numpy().copy(), st.hessian.numpy().copy()
    for name, fn in (("basic", "quantize_basic"), ("light", "quantize_sleekit_light"), ("heavy", "quantize_sleekit_heavy")):
        with torch.no_grad():
            lin.weight.copy_(torch.from_numpy(Wt))
            lin.bias.copy_(torch.from_numpy(bt))
        getattr(st, fn)(3)
        out[f"{name}_W"] = lin.weight.detach().numpy().copy()
        out[f"{name}_b"] = lin.bias.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "statistics.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "pivot":      # added later: only this fixture
        pivot_ordering()
        print("golden pivot vector written to", OUT)
        sys.exit(0)
    rounding()
    scales()
    factor_and_sweep()
    local_search()
    obq_scaling()
    statistics()
    with open(os.path.join(OUT, "VERSIONS.txt"), "w") as f:
        f.write(f"numpy {np.__version__}\ntorch {torch.__version__}\nreference {REF}\n")
    print("golden vectors written to", OUT)
