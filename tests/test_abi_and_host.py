"""CPU-side checks: the C-ABI library builds, loads without a GPU and exports every symbol the
header declares; the host layer mirrors the reference's interface and fails loudly (no CPU
fallback) when no CUDA device is present."""

import ctypes
import inspect
import os

import numpy as np
import pytest
import torch

from sleekit_b200 import _lib


def test_library_builds_loads_and_exports_header_symbols():
    lib = _lib.load()
    declared = _lib.header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sleekit_b200.h but not exported"
    # every declared function is typed in the binding, and nothing extra is bound
    assert sorted(_lib.SIGNATURES) == declared
    assert lib.slk_abi_version() == 1


def test_workspace_queries_run_without_gpu():
    lib = _lib.load()
    assert lib.slk_hinv_ws_bytes(768) >= 3 * 768 * 768 * 8
    assert lib.slk_hinv_ws_bytes(100) >= 3 * 128 * 128 * 8  # padded to 64
    assert lib.slk_local_search_ws_bytes(16, 64) >= 16 * 64 * 4 + 64 * 4
    assert lib.slk_hweighted_error_ws_bytes(10, 1000, 4) >= 10 * 8 * 4
    assert lib.slk_scale_search_fullh_ws_bytes(8, 64, 100, 1) > 0


def test_argument_errors_are_reported_not_crashed():
    lib = _lib.load()
    cb = _lib.SlkCodebook(0, 1, -1.0, 1.0, 2.0, None, None)  # size 1 is invalid for a uniform grid
    rc = lib.slk_round_f32(None, 4, ctypes.byref(cb), 0, None, None, None)
    assert rc == -1
    assert b"codebook" in lib.slk_last_error() or b"size" in lib.slk_last_error()
    rc = lib.slk_gptq_sweep_f32(None, None, 4, 4, None, None, ctypes.byref(_lib.SlkCodebook(0, 4, -1.0, 1.0, 2 / 3, None, None)), 64, 8, 0, None)
    assert rc == -1 and b"leaf" in lib.slk_last_error()


def test_public_interface_matches_reference_signatures():
    # names, argument order and defaults of SURVEY.md section 8b
    from sleekit_b200 import codebook, obq, scaling, statistics

    def sig(f):
        return [(p.name, p.default) for p in inspect.signature(f).parameters.values()]

    E = inspect.Parameter.empty
    assert sig(obq.quantize_opt) == [("W", E), ("H", E), ("quantizer", E), ("act_order", "diag"), ("damp", 0.01),
                                     ("nb_ls_moves", 0), ("min_block_size", 32), ("num_blocks", 8)]
    assert sig(scaling.quantize_with_scaling) == [("data", E), ("scale", E), ("quantizer", E), ("H", None),
                                                  ("act_order", "diag"), ("damp", 0.01), ("nb_ls_moves", 0)]
    assert sig(scaling.compute_min_mse_scaling) == [("data", E), ("codebook", E), ("axis", 0), ("H", None),
                                                    ("min_factor", 0.05), ("max_factor", 1.0), ("grid_size", 100)]
    assert sig(scaling.compute_obq_scaling) == [("data", E), ("codebook", E), ("axis", E), ("H", E), ("damp", 0.01),
                                                ("act_order", "diag"), ("min_factor", 0.05), ("max_factor", 1.0),
                                                ("grid_size", 100)]
    assert sig(scaling.compute_scaling) == [("data", E), ("codebook", E), ("H", E), ("mode", "mse"), ("axis", 0),
                                            ("min_factor", 0.05), ("max_factor", 1.0), ("grid_size", 100)]
    assert sig(statistics.Sleekit.quantize)[1:] == [("nbits", E), ("scaling_mode", "mse"), ("order_mode", "diag"),
                                                    ("bias_correction", False), ("damp", 0.01), ("nb_ls_moves", 0),
                                                    ("grid_size", 100), ("min_factor", 0.05), ("max_factor", 1.0)]
    for name in ("random_psd_matrix", "remove_input_bias", "remove_dead_values", "compute_hessian_chol",
                 "compute_hessian_order", "channelwise_error", "quantization_error", "compute_gain",
                 "LocalSearchQuantizer", "quantize_local_search", "_quantize_opt_block", "_quantize_opt_core"):
        assert hasattr(obq, name)
    for name in ("apply_scaling", "apply_scaling_in_place", "compute_norm_scaling",
                 "compute_non_saturating_scaling", "_quantize_opt_block"):
        assert hasattr(scaling, name)
    for name in ("UniformCodebook", "Codebook", "lloyd_max"):
        assert hasattr(codebook, name)
    cb = codebook.UniformCodebook(8, -1, 1)
    assert len(cb) == 8 and cb.min() == -1 and cb.max() == 1 and cb.zero == -1
    assert cb.scale == pytest.approx(2 / 7)
    np.testing.assert_allclose(cb.values, np.linspace(-1, 1, 8))
    tb = codebook.Codebook([4.0, -1.0, 2.0, 8.0])
    np.testing.assert_array_equal(tb.values, [-1, 2, 4, 8])
    np.testing.assert_array_equal(tb.thresholds, [0.5, 3, 6])
    assert len(codebook.Codebook.nf4()) == 16


def test_star_imports_leak_np_like_the_reference():
    # experiments/*.py never import numpy: they rely on `np` leaking through the star imports
    ns = {}
    exec("from sleekit.codebook import *\nfrom sleekit.obq import *\nfrom sleekit.scaling import *", ns)
    assert ns["np"] is np
    assert "UniformCodebook" in ns and "quantize_opt" in ns and "compute_obq_scaling" in ns
    import sleekit

    assert sleekit.Sleekit.__module__ == "sleekit_b200.statistics"


def test_errors_follow_reference_conventions():
    from sleekit_b200 import codebook, scaling

    with pytest.raises(AssertionError):
        codebook.UniformCodebook(1, -1, 1)
    with pytest.raises(AssertionError):
        codebook.UniformCodebook(4, 1, -1)
    with pytest.raises(RuntimeError):
        scaling.compute_non_saturating_scaling(np.ones((2, 2), np.float32), codebook.UniformCodebook(4, 0, 1))
    with pytest.raises(ValueError):
        from sleekit_b200 import Sleekit

        Sleekit(torch.nn.Embedding(4, 4))


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_hot_path_fails_loudly_without_cuda():
    from sleekit_b200 import codebook, obq, scaling

    cb = codebook.UniformCodebook(8, -1, 1)
    x = np.zeros((4, 4), np.float32)
    for call in (lambda: cb(x), lambda: scaling.compute_min_mse_scaling(x, cb),
                 lambda: obq.quantize_opt(x, np.eye(4, dtype=np.float32), cb),
                 lambda: obq.channelwise_error(x, x, np.eye(4, dtype=np.float32))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for pkg in ("sleekit_b200", "sleekit"):
        for dirpath, _, files in os.walk(os.path.join(root, pkg)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in text and "from oracle" not in text, os.path.join(dirpath, f)


def test_codebook_breakpoints_are_exact_on_the_host():
    """The kernels round through X[k] = min{x : index(x) >= k}, found on the host by bisection with
    IEEE fp32 ops; check against the oracle's quantize_index (codebook.py:43-54) on both sides of
    every breakpoint and on random values -- pure host code, no CUDA call."""
    import ctypes as C

    import numpy as np

    from oracle import sleekit_oracle as orc
    from sleekit_b200 import _lib
    from sleekit_b200._lib import SlkCodebook

    lib = _lib.load()
    rng = np.random.default_rng(5)
    for c, lo, hi in [(2, -1, 1), (3, -1, 1), (4, -1, 1), (8, -1, 1), (16, -1, 1), (9, -2, 2), (7, -0.5, 3.0)]:
        step = (hi - lo) / (c - 1)
        cb = SlkCodebook(0, c, lo, hi, step, None, None)
        out = (C.c_float * 16)()
        assert lib.slk_codebook_breaks_host(C.byref(cb), out) == 0
        X = np.array(out[:], dtype=np.float32)
        grid = orc.UniformGrid(c, lo, hi)
        assert np.all(np.isinf(X[c:])) and np.isinf(X[0])
        assert np.all(np.diff(X[1:c]) > 0)
        for k in range(1, c):
            at = X[k]
            below = np.nextafter(at, np.float32(-np.inf))
            assert grid.index(np.array([at], np.float32))[0] >= k
            assert grid.index(np.array([below], np.float32))[0] < k
        x = (rng.standard_normal(200000) * (hi - lo)).astype(np.float32)
        want = grid.index(x).astype(np.int64)
        got = (x[:, None] >= X[None, 1:c]).sum(axis=1)
        np.testing.assert_array_equal(got, want)


def test_issue_orders_are_permutations_with_the_stated_structure():
    from sleekit_b200 import workloads as wl
    from sleekit_b200.pipeline import issue_order

    shapes = wl.layer_shapes("opt-125m")
    for order in ("big", "interleaved", "model"):
        idx = issue_order(shapes, order)
        assert sorted(idx) == list(range(len(shapes)))
    big = issue_order(shapes, "big")
    assert [shapes[k][1] for k in big[:12]] == [3072] * 12                  # longest chains first
    assert [shapes[k] for k in big[12:24]] == [(3072, 768)] * 12            # then more rows first
    inter = issue_order(shapes, "interleaved")
    pos = [p for p, k in enumerate(inter) if shapes[k][1] == 3072]
    assert pos == list(range(0, 72, 6))                                     # one long chain, then five short ones
    assert issue_order([(8, 16), (8, 16)], "interleaved") == [0, 1]         # nothing to interleave
    with pytest.raises(ValueError):
        issue_order(shapes, "random")


def test_symmetric_upload_accounting_without_gpu():
    """slk_upload_symmetric_bytes: bytes of the block upper triangle (host arithmetic only)."""
    from sleekit_b200 import ops

    lib = _lib.load()
    for n in (32, 100, 768, 1000, 3072, 28672):
        bs = ops.symmetric_block_rows(n)
        assert bs % 32 == 0 and bs >= 32
        want = sum(min(bs, n - r0) * (n - r0) * 4 for r0 in range(0, n, bs))
        assert lib.slk_upload_symmetric_bytes(n, bs) == want
        assert want <= 4 * n * n
    assert lib.slk_upload_symmetric_bytes(3072, ops.symmetric_block_rows(3072)) / (4 * 3072 * 3072) < 0.54
    # argument errors are reported, not crashed
    assert lib.slk_upload_symmetric_f32(None, None, 16, 32, None) != 0


def test_upload_cache_fingerprint_detects_in_place_edits():
    """_convert._fingerprint (host only): the sampled content check of the identity-keyed upload cache changes
    under the in-place edits the reference's callers make -- dampening the diagonal (obq.py:45-47 style),
    rescaling or shifting the whole array, overwriting its tail -- and is stable for untouched content,
    for 1-D, rectangular, square and non-contiguous arrays."""
    from sleekit_b200 import _convert as cv

    rng = np.random.default_rng(0)
    for shape in [(3072, 3072), (768, 3072), (100000,), (5, 7)]:
        a = rng.standard_normal(shape).astype(np.float32)
        f0 = cv._fingerprint(a)
        assert cv._fingerprint(a) == f0 and cv._fingerprint(a.copy()) == f0
        b = a.copy()
        b *= np.float32(1.5)
        assert cv._fingerprint(b) != f0
        b = a.copy()
        b.reshape(-1)[-3:] += 1
        assert cv._fingerprint(b) != f0
        if len(shape) == 2 and shape[0] == shape[1]:
            b = a.copy()
            b[np.arange(shape[0]), np.arange(shape[0])] += np.float32(0.01)
            assert cv._fingerprint(b) != f0
    t = rng.standard_normal((64, 128)).astype(np.float32).T      # a transposed view
    assert cv._fingerprint(t) == cv._fingerprint(np.ascontiguousarray(t))
