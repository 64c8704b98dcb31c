import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def golden():
    return load_golden


# ---------------------------------------------------------------------------
# parity report: every GPU test that compares the CUDA path with the oracle / golden vectors records
# the MEASURED agreement next to the bar it asserts; written at session end (profiles/ keeps a copy)
# ---------------------------------------------------------------------------
PARITY = []


def record_parity(test, what, measured, bar):
    PARITY.append({"test": test, "what": what, "measured": float(measured), "bar": bar})


def scales_equivalent(got, want, row_error, test, what, min_same=0.999, tie_tol=2e-5):
    """Chosen scale-grid points: `got` (CUDA) against `want` (oracle / golden), both [rows].
    Bar: identical in >= min_same of the rows, and every differing row is a TIE -- the oracle's own
    per-row error (row_error(scales) -> [rows], scaling.py:84-95) at the two scales differs by less
    than tie_tol relative, i.e. by less than the summation-order noise of the fp32 error sums, so
    the reference's own choice between them depends on its BLAS (scaling.py:131 keeps the first
    strict minimum)."""
    got, want = np.asarray(got), np.asarray(want)
    same = float((got == want).mean())
    worst = 0.0
    diff = np.nonzero(got != want)[0]
    if diff.size:
        eg, ew = np.asarray(row_error(got), dtype=np.float64), np.asarray(row_error(want), dtype=np.float64)
        worst = float(np.max(np.abs(eg[diff] - ew[diff]) / np.maximum(np.abs(ew[diff]), 1e-300)))
    record_parity(test, what + ": rows with the reference's grid point", same, min_same)
    record_parity(test, what + ": worst relative error gap on differing rows (ties)", worst, tie_tol)
    assert same >= min_same, (what, same)
    assert worst <= tie_tol, (what, worst, diff[:8])
    return same


def pytest_sessionfinish(session, exitstatus):
    if not PARITY:
        return
    import json

    out = os.environ.get("SLK_PARITY_REPORT") or os.path.join(ROOT, "gpurun_out", "parity_report.json")
    try:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        with open(out, "w") as f:
            json.dump(PARITY, f, indent=1)
    except OSError:
        pass


def _has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
