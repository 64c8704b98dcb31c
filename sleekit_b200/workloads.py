"""Synthetic workloads of BASELINE.json's configs.

Pure numpy input generation shared by tests and bench.py so that every arm sees
byte-identical weights and calibration rows.  No arithmetic of the hot path lives here.
"""

import numpy as np


def synthetic_layer(rows, cols, layer_id, samples=2048, want_x=False):
    """W [rows, cols] fp32, H = X^T X / S fp32, mean [cols] fp32 (and X [S, cols] on request).

    W = 0.02 N(0,1), seed 1000+id.  X, seed 2000+id: a rank-64 correlated part plus noise,
    log-normal per-channel scales (outlier channels) and a non-zero mean."""
    W = (0.02 * np.random.default_rng(1000 + layer_id).standard_normal((rows, cols))).astype(np.float32)
    X = synthetic_calibration(cols, layer_id, samples)
    H = (X.T @ X) / np.float32(samples)
    m = X.mean(axis=0, dtype=np.float32)
    if want_x:
        return W, H, m, X
    return W, H, m


def synthetic_weight(rows, cols, layer_id):
    return (0.02 * np.random.default_rng(1000 + layer_id).standard_normal((rows, cols))).astype(np.float32)


def synthetic_calibration(cols, layer_id, samples=2048):
    g = np.random.default_rng(2000 + layer_id)
    base = g.standard_normal((samples, 64)).astype(np.float32) @ g.standard_normal((64, cols)).astype(np.float32)
    X = base + np.float32(0.3) * g.standard_normal((samples, cols)).astype(np.float32)
    X = X * np.exp(g.standard_normal(cols)).astype(np.float32) + np.float32(0.5)
    return np.ascontiguousarray(X, dtype=np.float32)


OPT125M_BLOCK = [(768, 768)] * 4 + [(3072, 768), (768, 3072)]


def layer_shapes(model="opt-125m"):
    """[out, in] = [r, n] of every linear layer, in the order of the reference's result files."""
    if model == "opt-125m":
        return OPT125M_BLOCK * 12
    if model == "opt-350m":
        return ([(1024, 1024)] * 4 + [(4096, 1024), (1024, 4096)]) * 24
    if model == "bloom-560m":
        return [(3072, 1024), (1024, 1024), (4096, 1024), (1024, 4096)] * 24
    if model == "llama2-7b-mlp":
        return [(11008, 4096), (11008, 4096), (4096, 11008)] * 32
    if model == "llama3-70b-down":
        return [(8192, 28672)]
    raise ValueError(model)
