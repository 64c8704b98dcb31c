"""The reference's on-disk statistics tree (Sleekit.export, statistics.py:89-105; walked by
experiments/compare.py:37-55) as the input of the layer-set driver: host-side logic on CPU, the
quantized result against the per-layer API on the GPU."""

import os

import numpy as np
import pytest

from sleekit_b200 import datatree
from sleekit_b200 import workloads as wl


def make_tree(tmp, shapes, dead_in=None, samples=256):
    layers = {}
    for i, (r, n) in enumerate(shapes):
        W, H, m = wl.synthetic_layer(r, n, 70 + i, samples=samples)
        if dead_in == i:
            H[5, :] = 0
            H[:, 5] = 0
        d = os.path.join(tmp, "model", f"{i // 2}", f"layer{i % 2}")
        os.makedirs(d)
        np.save(os.path.join(d, "weight.npy"), W.astype(np.float64) if i == 1 else W)   # any float dtype on disk
        np.save(os.path.join(d, "hessian.npy"), H)
        np.save(os.path.join(d, "mean.npy"), m)
        np.save(os.path.join(d, "bias.npy"), np.zeros(r, dtype=np.float32))
        layers[d] = (W, H, m)
    # a directory without mean.npy is not a layer (experiments/compare.py:43)
    bad = os.path.join(tmp, "model", "0", "incomplete")
    os.makedirs(bad)
    np.save(os.path.join(bad, "weight.npy"), np.zeros((2, 2), dtype=np.float32))
    np.save(os.path.join(bad, "hessian.npy"), np.zeros((2, 2), dtype=np.float32))
    return layers


def test_find_and_load_follow_the_reference_scripts(tmp_path):
    shapes = [(8, 32), (16, 32), (8, 64)]
    layers = make_tree(str(tmp_path), shapes, dead_in=2)
    found = datatree.find_layers(str(tmp_path))
    # the reference's own expression (experiments/compare.py:39-46)
    want = sorted(root for root, dirs, files in sorted(os.walk(str(tmp_path)))
                  if "weight.npy" in files and "hessian.npy" in files and "mean.npy" in files)
    assert found == want and len(found) == 3
    assert [datatree.layer_shape(d) for d in found] == [layers[d][0].shape for d in found]
    for d in found:
        W0, H0, m0 = layers[d]
        r, n = W0.shape
        for corr in (False, True):
            W, H = np.empty((r, n), dtype=np.float32), np.empty((n, n), dtype=np.float32)
            mean = datatree.load_layer_into(d, W, H, correct_input_bias=corr)
            # experiments/compare.py:51-55 + obq.py:14-35, restated
            Wr, Hr = W0.astype(np.float32).copy(), H0.astype(np.float32).copy()
            dead = Hr.diagonal() == 0
            Hr[dead, dead] = Hr.diagonal().mean()
            Wr[:, dead] = 0
            if corr:
                Hr = Hr - np.outer(m0, m0)
            np.testing.assert_array_equal(W, Wr)
            np.testing.assert_array_equal(H, Hr)
            np.testing.assert_array_equal(mean, m0)
            assert np.array_equal(H, H.T)


def test_plan_batches():
    shapes = [(8, 32)] * 5 + [(64, 256)] + [(8, 32)]
    small = 4 * (2 * 8 * 32 + 32 * 32)
    b = datatree.plan_batches(shapes, 2 * small)
    assert b == [[0, 1], [2, 3], [4], [5], [6]]          # the big layer exceeds the budget: own batch
    assert datatree.plan_batches(shapes, 1 << 40) == [list(range(7))]
    assert datatree.plan_batches([], 100) == []


@pytest.mark.gpu
def test_quantize_tree_equals_per_layer_api(tmp_path):
    from sleekit_b200 import codebook, obq, scaling

    shapes = [(48, 256), (64, 128), (48, 256), (64, 128), (32, 512)]
    make_tree(str(tmp_path), shapes, dead_in=3, samples=512)
    cb = codebook.UniformCodebook(8, -1, 1)
    small = 4 * (2 * 48 * 256 + 256 * 256)
    res = datatree.quantize_tree(str(tmp_path), cb, budget_bytes=2 * small, out_name="weight_q", streams=4)
    found = datatree.find_layers(str(tmp_path))
    assert [name for name, _ in res] == [os.path.relpath(d, str(tmp_path)) for d in found]
    for (name, err), d in zip(res, found):
        W = np.load(os.path.join(d, "weight.npy")).astype(np.float32)
        H = np.load(os.path.join(d, "hessian.npy")).astype(np.float32)
        obq.remove_dead_values(H, W)                                           # experiments/compare.py:54
        sc = scaling.compute_min_mse_scaling(W, cb, 0, H=H.diagonal())
        want = scaling.quantize_with_scaling(W, sc, cb, H=H)
        got = np.load(os.path.join(d, "weight_q.npy"))
        np.testing.assert_array_equal(got, want)
        assert abs(err - obq.quantization_error(W, want, H)) <= 1e-4 * abs(err)
