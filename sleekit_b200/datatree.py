"""The reference's on-disk statistics tree as the input of the layer-set driver.

``Sleekit.export(path, npy_format=True)`` (statistics.py:89-105) writes, per layer, ``weight.npy``,
``hessian.npy``, ``mean.npy`` and ``bias.npy`` into one directory; the experiment scripts walk such
a tree (``data/<model>/<layer>/``) and quantize layer after layer from host arrays
(experiments/compare.py:37-53, 84-95).  This module feeds the same tree through
``LayerSetQuantizer.host_plan``: the layers are cut into batches that fit a pinned-memory budget;
while the GPU replays the graph of batch k, a loader thread reads the files of batch k+1 straight
into the page-locked buffers of a second plan (double buffering), so that disk reads, PCIe copies
and kernels overlap.  The selection and preparation of the layers is the reference's:

* a directory is a layer iff it holds weight.npy, hessian.npy and mean.npy, and the layers are
  visited in sorted order                                    (experiments/compare.py:39-46)
* arrays are cast to fp32                                    (experiments/compare.py:51-53)
* remove_dead_values(H, W) in place                           (experiments/compare.py:54, obq.py:28-35)
* optional H - mean mean^T                                    (experiments/compare.py:55, obq.py:14-25)

Nothing here computes on the CPU beyond the O(n) dead-input fix-up the reference also does on the
host; the quantization itself is the CUDA path (``require_cuda`` raises without a device).
"""

from __future__ import annotations

import os
import threading

import numpy as np

REQUIRED = ("weight.npy", "hessian.npy", "mean.npy")


def find_layers(root_dir):
    """Sorted layer directories of a statistics tree (experiments/compare.py:39-46)."""
    return sorted(root for root, _dirs, files in sorted(os.walk(root_dir)) if all(f in files for f in REQUIRED))


def layer_shape(layer_dir):
    """(rows, cols) of a layer's weight, read from the .npy header only."""
    w = np.load(os.path.join(layer_dir, "weight.npy"), mmap_mode="r")
    assert w.ndim == 2, f"{layer_dir}: weight must be 2-D"
    return int(w.shape[0]), int(w.shape[1])


def load_layer_into(layer_dir, W_out, H_out, correct_input_bias=False):
    """Read one layer into caller-owned fp32 buffers (the plan's pinned numpy views) the way the
    reference's scripts prepare it: cast to fp32, dead inputs fixed in place, optionally
    H - mean mean^T.  Returns the mean vector (fp32)."""
    W_out[...] = np.load(os.path.join(layer_dir, "weight.npy"), mmap_mode="r")      # compare.py:51
    H_out[...] = np.load(os.path.join(layer_dir, "hessian.npy"), mmap_mode="r")     # compare.py:52
    mean = np.load(os.path.join(layer_dir, "mean.npy")).astype(np.float32)          # compare.py:53
    assert H_out.shape == (W_out.shape[1], W_out.shape[1]) and mean.shape == (W_out.shape[1],)
    d = H_out.diagonal()                                                            # obq.py:28-35
    dead = d == 0
    if dead.any():
        H_out[dead, dead] = d.mean()
        W_out[:, dead] = 0
    if correct_input_bias:                                                          # obq.py:14-25
        H_out -= np.outer(mean, mean)
    return mean


def plan_batches(shapes, budget_bytes):
    """Cut the layer list into consecutive batches whose pinned buffers (W, H and Q of every layer,
    fp32) stay under `budget_bytes`; a layer larger than the budget gets a batch of its own."""
    batches, cur, used = [], [], 0
    for i, (r, n) in enumerate(shapes):
        need = 4 * (2 * r * n + n * n)
        if cur and used + need > budget_bytes:
            batches.append(cur)
            cur, used = [], 0
        cur.append(i)
        used += need
    if cur:
        batches.append(cur)
    return batches


def quantize_tree(root_dir, codebook, scaling_mode="diag", act_order="diag", damp=0.01, grid_size=100,
                  min_factor=0.05, max_factor=1.0, correct_input_bias=False, budget_bytes=2 << 30,
                  out_name=None, streams=72):
    """Quantize every layer of a statistics tree.  Returns [(relative layer name, layer error)] in the
    reference's order; with `out_name` the de-scaled quantized weights are also written next to
    each layer's inputs as `<out_name>.npy`.

    Per batch: loader thread -> pinned buffers of plan A | GPU replays plan B's graph (H2D, scale
    search, GPTQ, layer error, D2H).  Plans are cached per shape list, so a model made of identical
    blocks records two graphs in total."""
    from . import ops
    from .pipeline import LayerSetQuantizer

    ops.require_cuda()
    layers = find_layers(root_dir)
    shapes = [layer_shape(d) for d in layers]
    batches = plan_batches(shapes, budget_bytes)
    lsq = LayerSetQuantizer(codebook, scaling_mode=scaling_mode, act_order=act_order, damp=damp, nb_ls_moves=0,
                            grid_size=grid_size, min_factor=min_factor, max_factor=max_factor, streams=streams)
    plans = {}          # (parity, shapes of the batch) -> HostPlan; two parities = double buffering

    def plan_for(k):
        key = (k & 1, tuple(shapes[i] for i in batches[k]))
        if key not in plans:
            plans[key] = lsq.host_plan(list(key[1]))
        return plans[key]

    def load(k, plan):
        for slot, i in enumerate(batches[k]):
            load_layer_into(layers[i], plan.W[slot], plan.H[slot], correct_input_bias)

    results = []
    if not batches:
        return results
    cur = plan_for(0)
    load(0, cur)
    for k in range(len(batches)):
        nxt, loader = None, None
        if k + 1 < len(batches):
            nxt = plan_for(k + 1)
            loader = threading.Thread(target=load, args=(k + 1, nxt))
            loader.start()                      # disk -> pinned memory while the GPU works on batch k
        Q, err = cur.run()
        for slot, i in enumerate(batches[k]):
            results.append((os.path.relpath(layers[i], root_dir), float(err[slot])))
            if out_name:
                np.save(os.path.join(layers[i], out_name + ".npy"), Q[slot])
        if loader is not None:
            loader.join()
        cur = nxt
    return results
