"""Multi-GPU sharding of the quantization path (one process per GPU, torch.distributed).

The path shards three ways, each with at most one exchange step:

* calibration samples across ranks  -> one all-reduce of the partial statistics
  (``allreduce_statistics``): every rank accumulates ``Sleekit``-style running means over its own
  rows of X; the count-weighted sums ``[count*H, count*mean, count]`` are packed into one
  ``[n*n + n + 1]`` buffer and summed (NCCL over NVLink on GPUs);
* output rows of W across ranks     -> no communication inside scale search / sweep / local search
  (rows never interact, obq.py:106-137, :264-346, scaling.py:127-133); H and U are replicated.
  ``quantize_rows_sharded`` runs a per-rank function on its row slice and all-gathers the rows;
  the only reductions are the column sums needed by ``act_order in {"err", "sqerr"}``
  (``allreduce_column_sums``) and the layer error (``allreduce_row_error_mean``);
* independent layers round-robin    -> no collective at all, errors gathered at the end
  (``layers_of_rank`` / ``gather_layer_values``).

Everything here is backend-agnostic plumbing (NCCL on GPUs, gloo in the CPU tests): the compute is
whatever callable the caller passes in (the sleekit_b200 device functions in production).
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def _world(group=None):
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def row_partition(rows, world):
    """Balanced contiguous row ranges: [(start, stop)] * world, sizes differ by at most one."""
    base, extra = divmod(int(rows), int(world))
    out, start = [], 0
    for k in range(world):
        size = base + (1 if k < extra else 0)
        out.append((start, start + size))
        start += size
    return out


def layers_of_rank(num_layers, rank, world):
    """Round-robin layer assignment: rank k owns layers k, k + world, ..."""
    return list(range(rank, num_layers, world))


def allreduce_statistics(hessian, mean, count, group=None, peer_buffer=None, barrier=None):
    """Combine per-rank running means (statistics.py:76-87 semantics) into the global ones, IN PLACE.

    hessian [n, n], mean [n] fp32 are this rank's running means over `count` samples; on return they
    hold the running means over the union of all ranks' samples (identical on every rank).  Returns
    (hessian, mean, total count).  The sample counts travel first (one scalar all-reduce), every rank
    then weights its statistics by count / total so that ONE sum all-reduce yields the global means
    with no temporaries of the size of H.  On CUDA the Hessian -- symmetric -- is exchanged as its block
    upper triangle (ops.sym_pack / sym_unpack: ~53 % of the bytes, all-reduced in place in the packed
    buffer); on CPU tensors (gloo tests) the full matrix is reduced in place.  peer_buffer (ops.PeerBuffer of
    at least sym_packed_len(n) + n floats) + barrier: the packed statistics are summed by our own NVLink
    kernel (slk_peer_allreduce_f32: every rank reduces one slice with peer loads and stores the sum into all
    copies) instead of NCCL."""
    rank, world = _world(group)
    if world == 1:
        return hessian, mean, count
    n = mean.numel()
    tot = torch.tensor([float(count)], dtype=torch.float64, device=hessian.device)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    total = float(tot.item())
    if total == 0:
        return hessian.zero_(), mean.zero_(), 0
    w = float(count) / total
    if hessian.is_cuda:
        from . import ops

        L = ops.sym_packed_len(n)
        if peer_buffer is not None:
            assert peer_buffer.count >= L + n and barrier is not None
            buf = peer_buffer.tensor()
            if peer_buffer.count > L + n:
                buf[L + n:].zero_()
        else:
            buf = torch.empty(L + n, dtype=torch.float32, device=hessian.device)
        ops.sym_pack(hessian, buf, w)
        torch.mul(mean, w, out=buf[L:L + n])
        if peer_buffer is not None:
            peer_buffer.allreduce(barrier)
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        ops.sym_unpack(buf, hessian, 1.0)
        mean.copy_(buf[L:L + n])
    else:
        hessian.mul_(w)
        mean.mul_(w)
        dist.all_reduce(hessian, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mean, op=dist.ReduceOp.SUM, group=group)
    return hessian, mean, int(round(total))


def allreduce_column_sums(local_sums, group=None):
    """Sum of per-rank column sums (the err / sqerr ordering keys, obq.py:60-69)."""
    rank, world = _world(group)
    if world > 1:
        dist.all_reduce(local_sums, op=dist.ReduceOp.SUM, group=group)
    return local_sums


def allreduce_row_error_mean(local_row_errors, total_rows, group=None):
    """Mean channel error over all rows (obq.py:98-103) from per-rank row errors."""
    s = local_row_errors.double().sum().reshape(1)
    rank, world = _world(group)
    if world > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    return (s / float(total_rows)).to(local_row_errors.dtype)[0]


def quantize_rows_sharded(W, fn, group=None):
    """Run `fn(W[start:stop], start, stop)` on this rank's row slice and all-gather the result rows.

    `fn` must return a [stop-start, n] tensor on W's device.  Rows never interact in the path, so
    the gathered matrix equals the unsharded result row for row."""
    rank, world = _world(group)
    r, n = W.shape
    parts = row_partition(r, world)
    start, stop = parts[rank]
    local = fn(W[start:stop].contiguous(), start, stop)
    if world == 1:
        return local
    width = max(b - a for a, b in parts)
    padded = torch.zeros((width, n), dtype=local.dtype, device=local.device)
    padded[: stop - start] = local
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)
    return torch.cat([g[: b - a] for g, (a, b) in zip(gathered, parts)], dim=0)


def gather_layer_values(local_values, num_layers, group=None):
    """Round-robin layers: rank k computed values for layers k, k+world, ...; return all of them in
    layer order on every rank.  local_values: 1-D tensor, one entry per owned layer."""
    rank, world = _world(group)
    if world == 1:
        return local_values
    per = (num_layers + world - 1) // world
    padded = torch.zeros(per, dtype=local_values.dtype, device=local_values.device)
    padded[: local_values.numel()] = local_values
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)
    out = torch.empty(num_layers, dtype=local_values.dtype, device=local_values.device)
    for k in range(world):
        idx = layers_of_rank(num_layers, k, world)
        out[idx] = gathered[k][: len(idx)]
    return out
