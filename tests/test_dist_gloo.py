"""World-size-2 gloo tests of the sharding logic (CPU; compute injected from the oracle).

They check that the three ways the path shards give the unsharded answer: sample-sharded
statistics + all-reduce, row-sharded quantization + all-gather (bitwise, rows never interact),
and round-robin layers."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sleekit_oracle as orc
from sleekit_b200 import dist as sdist
from sleekit_b200 import workloads as wl


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        grid = orc.UniformGrid(8, -1, 1)
        r, n, S = 10, 48, 96
        W, H_all, m_all, X = wl.synthetic_layer(r, n, 3, samples=S, want_x=True)

        # 1. calibration samples across ranks: uneven split, two batches on rank 0
        cuts = [0, 40, S]
        st = orc.RunningStats(n)
        mine = X[cuts[rank]: cuts[rank + 1]]
        if rank == 0:
            st.add_rows(mine[:15])
            st.add_rows(mine[15:])
        else:
            st.add_rows(mine)
        H, m, cnt = sdist.allreduce_statistics(torch.from_numpy(st.hessian), torch.from_numpy(st.mean), st.count)
        ref = orc.RunningStats(n)
        ref.add_rows(X)
        assert cnt == S
        np.testing.assert_allclose(H.numpy(), ref.hessian, rtol=2e-5, atol=1e-5)
        np.testing.assert_allclose(m.numpy(), ref.mean, rtol=2e-5, atol=1e-6)

        # 2. output rows across ranks: same H on every rank, rows quantized independently
        Hn = ref.hessian
        sc = orc.search_scale(W, grid, 0, H=Hn.diagonal())
        full = orc.quantize_scaled(W, sc, grid, H=Hn, rule="diag", damp=0.01)

        def fn(Wslice, a, b):
            return torch.from_numpy(orc.quantize_scaled(Wslice.numpy(), sc[a:b], grid, H=Hn, rule="diag", damp=0.01))

        got = sdist.quantize_rows_sharded(torch.from_numpy(W), fn)
        np.testing.assert_array_equal(got.numpy(), full)

        # err / sqerr ordering needs the column sums over all rows
        a, b = sdist.row_partition(r, world)[rank]
        Ws = orc.divide_rows(W, sc, 0)
        local = np.square(grid(Ws[a:b]) - Ws[a:b]).sum(axis=0)
        tot = sdist.allreduce_column_sums(torch.from_numpy(local.copy()))
        np.testing.assert_allclose(tot.numpy(), np.square(grid(Ws) - Ws).sum(axis=0), rtol=1e-6)

        rows = torch.from_numpy(orc.rowwise_error(W[a:b], full[a:b], Hn))
        mean = sdist.allreduce_row_error_mean(rows, r)
        np.testing.assert_allclose(float(mean), float(orc.mean_error(W, full, Hn)), rtol=1e-5)

        # 3. independent layers round-robin
        L = 5
        mine_ids = sdist.layers_of_rank(L, rank, world)
        vals = torch.tensor([10.0 * i + 1 for i in mine_ids])
        allv = sdist.gather_layer_values(vals, L)
        np.testing.assert_array_equal(allv.numpy(), [10.0 * i + 1 for i in range(L)])
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_row_partition_and_round_robin():
    assert sdist.row_partition(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert sdist.row_partition(8192, 8)[-1] == (7168, 8192)
    assert sdist.row_partition(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert sdist.layers_of_rank(72, 3, 8) == list(range(3, 72, 8))
    covered = sorted(i for k in range(8) for i in sdist.layers_of_rank(72, k, 8))
    assert covered == list(range(72))


def test_single_process_is_identity():
    H, m = torch.eye(3), torch.ones(3)
    assert sdist.allreduce_statistics(H, m, 7) == (H, m, 7)
    W = torch.arange(12.0).reshape(4, 3)
    out = sdist.quantize_rows_sharded(W, lambda x, a, b: x * 2)
    assert torch.equal(out, W * 2)


def test_world_size_2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
