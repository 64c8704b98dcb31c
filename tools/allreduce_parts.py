"""Times the pieces of the sample-sharded statistics exchange at n = 28672 (development aid):
pack, unpack + mirror, peer all-reduce kernel, NCCL all-reduce.  python / torchrun tools/allreduce_parts.py [n]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleekit_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28672
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", torch.cuda.current_device())
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
H = torch.randn((n, n), device=dev)
L = ops.sym_packed_len(n)
token = torch.zeros(1, device=dev)


def barrier():
    if world > 1:
        dist.all_reduce(token)


pb = ops.PeerBuffer(L + n) if world > 1 else None
buf = pb.tensor() if pb is not None else torch.empty(L + n, device=dev)
nbuf = torch.empty(L + n, device=dev)


def timed(name, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"{name:32s} {e0.elapsed_time(e1) / reps:8.3f} ms")


timed("sym_pack", lambda: ops.sym_pack(H, buf, 0.125))
timed("sym_unpack + mirror", lambda: ops.sym_unpack(buf, H, 1.0))
timed("barrier (token all-reduce)", barrier)
if world > 1:
    timed("peer all-reduce (+2 barriers)", lambda: pb.allreduce(barrier))
    timed("NCCL all-reduce", lambda: dist.all_reduce(nbuf))
    if rank == 0:
        print(f"bytes {4 * (L + n) / 1e9:.3f} GB, ranks {world}")
    pb.close()
    dist.destroy_process_group()
