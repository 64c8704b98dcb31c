"""Peer-to-peer copy bandwidth between GPU 0 and the other GPUs of the box (development aid)."""
import torch

n = torch.cuda.device_count()
src = torch.empty(1 << 28, dtype=torch.uint8, device="cuda:0")
for d in range(1, n):
    dst = torch.empty_like(src, device=f"cuda:{d}")
    for _ in range(2):
        dst.copy_(src)
    torch.cuda.synchronize(0); torch.cuda.synchronize(d)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.device(0):
        e0.record()
        for _ in range(10):
            dst.copy_(src)
        e1.record()
        torch.cuda.synchronize(0)
    print(f"cuda:0 -> cuda:{d}: {10 * src.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s, can_access_peer={torch.cuda.can_device_access_peer(0, d)}")
