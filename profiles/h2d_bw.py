"""Ceiling for the end-to-end number: pinned host -> device copy bandwidth on this box."""
import torch, time
for mb in (256, 1024, 4096):
    h = torch.empty(mb * 1024 * 1024, dtype=torch.uint8).pin_memory()
    d = torch.empty_like(h, device='cuda')
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(f"{mb} MiB pinned H2D: {4 * h.numel() / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
    del h, d
