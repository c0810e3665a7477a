"""Pinned host->device copy bandwidth of this box (what bounds bench.py's e2e).  Prints GB/s for a few sizes."""
import torch, time
for mb in (1, 4, 64, 512):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    reps = max(4, 2048 // mb)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(f"H2D {mb:4d} MiB x{reps}: {n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s", flush=True)
