"""Experiment: raw pinned H2D bandwidth on the box (context for the e2e number)."""
import time, torch
for gb in (0.25, 2.0):
    n = int(gb * (1 << 30) / 4)
    h = torch.empty(n, dtype=torch.float32, pin_memory=True); h.fill_(1.0)
    d = torch.empty(n, dtype=torch.float32, device="cuda")
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"H2D pinned {gb} GiB: {gb * 1.0737 / dt:.1f} GB/s")
    t0 = time.perf_counter()
    for _ in range(3): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"D2H pinned {gb} GiB: {gb * 1.0737 / dt:.1f} GB/s")
