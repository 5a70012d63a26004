"""H2D / D2H bandwidth of pinned host memory on this box (context for bench.py's e2e number)."""
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, (a, b) in {"h2d": (d, h), "d2h": (h, d)}.items():
    for sz in (1 << 24, 1 << 26, 1 << 28, 1 << 30):
        a[:sz].copy_(b[:sz], non_blocking=True); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            a[:sz].copy_(b[:sz], non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        print(name, sz >> 20, "MiB", "%.1f GB/s" % (4 * sz / (e0.elapsed_time(e1) * 1e-3) / 1e9))

# The same H2D copy while another stream keeps HBM busy (what ldvb_push's copies see while the chain runs).
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"); y = torch.empty_like(x)
side = torch.cuda.Stream()
sz = 1 << 30
for busy in (False, True):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if busy:
        with torch.cuda.stream(side):
            for _ in range(60):
                y.copy_(x)
    e0.record()
    for _ in range(4):
        d[:sz].copy_(h[:sz], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print("h2d 1024 MiB", "with a device-to-device copy loop on another stream" if busy else "alone", "%.1f GB/s" % (4 * sz / (e0.elapsed_time(e1) * 1e-3) / 1e9))
