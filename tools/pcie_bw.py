"""Developer probe: pinned H2D / D2H bandwidth of the box (floor of bench.py's e2e)."""
import time, torch
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, src, dst in (("h2d", h, d), ("d2h", d, h)):
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print(name, f"{n / dt / 1e9:.1f} GB/s")
# both directions at once
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print("bidir each", f"{n / dt / 1e9:.1f} GB/s")
# small 2.4 MB copies
m = 1228800 * 2
torch.cuda.synchronize(); t = time.perf_counter()
for i in range(400):
    h[i * m:(i + 1) * m].copy_(d[i * m:(i + 1) * m], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("d2h 2.4MB chunks", f"{400 * m / dt / 1e9:.1f} GB/s")
