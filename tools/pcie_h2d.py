"""Raw host-to-device bandwidth of pinned copies at the sizes of one tracked frame's inputs (vertex / normal maps 4.9 MB, colour 1.2 MB,
depth 0.6 MB): the ceiling of bench.py's e2e.  python tools/pcie_h2d.py"""
import torch, time
x = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
y = torch.empty_like(x, device="cuda")
for n in (256 << 20, 4900000, 1228800, 614400):
    torch.cuda.synchronize()
    reps = max(4, (1 << 30) // n)
    t0 = time.perf_counter()
    for i in range(reps):
        y[:n].copy_(x[:n], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"H2D {n/1e6:8.2f} MB copies: {n*reps/dt/1e9:6.1f} GB/s")
