"""Pinned host -> device copy bandwidth on this box for the bench's step size (upper bound of the e2e arm)."""
import time
import torch

for mb in (16, 64, 184, 512):
    n = mb * 1024 * 1024
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for streams in (1, 2, 4):
        ss = [torch.cuda.Stream() for _ in range(streams)]
        chunk = n // streams
        torch.cuda.synchronize()
        best = 0.0
        for rep in range(5):
            t0 = time.perf_counter()
            for k, s in enumerate(ss):
                with torch.cuda.stream(s):
                    d[k * chunk:(k + 1) * chunk].copy_(h[k * chunk:(k + 1) * chunk], non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = max(best, n / dt / 1e9)
        print(f"H2D {mb:4d} MB, {streams} stream(s): {best:6.1f} GB/s")
h = torch.empty(184 * 1024 * 1024, dtype=torch.uint8).pin_memory()
d = torch.empty_like(h, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize()
print(f"H2D 184 MB by CUDA events: {h.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s")

# sustained: 10 copies of 184.56 MB alternating over two streams and two device buffers (the bench's e2e pattern)
n = 184563712
h = torch.empty(n, dtype=torch.uint8).pin_memory()
ds = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
ss = [torch.cuda.Stream() for _ in range(2)]
for trial in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(10):
        with torch.cuda.stream(ss[k % 2]):
            ds[k % 2].copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"sustained 10 x 184.6 MB on 2 streams: {10 * n / dt / 1e9:.1f} GB/s ({dt * 100:.2f} ms per copy)")
# same with a compute kernel running concurrently on a third stream (HBM traffic)
big = torch.empty(256 * 1024 * 1024, dtype=torch.float32, device="cuda")
s3 = torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.stream(s3):
    for _ in range(60):
        big.mul_(1.0001)
for k in range(10):
    with torch.cuda.stream(ss[k % 2]):
        ds[k % 2].copy_(h, non_blocking=True)
for s in ss:
    s.synchronize()
dt = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"sustained with concurrent HBM-bound kernels: {10 * n / dt / 1e9:.1f} GB/s")
