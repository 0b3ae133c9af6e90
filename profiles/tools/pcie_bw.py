"""Pinned-memory PCIe bandwidth of the box (the bound of bench.py's e2e leg, which returns every observation row to the host):
D2H / H2D in the chunk sizes the host-buffer rollout uses (8.4 MB of observations per env step) and in one piece."""
import time

import torch

dev = torch.device('cuda')
for mb in (8.4, 64, 1024):
    n = int(mb * 1e6) // 4
    d = torch.empty(n, device=dev)
    h = torch.empty(n, pin_memory=True)
    reps = max(4, int(2e9 / (4 * n)))
    for name, (src, dst) in (('D2H', (d, h)), ('H2D', (h, d))):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f'{name} {mb:7.1f} MB x {reps}: {4 * n * reps / dt / 1e9:6.1f} GB/s', flush=True)
# both directions at once (copy engines are independent)
n = int(64e6) // 4
d1, h1, d2, h2 = torch.empty(n, device=dev), torch.empty(n, pin_memory=True), torch.empty(n, device=dev), torch.empty(n, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(16):
    with torch.cuda.stream(s1):
        h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f'D2H + H2D concurrently: {4 * n * 16 / dt / 1e9:6.1f} GB/s each way', flush=True)
