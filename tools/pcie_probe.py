import torch, time
n = 691_000_000
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for _ in range(3):
    h.copy_(d, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print("D2H %.2f ms per 691 MB = %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
h2 = torch.empty(n // 2, dtype=torch.uint8, pin_memory=True); h3 = torch.empty(n // 2, dtype=torch.uint8, pin_memory=True)
t0 = time.perf_counter()
for _ in range(10):
    h2.copy_(d[: n // 2], non_blocking=True); h3.copy_(d[n // 2 : 2 * (n // 2)], non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print("2 x D2H %.2f ms = %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
t0 = time.perf_counter()
for _ in range(10):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print("H2D %.2f ms = %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
