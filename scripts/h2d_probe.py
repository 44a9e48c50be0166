"""Dev helper: host->device copy rate of pinned memory through torch and through the C ABI."""
import sys, time
sys.path.insert(0, ".")
import torch
import cpvs_b200
n = 16384
host = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
host.fill_(0.5)
dev = torch.empty((n, n), dtype=torch.float32, device="cuda")
for _ in range(2):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 5
print("torch pinned H2D: %.2f ms  %.1f GB/s" % (dt * 1e3, n * n * 4 / dt / 1e9))
ctx = cpvs_b200.Context(0)
a = host.numpy()
for i in range(4):
    t = time.perf_counter()
    mm = cpvs_b200.MinMaxHierarchy(a, ctx)
    ctx.synchronize()
    dt = time.perf_counter() - t
    print("C ABI minmax from pinned host: %.2f ms (%s)" % (dt * 1e3, mm.timing()))
    mm.close()
