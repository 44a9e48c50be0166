"""Dev helper: what HBM delivers for streaming read/write mixes (plumbing-level torch kernels, not part of the product).
The per-column leaf builder moves 1.29 GB in and 0.50 GB out per launch (72/28); MEASURED_PEAKS.json's copy figure is 50/50."""
import torch

n = 128 * 1024 * 1024  # 512 MiB per fp32 tensor, far beyond L2
a, b, c, o = (torch.empty(n, dtype=torch.float32, device="cuda").normal_() for _ in range(4))


def time_ms(fn, reps=10):
    for _ in range(3):
        fn()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


cases = [("read only (sum)", lambda: torch.sum(a), 4 * n), ("copy 50/50", lambda: o.copy_(a), 8 * n),
         ("add 67/33", lambda: torch.add(a, b, out=o), 12 * n), ("addcmul 75/25", lambda: torch.addcmul(a, b, c, out=o), 16 * n)]
for name, fn, nbytes in cases:
    ms = time_ms(fn)
    print("%-18s %7.3f ms  %7.1f GB/s" % (name, ms, nbytes / ms / 1e6))
