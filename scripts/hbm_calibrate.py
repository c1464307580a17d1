"""Calibration of this box's HBM: pure write, pure read and copy bandwidth with torch kernels (context for roofline fractions)."""
import torch
dev = torch.device("cuda:0")
n = 2 * 1024 ** 3 // 4
a = torch.empty(n, device=dev, dtype=torch.float32); b = torch.empty(n, device=dev, dtype=torch.float32)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timeit(lambda: a.zero_()); print("write  2 GiB: %.3f ms  %.0f GB/s" % (ms, n * 4 / ms / 1e6))
ms = timeit(lambda: a.fill_(1.5)); print("fill   2 GiB: %.3f ms  %.0f GB/s" % (ms, n * 4 / ms / 1e6))
ms = timeit(lambda: b.copy_(a)); print("copy   2+2 GiB: %.3f ms  %.0f GB/s" % (ms, 2 * n * 4 / ms / 1e6))
ms = timeit(lambda: a.sum()); print("read   2 GiB: %.3f ms  %.0f GB/s" % (ms, n * 4 / ms / 1e6))
