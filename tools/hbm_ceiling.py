"""Measure what a pure streaming kernel reaches on this GPU: copy (read+write, the MEASURED_PEAKS.json definition) and
write-only (the pattern of the evaluation pass).  Context for roofline fractions; not a bench value."""
import json
import torch

n = 2 * 1024 ** 3  # 16 GiB of fp64
a = torch.empty(n // 2, dtype=torch.float64, device="cuda")
b = torch.empty(n // 2, dtype=torch.float64, device="cuda")
big = torch.empty(n, dtype=torch.float64, device="cuda")


def timeit(f, reps=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


t_copy = timeit(lambda: b.copy_(a))
t_fill = timeit(lambda: big.fill_(1.5))
t_zero = timeit(lambda: big.zero_())
print(json.dumps({"copy_GBs": 2 * a.numel() * 8 / t_copy / 1e6, "fill_GBs": big.numel() * 8 / t_fill / 1e6,
                  "memset_GBs": big.numel() * 8 / t_zero / 1e6}))

# PCIe: pinned host <-> device, 2 GiB
h = torch.empty(256 * 1024 ** 2, dtype=torch.float64).pin_memory()
d = torch.empty_like(h, device="cuda")
t_h2d = timeit(lambda: d.copy_(h, non_blocking=True), reps=3)
t_d2h = timeit(lambda: h.copy_(d, non_blocking=True), reps=3)
print(json.dumps({"h2d_GBs": h.numel() * 8 / t_h2d / 1e6, "d2h_GBs": h.numel() * 8 / t_d2h / 1e6}))
