"""HBM bandwidth by access mix (torch ops, CUDA events, best of 10): the copy peak in MEASURED_PEAKS.json is a 1:1
read/write stream; a write-only stream is much slower on B200, which sets the roofline of the write-dominated kernels
(stem 2 B in / 128 B out per pixel, pack 2 / 8, bilinear upsample 1 / 4).   python tools/probe/bw_probe.py"""
import torch

dev = "cuda:0"
n = 1 << 30  # 1 Gi bf16 = 2 GB
a = torch.empty(n, dtype=torch.bfloat16, device=dev)
b = torch.empty_like(a)


def best_ms(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


print(f"copy   (1 read : 1 write): {4 * n / best_ms(lambda: b.copy_(a)) / 1e6:.0f} GB/s")
print(f"fill   (write only)      : {2 * n / best_ms(lambda: a.fill_(1.0)) / 1e6:.0f} GB/s")
print(f"memset (write only)      : {2 * n / best_ms(lambda: a.zero_()) / 1e6:.0f} GB/s")
