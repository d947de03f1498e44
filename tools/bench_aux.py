#!/usr/bin/env python
"""Throughput of the non-GEMM kernels of the path (through the C ABI) against the measured HBM copy peak:
pack_pair, head_post, bilinear upsample, stem conv, SSIM+PSNR. Writes one JSON object per kernel.

    python tools/bench_aux.py > profiles/r01_aux_kernels.jsonl       (on the GPU box)

Timing: CUDA events on the launching stream, 5 warm-up + 20 timed launches; every working set is > 126 MB (L2)."""
import ctypes as C
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "ai-based-frame-interpolation_b200"))
from model import _engine as E  # noqa: E402

PEAK = 6535.7
pk = ROOT / "MEASURED_PEAKS.json"
if pk.exists():
    PEAK = json.loads(pk.read_text())["hbm_gbs"]


def timed(fn, iters=20, warm=5):
    import os
    iters = int(os.environ.get("FI_AUX_ITERS", iters))   # ncu captures: FI_AUX_ITERS=1 FI_AUX_WARM=0 (one launch each)
    warm = int(os.environ.get("FI_AUX_WARM", warm))
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def report(name, ms, algo_bytes, note):
    gbs = algo_bytes / ms / 1e6
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "algorithmic_bytes": algo_bytes, "achieved_gbs": round(gbs, 1),
                      "hbm_peak_gbs": PEAK, "frac_of_hbm_peak": round(gbs / PEAK, 3), "note": note}))


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    st = E.current_stream()
    H, W = 1080, 1920
    n = 32
    f1 = torch.randint(0, 256, (n, 1, H, W), dtype=torch.uint8, device=dev)
    f2 = torch.randint(0, 256, (n, 1, H, W), dtype=torch.uint8, device=dev)
    out = torch.empty((n, 2, H, W), dtype=torch.float32, device=dev)
    ms = timed(lambda: E.check(E.lib().fiPackPairU8(f1.data_ptr(), f2.data_ptr(), out.data_ptr(), n, 1, H, W, st)))
    report("pack_pair_u8", ms, n * H * W * (2 + 8), f"{n} 1080p grey pairs: 2 B in + 8 B out per pixel")

    y = torch.randn((n * 2, 1, H, W), device=dev)
    o8 = torch.empty(y.shape, dtype=torch.uint8, device=dev)
    ms = timed(lambda: E.check(E.lib().fiHeadPostU8(y.data_ptr(), o8.data_ptr(), y.numel(), st)))
    report("head_post_u8", ms, y.numel() * 5, f"{2 * n} 1080p fp32 frames -> u8: 4 B in + 1 B out per value")

    c, h, w, nb = 64, 540, 960, 8
    src = torch.randn((nb, h, w, c), device=dev).to(torch.bfloat16)
    dst = torch.empty((nb, 2 * h, 2 * w, c), dtype=torch.bfloat16, device=dev)
    ms = timed(lambda: E.check(E.lib().fiUpsample2x(src.data_ptr(), dst.data_ptr(), nb, h, w, c, st)))
    report("upsample2x_bilinear", ms, nb * h * w * c * 2 * 5, f"{nb}x{h}x{w}x{c} bf16 -> x2: 1 read + 4 writes per element")

    kp = E.lib().fiStemPackedK(2)
    wt = (torch.randn(64, 2, 3, 3) * 0.3).contiguous()
    packed = torch.empty((64, kp), dtype=torch.int16)
    E.check(E.lib().fiStemPackWeights(wt.data_ptr(), 2, packed.data_ptr()))
    wk, bias = packed.to(dev), torch.zeros(64, device=dev)
    ns = 8
    d64 = torch.empty((ns, H, W, 64), dtype=torch.bfloat16, device=dev)
    p0, p1 = E.planes_of(f1[:ns]), E.planes_of(f2[:ns])
    ms = timed(lambda: E.check(E.lib().fiStemConv(C.byref(p0), C.byref(p1), 1, wk.data_ptr(), bias.data_ptr(),
                                                  d64.data_ptr(), ns, H, W, st)))
    report("stem_conv (tcgen05, hi/lo split)", ms, ns * H * W * (2 + 128),
           f"{ns} 1080p u8 pairs -> 64ch bf16 NHWC: 2 B in + 128 B out per pixel")

    for (hh, ww, nn, tag) in ((1080, 1920, 64, "1080p"), (2160, 3840, 32, "4K")):
        a = torch.randint(0, 256, (nn, hh, ww), dtype=torch.uint8, device=dev)
        b = (a.to(torch.int16) + torch.randint(-9, 10, a.shape, device=dev, dtype=torch.int16)).clamp(0, 255).to(torch.uint8)
        ws = torch.empty(max(16, E.lib().fiSsimPsnrWorkspaceBytes(nn, hh, ww)), dtype=torch.uint8, device=dev)
        res = torch.empty((nn, 2), dtype=torch.float64, device=dev)
        ms = timed(lambda: E.check(E.lib().fiSsimPsnrU8(a.data_ptr(), b.data_ptr(), nn, hh, ww, res.data_ptr(),
                                                        ws.data_ptr(), st)))
        report(f"ssim_psnr_u8 {tag}", ms, nn * hh * ww * 2,
               f"{nn} {tag} u8 pairs, fused SSIM+PSNR: 2 B in per pixel (ALU-bound kernel: ~60 integer/fp ops per pixel)")


if __name__ == "__main__":
    main()
