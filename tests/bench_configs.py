#!/usr/bin/env python
"""(Lives under tests/ because it runs the oracle's torch ops as the library baseline.)
Side measurements for the BASELINE.json configs that are not the headline bench line (on the GPU box):

  config[1]  UNet bf16 forward, batch 32 of 256x256 pairs on one B200 — this path vs the same network run by stock
             PyTorch on the same GPU (the oracle's functional ops = the reference module's ops: eager fp32 with TF32
             convs, and bf16 channels_last autocast), i.e. the "library baseline" of SURVEY.md §2.1.
  latency    batch 1 at 256x256 (what one `POST /interpolate` request costs) and at 1080p.
  config[3]  4K (3840x2160), pairs + fused SSIM/PSNR of the result against a synthetic ground truth.
  colour     the 6-in / 3-out UNet of the README (UNet(6, 3), colour frame pairs) at 1080p, u8 planes in, u8 frames out,
             and the bilinear (Upsample) decoder variant of the grey network.

    python tests/bench_configs.py > profiles/r01_configs.jsonl"""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "ai-based-frame-interpolation_b200"))
sys.path.insert(0, str(ROOT))
from model import _engine as E  # noqa: E402
from oracle import unet_oracle as O  # noqa: E402


def timed(fn, iters, warm=3):
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


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    sd = O.init_state_dict(0, 2, 1, False)
    net = E.Net(dev, 2, 1, False)
    net.load_state_dict(sd)
    sd_gpu = {k: v.to(dev) for k, v in sd.items()}

    def ours(f1, f2):
        return lambda: net.forward(f1, f2, want_f32=True)[0]

    # ---- config[1]: batch 32 x 256^2
    n, h, w = 32, 256, 256
    x1 = torch.rand(n, 1, h, w, device=dev) * 2 - 1
    x2 = torch.rand(n, 1, h, w, device=dev) * 2 - 1
    flops = O.flops_per_forward(n, h, w)
    ms = timed(ours(x1, x2), 20)
    x = torch.cat([x1, x2], 1)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    ms_tf32 = timed(lambda: O.unet_forward(sd_gpu, x), 10)
    xcl = x.contiguous(memory_format=torch.channels_last)

    def bf16_eager():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return O.unet_forward(sd_gpu, xcl)
    ms_bf16 = timed(bf16_eager, 10)
    ref = O.unet_forward(sd_gpu, x)
    got = net.forward(x1, x2)[0]
    print(json.dumps({"config": "UNet bf16 forward, batch 32 of 256x256 pairs, 1xB200", "gflop": flops / 1e9,
                      "this_path": {"ms": round(ms, 3), "pairs_per_s": round(n / ms * 1e3, 1),
                                    "tflops": round(flops / ms / 1e9, 1)},
                      "torch_eager_fp32_tf32_nchw": {"ms": round(ms_tf32, 3), "pairs_per_s": round(n / ms_tf32 * 1e3, 1)},
                      "torch_eager_bf16_channels_last": {"ms": round(ms_bf16, 3),
                                                         "pairs_per_s": round(n / ms_bf16 * 1e3, 1)},
                      "speedup_vs_torch_bf16": round(ms_bf16 / ms, 2), "speedup_vs_torch_tf32": round(ms_tf32 / ms, 2),
                      "max_abs_vs_torch_fp32_pixel_units": float((got - ref).abs().max()) / 2}))

    # ---- latency, batch 1
    for (hh, ww) in ((256, 256), (1080, 1920)):
        a = torch.rand(1, 1, hh, ww, device=dev) * 2 - 1
        b = torch.rand(1, 1, hh, ww, device=dev) * 2 - 1
        fn = ours(a, b)
        ms_dev = timed(fn, 50)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            fn()
            torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 50 * 1e3
        print(json.dumps({"config": f"latency, 1 pair {hh}x{ww}", "device_ms_back_to_back": round(ms_dev, 4),
                          "wall_ms_per_synchronous_call": round(wall, 4),
                          "tflops": round(O.flops_per_forward(1, hh, ww) / ms_dev / 1e9, 1)}))

    # ---- config[3]: 4K with SSIM/PSNR evaluation
    hh, ww, nn = 2160, 3840, 2
    f = torch.randint(0, 256, (nn + 1, 1, hh, ww), dtype=torch.uint8, device=dev)
    gt = torch.randint(0, 256, (nn, hh, ww), dtype=torch.uint8, device=dev)

    def step4k():
        out = net.forward(f[:nn], f[1:], want_f32=False, want_u8=True)[1]
        return E.ssim_psnr_u8(out[:, 0], gt)
    ms4k = timed(step4k, 10)
    ms_metric = timed(lambda: E.ssim_psnr_u8(gt, gt), 20)
    print(json.dumps({"config": "4K (3840x2160) interpolation + SSIM/PSNR evaluation, 2 pairs per forward",
                      "ms_per_step": round(ms4k, 3), "frames_per_s": round(nn / ms4k * 1e3, 1),
                      "tflops": round(O.flops_per_forward(nn, hh, ww) / ms4k / 1e9, 1),
                      "ssim_psnr_ms_per_4k_pair": round(ms_metric / nn, 4)}))

    # ---- colour UNet(6,3) and the bilinear decoder at 1080p, 4 pairs per forward
    del net
    torch.cuda.empty_cache()
    hh, ww, nn = 1080, 1920, 4
    for name, (cin, cout, bil) in (("UNet(6,3) colour pairs, ConvT decoder", (6, 3, False)),
                                   ("UNet(2,1) grey pairs, bilinear decoder", (2, 1, True))):
        m = E.Net(dev, cin, cout, bil)
        m.load_state_dict(O.init_state_dict(0, cin, cout, bil))
        c = cin // 2
        fa = torch.randint(0, 256, (nn, c, hh, ww), dtype=torch.uint8, device=dev)
        fb = torch.randint(0, 256, (nn, c, hh, ww), dtype=torch.uint8, device=dev)
        ms_c = timed(lambda: m.forward(fa, fb, want_f32=False, want_u8=True)[1], 20)
        fl = O.flops_per_forward(nn, hh, ww, cin, cout, bil)
        print(json.dumps({"config": f"1080p, {name}, {nn} pairs per forward, u8 in / u8 out",
                          "ms_per_step": round(ms_c, 3), "frames_per_s": round(nn / ms_c * 1e3, 1),
                          "tflops": round(fl / ms_c / 1e9, 1)}))
        del m
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
