#!/usr/bin/env python
"""Per-launch device times of one forward at a small frame size (default: ONE 256x256 pair, the shape every
POST /interpolate produces), to see which layers the fixed per-kernel cost and the low tile counts hurt.

    python tools/profile_small.py [--n 1 --h 256 --w 256 --bilinear] > profiles/r02_small_profile.json

CUDA events around every launch (serialised: no programmatic-dependent-launch overlap), 50 profiled forwards."""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "ai-based-frame-interpolation_b200"))
from model import _engine as E  # noqa: E402
from model.unet import FrameInterpolationUNet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1)
    ap.add_argument("--h", type=int, default=256)
    ap.add_argument("--w", type=int, default=256)
    ap.add_argument("--bilinear", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net = E.Net(dev, 2, 1, a.bilinear)
    net.load_state_dict(FrameInterpolationUNet(bilinear=a.bilinear).state_dict())
    f = torch.randint(0, 256, (a.n + 1, 1, a.h, a.w), dtype=torch.uint8, device=dev)
    run = lambda: net.forward(f[:-1], f[1:], want_f32=False, want_u8=True)  # noqa: E731
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        run()
    e1.record()
    torch.cuda.synchronize()
    back_to_back = e0.elapsed_time(e1) / 200
    net.set_profiling(True)
    for _ in range(50):
        run()
    prof = net.profile()
    net.set_profiling(False)
    rows = [{"launch": p["name"], "us": round(1e3 * p["ms_total"] / p["calls"], 2),
             "tflops": round(p["flops"] / (p["ms_total"] / p["calls"]) / 1e9, 1)} for p in prof]
    print(json.dumps({"shape": [a.n, a.h, a.w], "bilinear": a.bilinear, "forward_ms_back_to_back": round(back_to_back, 4),
                      "sum_of_serialised_launches_ms": round(sum(r["us"] for r in rows) / 1e3, 4), "launches": rows}, indent=1))


if __name__ == "__main__":
    main()
