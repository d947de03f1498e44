#!/usr/bin/env python
"""Throughput of the file-to-file video path — `FrameInterpolator.interpolate_video(input.mp4, output.mp4, factor)` (what
reference main.py:118-129 calls) — on a synthetic 1080p clip, next to its stages measured alone: cv2/FFMPEG decode,
the GPU stage on decoded frames (BGR frames through the grey model = three planes per frame), cv2 `mp4v` encode.
The GPU stage is 100x a software codec, so the pipeline is codec-bound; this tool says by how much.

    python tools/bench_video.py [--frames 120 --gpus 1] > profiles/r02_video_path.json"""
import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import cv2
import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "ai-based-frame-interpolation_b200"))
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from model.inference import FrameInterpolator  # noqa: E402
from model.unet import FrameInterpolationUNet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=120)
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args()
    grey = bench.synthetic_frames(a.frames)[:, 0]
    with tempfile.TemporaryDirectory() as tmp:
        src, dst, ckpt = os.path.join(tmp, "in.mp4"), os.path.join(tmp, "out.mp4"), os.path.join(tmp, "m.pth")
        t0 = time.perf_counter()
        wr = cv2.VideoWriter(src, cv2.VideoWriter_fourcc(*"mp4v"), 30.0, (bench.W, bench.H), True)
        for f in grey:
            wr.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
        wr.release()
        encode_fps = a.frames / (time.perf_counter() - t0)
        t0 = time.perf_counter()
        cap, decoded = cv2.VideoCapture(src), []
        while True:
            ok, fr = cap.read()
            if not ok:
                break
            decoded.append(fr)
        decode_fps = len(decoded) / (time.perf_counter() - t0)
        torch.manual_seed(0)
        torch.save(FrameInterpolationUNet(bilinear=False).state_dict(), ckpt)
        fi = FrameInterpolator(ckpt, "cuda", gpus=a.gpus)
        fi.interpolate_sequence(decoded[:10], 2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        seq = fi.interpolate_sequence(decoded, 2)
        gpu_stage_fps = (len(seq) - len(decoded)) / (time.perf_counter() - t0)
        t0 = time.perf_counter()
        written = fi.interpolate_video(src, dst, 2)
        total = time.perf_counter() - t0
        fi.close()
    print(json.dumps({"clip": f"{a.frames} synthetic 1080p frames, mp4v, BGR", "gpus": a.gpus, "host_cores": os.cpu_count(),
                      "decode_frames_per_s": round(decode_fps, 1), "encode_frames_per_s": round(encode_fps, 1),
                      "gpu_stage_new_bgr_frames_per_s": round(gpu_stage_fps, 1),
                      "gpu_stage_note": "BGR frames through the grey model: 3 forwards per new frame",
                      "interpolate_video_seconds": round(total, 3), "frames_written": written,
                      "interpolate_video_output_frames_per_s": round(written / total, 1),
                      "interpolate_video_new_frames_per_s": round((written - a.frames) / total, 1)}))


if __name__ == "__main__":
    main()
