#!/usr/bin/env python
"""CLI with the reference main.py's sub-commands and flags (reference main.py:40-72): train / infer / video / serve /
info. `infer` and `video` drive model.inference.FrameInterpolator, `train` drives model.train.main (the B200
training step)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def build_parser():
    p = argparse.ArgumentParser(description="AI-Based Frame Interpolation (B200-native path)")
    sub = p.add_subparsers(dest="command", help="Available commands")
    t = sub.add_parser("train", help="Train the model")
    t.add_argument("--data-dir", required=True)
    t.add_argument("--epochs", type=int, default=100)
    t.add_argument("--batch-size", type=int, default=8)
    t.add_argument("--lr", type=float, default=1e-4)
    t.add_argument("--device", default="auto")
    i = sub.add_parser("infer", help="Interpolate one frame pair")
    i.add_argument("--model", required=True)
    i.add_argument("--frame1", required=True)
    i.add_argument("--frame2", required=True)
    i.add_argument("--output", default="interpolated.png")
    i.add_argument("--device", default="auto")
    v = sub.add_parser("video", help="Interpolate a video")
    v.add_argument("--model", required=True)
    v.add_argument("--input", required=True)
    v.add_argument("--output", required=True)
    v.add_argument("--factor", type=int, default=2)
    v.add_argument("--device", default="auto")
    s = sub.add_parser("serve", help="Start the HTTP API")
    s.add_argument("--host", default="0.0.0.0")
    s.add_argument("--port", type=int, default=8000)
    s.add_argument("--reload", action="store_true")
    sub.add_parser("info", help="Show model information")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    if not args.command:
        build_parser().print_help()
        return 0
    device = getattr(args, "device", "auto")  # `serve` / `info` have no --device (the reference crashes here)
    device = "cuda" if device == "auto" else device
    try:
        if args.command == "train":
            from model.train import main as train_main
            train_main(["--data-dir", args.data_dir, "--epochs", str(args.epochs), "--batch-size", str(args.batch_size),
                        "--lr", str(args.lr), "--device", args.device])
            return 0
        if args.command == "infer":
            import cv2
            from model.inference import FrameInterpolator
            interpolator = FrameInterpolator(args.model, device)
            frame1, frame2 = cv2.imread(args.frame1), cv2.imread(args.frame2)
            if frame1 is None or frame2 is None:
                print("Error: Could not read input frames")
                return 1
            print("Generating intermediate frame...")
            cv2.imwrite(args.output, interpolator.interpolate_frames(frame1, frame2))
            print(f"Interpolated frame saved to: {args.output}")
        elif args.command == "video":
            from model.inference import FrameInterpolator
            interpolator = FrameInterpolator(args.model, device)
            print(f"Interpolating video: {args.input}\nOutput: {args.output}\nFactor: {args.factor}x")
            n = interpolator.interpolate_video(args.input, args.output, args.factor)
            print(f"Video interpolation completed! ({n} frames written)")
        elif args.command == "serve":
            import uvicorn
            print(f"Starting API server on {args.host}:{args.port}")
            uvicorn.run("api.app:app", host=args.host, port=args.port, reload=args.reload)
        elif args.command == "info":
            from model.unet import FrameInterpolationUNet, count_parameters
            for bilinear in (False, True):
                m = FrameInterpolationUNet(bilinear=bilinear)
                print(f"FrameInterpolationUNet(bilinear={bilinear}): {count_parameters(m):,} parameters")
            print("Input: two grayscale frames [B,1,H,W]; output: one intermediate frame [B,1,H,W]")
    except ImportError as e:
        print(f"Import error: {e}")
        return 1
    except Exception as e:
        print(f"Error: {e}")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
