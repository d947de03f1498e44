#!/usr/bin/env python
"""Headline benchmark: interpolated frames/s of the UNet forward on 1080p 2x video interpolation (BASELINE.json
configs[2]: 600 synthetic 1920x1080 frames, frame pairs sharded across the GPUs of one node, no collective).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU algorithm (oracle port) on the host cores

A step = one forward of `--pairs` consecutive frame pairs (u8 frames resident in HBM -> u8 interpolated frames in HBM).
`value` is device-timed (CUDA events, max over ranks); `e2e` goes through the C-ABI host-buffer entry point
(fiNetInterpolateClipHostU8: pinned H2D of every frame, forward, D2H of every result, all inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT / "ai-based-frame-interpolation_b200"), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "interpolated frames/sec, 1080p UNet fwd"
H, W = 1080, 1920
N_FRAMES = 600


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        load = [v for v in sm if mx and v > 0.3 * mx] or sm
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_frames(n, first=0):
    """Frames [first, first+n) of the synthetic clip: a moving bright disc over a gradient + noise (the reference's
    only data generator is of this kind, demo_simple.py:17-40), seeded per frame index."""
    import numpy as np
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    base = (xx / W * 96 + yy / H * 64).astype(np.float32)
    out = np.empty((n, 1, H, W), dtype=np.uint8)
    for k in range(n):
        i = first + k
        rs = np.random.RandomState(1000 + i)
        cx, cy = 200 + 2.5 * i, 540 + 120 * np.sin(i / 7.0)
        disc = ((xx - cx) ** 2 + (yy - cy) ** 2 < 90 ** 2) * 120.0
        out[k, 0] = np.clip(base + disc + rs.randint(0, 24, size=(H, W)), 0, 255).astype(np.uint8)
    return out


def cpu_forward_seconds(pairs, threads):
    """The oracle (CPU port of reference model/unet.py, fp32) on `pairs` 1080p frame pairs; returns seconds."""
    import torch
    from oracle import unet_oracle as O
    torch.set_num_threads(threads)
    sd = O.init_state_dict(0, 2, 1, False)
    fr = synthetic_frames(pairs + 1)
    x = torch.cat([O.preprocess_u8(fr[:-1]), O.preprocess_u8(fr[1:])], 1)
    O.unet_forward(sd, x[:1, :, :64, :64])  # thread-pool / allocator warm-up on a tiny crop
    t0 = time.perf_counter()
    for i in range(pairs):
        O.postprocess(O.unet_forward(sd, x[i:i + 1]))
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    import torch
    from oracle import unet_oracle as O
    torch.set_num_threads(threads)
    sd = O.init_state_dict(0, 2, 1, False)
    fr = synthetic_frames(2)
    x = torch.cat([O.preprocess_u8(fr[:1]), O.preprocess_u8(fr[1:])], 1)
    O.unet_forward(sd, x[:, :, :64, :64])
    budget_s, times = 200.0, []
    steps = args.steps
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        O.postprocess(O.unet_forward(sd, x))
        dt = time.perf_counter() - t0
        if i >= args.warmup or args.warmup == 0:
            times.append(dt)
        if i == 0 and args.warmup > 0:
            # a 1080p CPU forward takes seconds: keep the whole run inside a few minutes
            affordable = max(1, int(budget_s / dt) - 1)
            if args.warmup + args.steps > affordable + 1:
                steps = max(1, affordable)
                args.warmup, args.steps = 1, steps
        if len(times) >= steps:
            break
    total = sum(times)
    fps = len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "1080p (1920x1080) 2x video interpolation, UNet(2,1,bilinear=False) random-init",
                   "pairs_per_step": 1},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{len(times)} timed 1080p frame pairs, 1 per step, oracle/unet_oracle.py "
                                   "(torch fp32 CPU restatement of reference model/unet.py; the Python reference "
                                   "cannot travel to the GPU box)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=4, help="frame pairs per forward (per GPU)")
    ap.add_argument("--bilinear", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="bf16 (headline) or the hi/lo-split fp32-grade path (3x the tensor work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default=None, help="write the per-launch table (JSON) here")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    from model import _engine as E
    from model.unet import UNet

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = args.pairs
    net = E.Net(dev, 2, 1, args.bilinear, args.precision)
    torch.manual_seed(0)   # random-init weights of the architecture (PyTorch default init, as the reference's modules)
    net.load_state_dict(UNet(2, 1, args.bilinear).state_dict())

    # this rank's contiguous shard of the 599 frame pairs (neighbouring ranks share one boundary frame); the timed
    # steps rotate over a window of the shard (generating all 600 1080p frames on the host would only slow start-up)
    from model.sharding import shard_pairs
    first_pair, n_pairs = shard_pairs(N_FRAMES, world, rank)
    n_local = max(B + 1, min(B * 4 + 1, n_pairs + 1))
    host = synthetic_frames(n_local, first=first_pair)
    frames = torch.from_numpy(host).to(dev)

    def step(i):
        s = (i * B) % (n_local - B)
        return net.forward(frames[s:s + B], frames[s + 1:s + B + 1], want_f32=False, want_u8=True)[1]

    flops_step, launches_step = net.cost(B, H, W)
    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    net.set_profiling(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        out = step(i)
    ev1.record()
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if rank == 0 else None
    prof = net.profile()
    net.set_profiling(False)
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end through the host-buffer C-ABI call: one clip of steps*B frame pairs on the HOST in, the
    # interpolated frames on the HOST out (pinned H2D of every frame, forward, D2H of every result inside the region)
    e2e_steps = args.steps
    idx = np.arange(e2e_steps * B + 1) % n_local
    clip = np.ascontiguousarray(host[idx])
    net.interpolate_clip_host_u8(clip[:2 * B + 1], B)  # staging buffers + streams allocated outside the timed region
    barrier()
    t0 = time.perf_counter()
    res = net.interpolate_clip_host_u8(clip, B)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * B * e2e_steps / e2e_s

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    conv = [p for p in prof if p["kind"] == 1]
    conv_ms = sum(p["ms_total"] for p in conv)
    conv_flops = sum(p["flops"] * p["calls"] for p in conv)
    conv_launches = sum(p["calls"] for p in conv)
    achieved = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])  # kernels are timed inside a long step
    all_ms = sum(p["ms_total"] for p in prof)
    traffic = None  # DRAM bytes per conv launch from the committed ncu --set full capture, scaled to this batch
    tf = ROOT / "profiles" / "ncu_traffic.json"
    if tf.exists() and not args.bilinear and args.precision == "bf16":
        t = json.loads(tf.read_text())
        traffic = t["dram_bytes"] / t["pairs"] * B / t["conv_launches"]
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak else None, "traffic": traffic,
                "traffic_note": "avg DRAM read+write bytes per tcgen05 conv launch: profiles/ncu_traffic.json "
                                "(ncu --set full at 1 pair) x pairs_per_step; algorithmic bytes per launch avg = %.3e"
                                % (sum(p["bytes"] for p in conv) / max(1, len(conv))),
                "kernel": "conv_gemm_kernel (tcgen05 implicit GEMM), all instantiations",
                "peak_source": pk_src + ", sustained bf16 cuBLAS figure",
                "avg_launch_ms": conv_ms / max(1, conv_launches), "launches": conv_launches,
                "flops_per_launch_avg": conv_flops / max(1, conv_launches),
                "share_of_step": conv_ms / all_ms if all_ms else None,
                "whole_step_tflops": flops_step * args.steps / (ms / 1e3) / 1e12}
    if args.profile_out:
        Path(args.profile_out).write_text(json.dumps(prof, indent=1))

    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        secs = cpu_forward_seconds(2, threads)
        cpu = {"value": 2 / secs, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "2 of the 599 1080p frame pairs (1 per forward) through oracle/unet_oracle.py, the fp32 torch "
                         "CPU restatement of reference model/unet.py"}

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (hi/lo split, fp32-grade)",
        "data": "synthetic",
        "config": {"workload": "1080p (1920x1080) 2x video interpolation, 600 synthetic frames, "
                               "UNet(2,1,bilinear=%s) random-init, frame pairs sharded across GPUs" % args.bilinear,
                   "pairs_per_step": B, "frame": [H, W],
                   "l2": "inputs larger than L2: %.1f GB of activations written and re-read per step vs 126 MB L2; "
                         "frame window rotates every step" % (2.3 * B)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": (B + 1) * H * W,
                "d2h_bytes_per_step": B * H * W, "steps": e2e_steps,
                "api": "fiNetInterpolateClipHostU8: host u8 clip -> host u8 interpolated frames, one synchronous "
                       "call over steps*pairs_per_step pairs; copies overlap compute inside the library"},
        "gpu_launches": launches_step * args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "flops_per_step": flops_step,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
