"""Training-step throughput on BASELINE config 5 (MSE + Adam, batch 16 at 256x256, FrameInterpolationUNet bilinear):
the B200 TrainStep against eager torch (fp32/TF32 and bf16 autocast) on the same GPU. One JSON line per arm.

    python tools/bench_train.py [--batch 16 --size 256 --steps 20 --warmup 5 --criterion mse|combined]
    torchrun --nproc-per-node N tools/bench_train.py ...     # data parallel: NCCL all-reduce of the flat gradient
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ai-based-frame-interpolation_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from model.train import CombinedLoss, TrainStep  # noqa: E402
from model.unet import FrameInterpolationUNet  # noqa: E402


def torch_step_factory(model, lr, amp, criterion):
    from test_gpu_train_step import ref_forward_train
    opt = torch.optim.Adam(model.parameters(), lr=lr)

    def step(f0, f1, gt):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            y = ref_forward_train(model, torch.cat([f0, f1], 1))
        loss = criterion(y.float(), gt)
        loss.backward()
        opt.step()
        return loss.detach()
    return step


def timed(fn, args, steps, warmup):
    for _ in range(warmup):
        fn(*args)
    torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = fn(*args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if dist.is_initialized():
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    return ms, float(loss)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--criterion", default="mse")
    ap.add_argument("--skip-torch", action="store_true")
    ap.add_argument("--graph", action="store_true", help="capture the step in a CUDA graph")
    ap.add_argument("--no-overlap", action="store_true", help="weight gradients on the main stream")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl")
    dev = torch.device("cuda", local)
    g = torch.Generator().manual_seed(local)
    f0, f1 = (torch.rand(a.batch, 1, a.size, a.size, generator=g).to(dev) for _ in range(2))
    gt = (f0 + f1) / 2
    rows = []
    torch.manual_seed(0)
    model = FrameInterpolationUNet(bilinear=True).to(dev).train()
    step = TrainStep(model, lr=1e-4, criterion=CombinedLoss() if a.criterion == "combined" else None, cuda_graph=a.graph, overlap_wgrad=not a.no_overlap)
    ms, loss = timed(step, (f0, f1, gt), a.steps, a.warmup)
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step(f0, f1, gt)                      # host time to enqueue one step (GPU idle at the start)
    host_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    rows.append({"arm": "b200_train_step" + ("_graph" if a.graph else ""), "ms_per_step": ms, "samples_per_s": world * a.batch * 1e3 / ms, "loss": loss,
                 "host_enqueue_ms": host_ms})
    if not a.skip_torch and world == 1:
        for name, amp, tf32 in (("torch_eager_fp32", False, False), ("torch_eager_tf32", False, True),
                                ("torch_eager_bf16_autocast", True, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.manual_seed(0)
            m = FrameInterpolationUNet(bilinear=True).to(dev).train()
            crit = CombinedLoss().to(dev) if a.criterion == "combined" else F.mse_loss
            ms, loss = timed(torch_step_factory(m, 1e-4, amp, crit), (f0, f1, gt), a.steps, a.warmup)
            rows.append({"arm": name, "ms_per_step": ms, "samples_per_s": a.batch * 1e3 / ms, "loss": loss})
    if local == 0:
        for r in rows:
            r.update({"workload": f"train step: FrameInterpolationUNet(bilinear) batch {a.batch} x {a.size}x{a.size}, "
                                  f"{a.criterion} + Adam", "n_gpus": world, "steps": a.steps, "warmup": a.warmup})
            print(json.dumps(r))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
