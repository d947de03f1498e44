"""GPU tier, needs >= 2 devices (gpurun --gpus 2): the data-parallel half of the training step — the only collective
of the whole path (BASELINE configs[4]: "with NCCL gradient allreduce"; reference loop model/train.py:183-199 run on N
replicas). Two NCCL ranks are spawned from the test.

What is checked, eager and CUDA-graph mode:
  * start-up sync: ranks seeded differently hold rank 0's parameters after TrainStep() (broadcast);
  * the two-bucket, overlapped all-reduce turns every element of the flat gradient into the MEAN over ranks — exactly:
    every all_reduce call is spied on (input cloned before, gathered from both ranks afterwards), the reduced slices tile
    the flat gradient with no gap and no overlap;
  * replicas stay bit-identical: after several steps on DIFFERENT batches the flat parameters and Adam moments are equal
    on both ranks;
  * fed the SAME batch, the two-rank gradient of the head (the layer nearest the loss: no ReLU-flip amplification) equals
    the one-rank gradient to bf16 noise, and so does the loss;
  * BatchNorm running estimates stay per replica (DDP semantics) until average_bn_buffers() equalises them.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _batch(seed, n=4, size=64):
    g = torch.Generator().manual_seed(seed)
    f0, f1 = torch.rand(n, 1, size, size, generator=g), torch.rand(n, 1, size, size, generator=g)
    return f0, f1, (f0 + f1) / 2


def _worker(rank, world, port, graph, out_dir):
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    for p in (str(root / "ai-based-frame-interpolation_b200"), str(root)):
        if p not in sys.path:
            sys.path.insert(0, p)
    from model.train import TrainStep
    from model.unet import FrameInterpolationUNet
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)

    # ---- one-rank baseline on the shared batch, before any process group exists
    torch.manual_seed(0)
    solo = FrameInterpolationUNet(bilinear=True).to(dev).train()
    solo_step = TrainStep(solo, lr=1e-3)
    shared = [t.to(dev) for t in _batch(100)]
    solo_loss = solo_step(*shared).item()
    solo_head = solo_step.grad_view[solo.unet.outc.conv.weight].clone()

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        torch.manual_seed(rank * 17)                       # replicas deliberately start from DIFFERENT weights
        model = FrameInterpolationUNet(bilinear=True).to(dev).train()
        step = TrainStep(model, lr=1e-3, cuda_graph=graph)
        n = step.flat_param.numel()
        gathered = [torch.empty_like(step.flat_param) for _ in range(world)]
        dist.all_gather(gathered, step.flat_param)
        assert torch.equal(gathered[0], gathered[1]), "parameters were not broadcast from rank 0"
        torch.manual_seed(0)
        want = torch.cat([p.detach().reshape(-1) for p in FrameInterpolationUNet(bilinear=True).parameters()])
        assert torch.equal(step.flat_param.cpu(), want), "rank 0's parameters are the seed-0 initialisation"

        # ---- same batch on both ranks: loss and head gradient as on one rank
        loss = step(*shared).item()
        assert abs(loss - solo_loss) <= 1e-2 * abs(solo_loss) + 1e-6, (loss, solo_loss)
        head = step.grad_view[model.unet.outc.conv.weight]
        rel = ((head - solo_head).norm() / solo_head.norm()).item()
        assert rel < 5e-2, f"head gradient differs from the one-rank step: rel L2 {rel:.4f}"

        # ---- different batches: spy on every all-reduce of the following steps
        calls = []
        real = dist.all_reduce

        def spy(t, op=dist.ReduceOp.SUM, group=None, async_op=False):
            calls.append((t, t.clone(), op))
            return real(t, op=op, group=group, async_op=async_op)

        dist.all_reduce = spy
        own = [t.to(dev) for t in _batch(200 + rank)]
        for _ in range(4 if graph else 2):                 # graph mode: two eager warm-up steps, capture, replay
            step(*own)
        dist.all_reduce = real
        torch.cuda.synchronize()
        per_step = 2
        assert len(calls) % per_step == 0 and len(calls) >= per_step
        last = calls[-per_step:]
        base = step.flat_grad.data_ptr()
        covered = sorted(((t.data_ptr() - base) // 4, t.numel()) for t, _, _ in last)
        assert covered[0][0] == 0 and covered[0][1] == covered[1][0] and covered[1][0] + covered[1][1] == n, covered
        for t, before, op in last:
            assert op == dist.ReduceOp.AVG
            parts = [torch.empty_like(before) for _ in range(world)]
            dist.all_gather(parts, before)
            mean = (parts[0] + parts[1]) / 2
            assert not torch.equal(parts[0], parts[1]), "the ranks worked on different batches"
            assert torch.allclose(t, mean, rtol=1e-6, atol=1e-12), "all-reduce result is not the mean gradient"

        # ---- replicas are bit-identical after the optimizer steps
        for name, buf in (("parameters", step.flat_param), ("exp_avg", step.m), ("exp_avg_sq", step.v)):
            parts = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(parts, buf)
            assert torch.equal(parts[0], parts[1]), f"{name} diverged between the replicas"
        assert not torch.equal(step.flat_param.cpu(), want), "the optimizer moved the parameters"

        # ---- BatchNorm running estimates: per replica until averaged
        bn = model.unet.inc.double_conv[1]
        parts = [torch.empty_like(bn.running_mean) for _ in range(world)]
        dist.all_gather(parts, bn.running_mean)
        assert not torch.equal(parts[0], parts[1])
        mean_before = (parts[0] + parts[1]) / 2
        step.average_bn_buffers()
        dist.all_gather(parts, bn.running_mean)
        assert torch.equal(parts[0], parts[1]) and torch.allclose(parts[0], mean_before, rtol=1e-6, atol=1e-9)
        (Path(out_dir) / f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("graph", [False, True])
def test_two_rank_training_step(graph, tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, graph, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
