"""Training step against torch autograd on the same parameters: loss, every parameter gradient,
BatchNorm running statistics and the Adam update. bf16 activations / activation gradients -> gradients are compared by
relative L2 error per tensor (the stated bar for this mixed-precision step: <= 5e-2; typical 1e-2)."""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from model.train import TrainStep
from model.unet import FrameInterpolationUNet

pytestmark = pytest.mark.gpu


class _RoundBf16(torch.autograd.Function):
    """Round to bf16 in the forward and in the backward: what a bf16 activation / activation-gradient store does."""

    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


def ref_forward_train(model, x, emulate_bf16=False, training=True):
    """The reference network's forward (model/unet.py:84-95, either decoder) in training mode, plain torch ops.
    emulate_bf16=True rounds weights, pre-BN outputs, activations and their gradients to bf16 the way any bf16
    training step stores them: the distance between the two torch results is the noise floor of the precision."""
    u = model.unet
    R = _RoundBf16.apply if emulate_bf16 else (lambda t: t)

    def dconv(dc, t, stem=False):
        s = dc.double_conv
        for k, (conv, bn) in enumerate(((s[0], s[1]), (s[3], s[4]))):
            wq = conv.weight if (stem and k == 0) else R(conv.weight)
            z = R(F.conv2d(t, wq, None, padding=1))
            t = R(F.relu(F.batch_norm(z, bn.running_mean, bn.running_var, bn.weight, bn.bias, training, 0.1, bn.eps)))
        return t

    x1 = dconv(u.inc, x, stem=True)
    feats = [x1]
    for d in (u.down1, u.down2, u.down3, u.down4):
        feats.append(dconv(d.maxpool_conv[1], F.max_pool2d(feats[-1], 2)))
    y = feats[4]
    for i, up in enumerate((u.up1, u.up2, u.up3, u.up4)):
        if isinstance(up.up, nn.ConvTranspose2d):   # class default decoder, model/unet.py:43
            y = R(F.conv_transpose2d(y, R(up.up.weight), up.up.bias, stride=2))
        else:
            y = R(F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=True))
        y = dconv(up.conv, torch.cat([feats[3 - i], y], 1))
    return u.outc.conv(y)


def make_model(seed, bilinear=True):
    torch.manual_seed(seed)
    m = FrameInterpolationUNet(bilinear=bilinear)
    g = torch.Generator().manual_seed(seed + 1)
    for mod in m.modules():  # non-trivial affine parameters so that their gradients are exercised
        if isinstance(mod, nn.BatchNorm2d):
            mod.weight.data = torch.rand(mod.num_features, generator=g) + 0.5
            mod.bias.data = torch.randn(mod.num_features, generator=g) * 0.1
    return m


@pytest.mark.parametrize("bilinear", [True, False], ids=["bilinear", "convt"])
@pytest.mark.parametrize("n,h,w", [(2, 32, 32), (3, 48, 64)])
def test_gradients_match_autograd(cuda_device, n, h, w, bilinear):
    """Loss, output, every parameter gradient and the BatchNorm running statistics against fp32 torch autograd.

    The step stores activations and activation gradients in bf16. Through 18 BatchNorm+ReLU stages that rounding
    flips ReLU / max-pool decisions, so ANY bf16 step drifts from the fp32 gradient (about 4 % on the output and up
    to ~50 % relative L2 on the first layers' gradients at random initialisation). The bar is therefore the noise
    floor itself: per tensor, our distance to fp32 autograd must not exceed 1.5x the distance of a torch emulation
    of the same bf16 stores (+0.03). The kernels on their own are held to 1 bf16 ulp in test_gpu_train_kernels.py."""
    ref = make_model(0, bilinear).train()
    emu = copy.deepcopy(ref)
    ours = copy.deepcopy(ref).to(cuda_device).train()
    g = torch.Generator().manual_seed(5)
    f1, f2 = torch.rand(n, 1, h, w, generator=g) * 2 - 1, torch.rand(n, 1, h, w, generator=g) * 2 - 1
    tgt = torch.rand(n, 1, h, w, generator=g) * 2 - 1
    out = ref_forward_train(ref, torch.cat([f1, f2], 1))
    loss_ref = F.mse_loss(out, tgt)
    loss_ref.backward()
    out_emu = ref_forward_train(emu, torch.cat([f1, f2], 1), emulate_bf16=True)
    F.mse_loss(out_emu, tgt).backward()

    step = TrainStep(ours, lr=0.0)   # lr 0: gradients and statistics are produced, parameters stay put
    loss = step(f1.to(cuda_device), f2.to(cuda_device), tgt.to(cuda_device))
    assert abs(loss.item() - loss_ref.item()) <= 5e-3 * abs(loss_ref.item()) + 1e-4, (loss.item(), loss_ref.item())

    def rel(a, b):
        return ((a - b).norm() / (b.norm() + 1e-12)).item()

    floor = rel(out_emu.detach(), out.detach())
    assert rel(step.last_output.cpu(), out.detach()) <= 1.5 * floor + 0.03, floor
    worst = 0.0
    for (name, p_ref), (_, p_emu), (_, p) in zip(ref.named_parameters(), emu.named_parameters(), ours.named_parameters()):
        e_ours, e_floor = rel(step.grad_view[p].cpu(), p_ref.grad), rel(p_emu.grad, p_ref.grad)
        worst = max(worst, e_ours / (1.5 * e_floor + 0.03))
        assert e_ours <= 1.5 * e_floor + 0.03, f"{name}: gradient error {e_ours:.4f}, bf16 noise floor {e_floor:.4f}"
    print(f"worst gradient error / allowed {worst:.3f}")
    for (name, b_ref), (_, b) in zip(ref.named_buffers(), ours.named_buffers()):
        if name.endswith("running_mean") or name.endswith("running_var"):
            assert torch.allclose(b.cpu(), b_ref, rtol=5e-2, atol=5e-3), name
        if name.endswith("num_batches_tracked"):
            assert int(b) == 1


@pytest.mark.parametrize("bilinear", [True, False], ids=["bilinear", "convt"])
def test_adam_steps_follow_torch(cuda_device, bilinear):
    """The optimiser: after one step every parameter equals torch.optim.Adam applied to OUR gradient (exact check of
    the update), and three steps reduce the loss along the torch trajectory (within the bf16 drift)."""
    ref = make_model(3, bilinear).train()
    ours = copy.deepcopy(ref).to(cuda_device).train()
    shadow = copy.deepcopy(ref)
    opt = torch.optim.Adam(ref.parameters(), lr=1e-4)
    opt_shadow = torch.optim.Adam(shadow.parameters(), lr=1e-4)
    step = TrainStep(ours, lr=1e-4)
    g = torch.Generator().manual_seed(9)
    f1, f2 = torch.rand(2, 1, 32, 48, generator=g) * 2 - 1, torch.rand(2, 1, 32, 48, generator=g) * 2 - 1
    tgt = (f1 + f2) / 2
    losses_ref, losses = [], []
    for it in range(3):
        opt.zero_grad()
        l = F.mse_loss(ref_forward_train(ref, torch.cat([f1, f2], 1)), tgt)
        l.backward()
        opt.step()
        losses_ref.append(l.item())
        losses.append(step(f1.to(cuda_device), f2.to(cuda_device), tgt.to(cuda_device)).item())
        if it == 0:
            for p_s, p in zip(shadow.parameters(), ours.parameters()):
                p_s.grad = step.grad_view[p].cpu().clone()
            opt_shadow.step()
            for (name, p_s), (_, p) in zip(shadow.named_parameters(), ours.named_parameters()):
                assert torch.allclose(p.detach().cpu(), p_s.detach(), rtol=1e-5, atol=2e-7), name
    assert abs(losses[0] - losses_ref[0]) <= 5e-3 * losses_ref[0]
    assert np.allclose(losses, losses_ref, rtol=0.1), (losses, losses_ref)
    assert losses[-1] < losses[0]
    # the eval-mode inference path picks up the trained parameters (running stats included)
    ours.eval()
    y = ours(f1.to(cuda_device), f2.to(cuda_device))
    assert torch.isfinite(y).all()
    with torch.no_grad():
        y_ref = ref_forward_train(ref, torch.cat([f1, f2], 1), training=False)
    assert (y.cpu() - y_ref).abs().max() < 0.25   # parameters differ by the drift above; same function family


@pytest.mark.parametrize("tag", ["mse", "combined", "mse_convt"])
def test_one_step_against_reference_golden(cuda_device, tag):
    """One optimisation step of the UNMODIFIED reference (tests/golden/train_golden.npz, made by
    oracle/make_train_golden.py: default init seed 0, batch 2x32x32, Adam lr 1e-4) — loss, output, the gradients next
    to the output (where the bf16 drift is < 2 %), every gradient's norm, and the updated head / running statistics."""
    import os
    GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    from model.train import CombinedLoss
    gold = np.load(os.path.join(GOLDEN_DIR, "train_golden.npz"))
    torch.manual_seed(0)
    model = FrameInterpolationUNet(bilinear=tag != "mse_convt").to(cuda_device).train()
    assert [k for k, _ in model.named_parameters()] == list(gold[f"{tag}_param_names"])
    g = torch.Generator().manual_seed(21)
    f0, f1 = torch.rand(2, 1, 32, 32, generator=g), torch.rand(2, 1, 32, 32, generator=g)
    gt = ((f0 + f1) / 2 + 0.05 * torch.randn(2, 1, 32, 32, generator=g)).clamp(0, 1)
    step = TrainStep(model, lr=1e-4, criterion=CombinedLoss() if tag == "combined" else None)
    loss = step(f0.to(cuda_device), f1.to(cuda_device), gt.to(cuda_device))
    assert abs(loss.item() - float(gold[f"{tag}_loss"])) <= 5e-3 * float(gold[f"{tag}_loss"])
    out_ref = torch.from_numpy(gold[f"{tag}_output"])
    assert ((step.last_output.cpu() - out_ref).norm() / out_ref.norm()).item() < 0.08
    params = dict(model.named_parameters())
    for key in gold.files:
        if key.startswith(f"{tag}_grad:"):   # "mse_grad:" does not prefix "mse_convt_grad:"
            name = key.split(":", 1)[1]
            g_ref = torch.from_numpy(gold[key])
            rel = ((step.grad_view[params[name]].cpu() - g_ref).norm() / g_ref.norm()).item()
            print(f"{tag} {name}: gradient rel L2 vs the reference step {rel:.4f}")
            # the transposed conv sits one conv + BatchNorm + ReLU stage further from the loss than the other stored
            # gradients: its bf16 drift is correspondingly larger
            assert rel < (0.25 if ".up." in name else 0.12), (name, rel)
    norms = np.array([step.grad_view[p].norm().item() for p in model.parameters()])
    assert np.allclose(norms, gold[f"{tag}_grad_norms"], rtol=0.35), np.abs(norms / gold[f"{tag}_grad_norms"] - 1).max()
    sd = model.state_dict()
    assert np.allclose(sd["unet.outc.conv.weight"].cpu().numpy(), gold[f"{tag}_after:unet.outc.conv.weight"], atol=2.1e-4)
    for k in ("running_mean", "running_var"):
        assert np.allclose(sd[f"unet.inc.double_conv.1.{k}"].cpu().numpy(),
                           gold[f"{tag}_after:unet.inc.double_conv.1.{k}"], rtol=2e-2, atol=2e-3)


@pytest.mark.parametrize("criterion,bilinear", [("mse", True), ("combined", True), ("mse", False)],
                         ids=["mse", "combined", "mse-convt"])
def test_cuda_graph_replay_matches_eager(cuda_device, criterion, bilinear):
    """cuda_graph=True captures the step after two eager ones; replayed steps follow the eager trajectory, and the
    learning rate / step count reach the Adam kernel through device memory (lr = 0 after capture freezes the weights).

    Two runs of the same step are not bit-identical: BatchNorm sums use fp32 atomics, a last-bit change of a mean flips
    a few bf16 roundings and every later layer amplifies that up to the bf16 noise floor (tools/debug_determinism.py) —
    so trajectories are compared loosely (25 %), the first step (same starting state) to 1 %."""
    from model.train import CombinedLoss
    crit = (lambda: CombinedLoss()) if criterion == "combined" else (lambda: None)
    eager_model = make_model(7, bilinear).to(cuda_device).train()
    graph_model = copy.deepcopy(eager_model)
    eager = TrainStep(eager_model, lr=1e-4, criterion=crit())
    graphed = TrainStep(graph_model, lr=1e-4, criterion=crit(), cuda_graph=True)
    g = torch.Generator().manual_seed(2)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, 64), torch.linspace(0, 1, 64), indexing="ij")
    le, lg = [], []
    for it in range(6):   # smooth moving patterns plus a little noise: better conditioned than white noise
        ph = torch.rand(4, 1, 1, 1, generator=g) * 6.28
        f1 = (0.5 + 0.4 * torch.sin(9 * xx + 5 * yy + ph) + 0.02 * torch.randn(4, 1, 64, 64, generator=g)).to(cuda_device)
        f2 = (0.5 + 0.4 * torch.sin(9 * xx + 5 * yy + ph + 0.6) + 0.02 * torch.randn(4, 1, 64, 64, generator=g)).to(cuda_device)
        tgt = (f1 + f2) / 2
        le.append(eager(f1, f2, tgt).item())
        lg.append(graphed(f1, f2, tgt).item())
    assert len(graphed._graphs) == 1
    assert abs(le[0] - lg[0]) <= 1e-2 * le[0], (le, lg)
    assert np.allclose(le, lg, rtol=0.25), (le, lg)
    assert min(lg[2:]) < lg[0], lg
    assert eager.step_count == graphed.step_count == 6
    assert int(graph_model.unet.inc.double_conv[1].num_batches_tracked) == 6
    # schedule change after capture: lr = 0 must freeze every parameter on the replayed graph
    before = graphed.flat_param.clone()
    graphed.lr = 0.0
    graphed(f1, f2, tgt)
    assert torch.equal(before, graphed.flat_param)
    graphed.lr = 1e-4
    graphed(f1, f2, tgt)
    moved = (graphed.flat_param - before).abs()
    assert moved.max() <= 1e-4 * 4 and moved.mean() > 1e-6, (moved.max().item(), moved.mean().item())   # Adam steps are O(lr)


def test_rgb_unet_training_step(cuda_device):
    """UNet(6, 3, bilinear=True) — colour frame pairs in, colour frame out: the 6-channel stem, the 3-class head and
    their gradients (frame2=None: the input tensor already holds both frames)."""
    from model.unet import UNet
    torch.manual_seed(11)
    ref = UNet(6, 3, bilinear=True).train()
    ours = copy.deepcopy(ref).to(cuda_device).train()
    g = torch.Generator().manual_seed(12)
    x = torch.rand(2, 6, 32, 48, generator=g) * 2 - 1
    tgt = torch.rand(2, 3, 32, 48, generator=g) * 2 - 1

    class Wrap:          # ref_forward_train expects the FrameInterpolationUNet attribute layout
        def __init__(self, u):
            self.unet = u
    out = ref_forward_train(Wrap(ref), x)
    loss_ref = F.mse_loss(out, tgt)
    loss_ref.backward()
    step = TrainStep(ours, lr=0.0)
    loss = step(x.to(cuda_device), None, tgt.to(cuda_device))
    assert abs(loss.item() - loss_ref.item()) <= 5e-3 * loss_ref.item()
    assert step.last_output.shape == (2, 3, 32, 48)
    rel = lambda a, b: ((a - b).norm() / (b.norm() + 1e-12)).item()  # noqa: E731
    assert rel(step.last_output.cpu(), out.detach()) < 0.08
    # head gradients sit next to the output (noise ~1 %), the stem's at the far end (bf16 noise floor ~50 %)
    assert rel(step.grad_view[ours.outc.conv.weight].cpu(), ref.outc.conv.weight.grad) < 0.05
    assert rel(step.grad_view[ours.outc.conv.bias].cpu(), ref.outc.conv.bias.grad) < 0.05
    g_stem, g_ref = step.grad_view[ours.inc.double_conv[0].weight].cpu(), ref.inc.double_conv[0].weight.grad
    assert g_stem.shape == (64, 6, 3, 3) and rel(g_stem, g_ref) < 1.0
    cos = (g_stem * g_ref).sum() / (g_stem.norm() * g_ref.norm())
    assert cos > 0.5


def test_train_model_end_to_end_on_a_frame_directory(cuda_device, tmp_path, monkeypatch):
    """The reference's training entry point on a tiny synthetic data set: FrameTripletDataset -> DataLoader ->
    train_model (CombinedLoss, Adam, plateau schedule) -> best_model.pth with the reference's checkpoint keys, which
    model.inference.load_model reads back for inference."""
    import cv2
    from model import train as T
    from model.inference import load_model
    rs = np.random.RandomState(0)
    for video in ("a", "b", "c"):
        d = tmp_path / "data" / video
        d.mkdir(parents=True)
        for i in range(6):
            img = np.zeros((64, 80), np.uint8)
            cv2.circle(img, (12 + 9 * i, 30 + 2 * i), 9, 255, -1)
            img = cv2.GaussianBlur(img, (5, 5), 0) + rs.randint(0, 8, img.shape).astype(np.uint8)
            cv2.imwrite(str(d / f"frame_{i:03d}.png"), img)
    ds = T.FrameTripletDataset(str(tmp_path / "data"))
    assert len(ds) == 12 and ds[0][0].shape == (1, 256, 256)
    monkeypatch.chdir(tmp_path)   # train_model writes best_model.pth into the cwd, like the reference
    train_set, val_set = torch.utils.data.random_split(ds, [9, 3], generator=torch.Generator().manual_seed(0))
    mk = lambda d, sh: torch.utils.data.DataLoader(d, batch_size=3, shuffle=sh, num_workers=0)  # noqa: E731
    torch.manual_seed(0)
    model = FrameInterpolationUNet(bilinear=True).to(cuda_device)
    train_losses, val_losses = T.train_model(model, mk(train_set, True), mk(val_set, False), num_epochs=3,
                                             device=cuda_device, lr=1e-3)
    assert len(train_losses) == len(val_losses) == 3 and all(np.isfinite(train_losses + val_losses))
    assert train_losses[-1] < train_losses[0]
    ck = torch.load(tmp_path / "best_model.pth", map_location="cpu")
    assert {"epoch", "model_state_dict", "optimizer_state_dict", "train_loss", "val_loss", "train_losses",
            "val_losses"} <= set(ck)
    assert set(ck["model_state_dict"]) == set(model.state_dict())
    st = ck["optimizer_state_dict"]
    assert st["param_groups"][0]["lr"] == 1e-3 and len(st["state"]) == len(list(model.parameters()))
    loaded = load_model(str(tmp_path / "best_model.pth"), cuda_device)
    f0, f1, gt = (t[None].to(cuda_device) for t in ds[0])
    out = loaded(f0, f1)
    assert out.shape == (1, 1, 256, 256) and torch.isfinite(out).all()
    # the CLI wrapper (main.py train -> model.train.main) runs the same path
    assert T.main(["--data-dir", str(tmp_path / "data"), "--epochs", "1", "--batch-size", "4", "--device", "cuda"]) is None


def test_optimizer_state_round_trip(cuda_device):
    """TrainStep.state_dict() has torch.optim.Adam's layout (what train_model saves as optimizer_state_dict) and
    load_state_dict() resumes from it: step count, learning rate and both moment vectors."""
    model = make_model(5).to(cuda_device).train()
    step = TrainStep(model, lr=3e-4)
    g = torch.Generator().manual_seed(1)
    f1, f2 = torch.rand(2, 1, 32, 32, generator=g).to(cuda_device), torch.rand(2, 1, 32, 32, generator=g).to(cuda_device)
    for _ in range(2):
        step(f1, f2, (f1 + f2) / 2)
    sd = step.state_dict()
    names = [n for n, _ in model.named_parameters()]
    assert len(sd["state"]) == len(names) and sd["param_groups"][0]["params"] == list(range(len(names)))
    ref_opt = torch.optim.Adam(model.parameters(), lr=1.0)
    ref_opt.load_state_dict(sd)                       # torch accepts the layout as is
    assert ref_opt.param_groups[0]["lr"] == 3e-4
    resumed = TrainStep(copy.deepcopy(model), lr=1.0)
    resumed.load_state_dict(sd)
    assert resumed.step_count == 2 and resumed.lr == 3e-4 and resumed.betas == (0.9, 0.999)
    assert torch.equal(resumed.m, step.m) and torch.equal(resumed.v, step.v)
    before = resumed.flat_param.clone()
    resumed(f1, f2, (f1 + f2) / 2)
    assert resumed.step_count == 3 and (resumed.flat_param - before).abs().max() <= 3e-4 * 1.5


@pytest.mark.parametrize("criterion", ["mse", "combined", "l1"])
def test_validation_loss_matches_torch_criterion(cuda_device, criterion):
    """TrainStep.loss_value (the validation loss of train_model, reference model/train.py:204-219) against the torch
    criterion on the same tensors; a criterion the library has no kernel for is evaluated by torch itself."""
    from model.train import CombinedLoss

    crit = {"mse": None, "combined": CombinedLoss(), "l1": nn.L1Loss()}[criterion]
    step = TrainStep(make_model(3).to(cuda_device), criterion=crit)
    g = torch.Generator().manual_seed(11)
    y = torch.rand(3, 1, 40, 56, generator=g).to(cuda_device)
    t = (y.cpu() * 0.7 + 0.3 * torch.rand(3, 1, 40, 56, generator=g)).to(cuda_device)
    got = step.loss_value(y, t)
    assert got.shape == (1,) and got.device.type == "cuda"
    want = (crit if crit is not None else nn.MSELoss())(y, t)
    assert abs(got.item() - want.item()) <= 2e-5 * max(1.0, abs(want.item()))
