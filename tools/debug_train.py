"""Per-layer activation and per-parameter gradient errors of the training step vs torch autograd (diagnostic)."""
import copy, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "ai-based-frame-interpolation_b200"))
import torch, torch.nn.functional as F
from test_gpu_train_step import make_model  # noqa: E402
from model.train import TrainStep

def run(n, h, w, emulate):
    dev = torch.device("cuda:0")
    ref = make_model(0).train()
    ours = copy.deepcopy(ref).to(dev).train()
    g = torch.Generator().manual_seed(5)
    f1, f2 = torch.rand(n, 1, h, w, generator=g) * 2 - 1, torch.rand(n, 1, h, w, generator=g) * 2 - 1
    tgt = torch.rand(n, 1, h, w, generator=g) * 2 - 1
    u = ref.unet
    acts = {}
    rb = (lambda t: t.to(torch.bfloat16).float()) if emulate else (lambda t: t)
    class RoundSTE(torch.autograd.Function):
        @staticmethod
        def forward(ctx, t): return rb(t)
        @staticmethod
        def backward(ctx, gr): return rb(gr)
    R = RoundSTE.apply
    def dconv(dc, t, name):
        s = dc.double_conv
        for k, (conv, bn) in zip((0, 3), ((s[0], s[1]), (s[3], s[4]))):
            wq = R(conv.weight) if (emulate and name + str(k) != "inc.0") else conv.weight
            z = R(F.conv2d(t, wq, None, padding=1))
            t = R(F.relu(F.batch_norm(z, None, None, bn.weight, bn.bias, True, 0.1, bn.eps)))
            acts[f"{name}.{k}"] = t
        return t
    x = torch.cat([f1, f2], 1)
    x1 = dconv(u.inc, x, "inc")
    feats = [x1]
    for i, d in enumerate((u.down1, u.down2, u.down3, u.down4), 1):
        feats.append(dconv(d.maxpool_conv[1], F.max_pool2d(feats[-1], 2), f"down{i}"))
    y = feats[4]
    for i, up in enumerate((u.up1, u.up2, u.up3, u.up4)):
        y = R(F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=True))
        y = dconv(up.conv, torch.cat([feats[3 - i], y], 1), f"up{i+1}")
    out = u.outc.conv(y)
    loss_ref = F.mse_loss(out, tgt); loss_ref.backward()
    step = TrainStep(ours, lr=0.0)
    step.keep_activations = True
    loss = step(f1.to(dev), f2.to(dev), tgt.to(dev))
    print(f"--- n={n} {h}x{w} emulate_bf16={emulate}: loss {loss.item():.6f} ref {loss_ref.item():.6f}")
    for k, v in acts.items():
        a = step.last_activations[k].float().cpu().permute(0, 3, 1, 2)
        print(f"  act {k:10s} rel {((a - v.detach()).norm() / v.detach().norm()).item():.5f}")
    print(f"  out rel {((step.last_output.cpu() - out.detach()).norm() / out.detach().norm()).item():.5f}")
    for (name, p_ref), (_, p) in zip(ref.named_parameters(), ours.named_parameters()):
        g_ref, g_ours = p_ref.grad, step.grad_view[p].cpu()
        print(f"  grad {name:45s} rel {((g_ours - g_ref).norm() / (g_ref.norm() + 1e-12)).item():.5f} |g| {g_ref.norm().item():.3e}")

for (n, h, w) in ((2, 32, 32), (4, 128, 128)):
    for em in (False, True):
        run(n, h, w, em)
