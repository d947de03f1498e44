"""Generates tests/golden/train_golden.npz by running the UNMODIFIED reference training code (read-only import of
/root/reference/model/train.py and unet.py) in the build container: CombinedLoss values and gradients on seeded
images, and one optimisation step (train-mode forward, CombinedLoss / MSE, backward, Adam lr=1e-4) of the default
initialised FrameInterpolationUNet(bilinear=True) on a seeded 2x32x32 batch.

    python oracle/make_train_golden.py

Test infrastructure: nothing at test or bench time reads /root/reference — only this script does.
"""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, "/root/reference/model")
spec = importlib.util.spec_from_file_location("ref_train", "/root/reference/model/train.py")
ref_train = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_train)
import unet as ref_unet  # noqa: E402


def batch(seed, n, h, w):
    g = torch.Generator().manual_seed(seed)
    f0, f1 = torch.rand(n, 1, h, w, generator=g), torch.rand(n, 1, h, w, generator=g)
    return f0, f1, ((f0 + f1) / 2 + 0.05 * torch.randn(n, 1, h, w, generator=g)).clamp(0, 1)


def main():
    out = {}
    # ---- the loss alone
    g = torch.Generator().manual_seed(11)
    pred = torch.rand(2, 1, 40, 56, generator=g).requires_grad_(True)
    target = (pred.detach() + 0.1 * torch.randn(2, 1, 40, 56, generator=g)).clamp(0, 1)
    loss = ref_train.CombinedLoss()(pred, target)
    loss.backward()
    out["loss_pred"], out["loss_target"] = pred.detach().numpy(), target.numpy()
    out["loss_value"], out["loss_grad"] = np.float64(loss.item()), pred.grad.numpy()
    out["ssim_loss_value"] = np.float64(ref_train.SSIMLoss()(pred.detach(), target).item())
    # ---- one training step, both criteria
    # "mse_convt": the class-default ConvTranspose2d decoder (model/unet.py:99), which reference train.py never builds but
    # its model file defines; same recipe
    for tag, crit, bilinear in (("combined", ref_train.CombinedLoss(), True), ("mse", nn.MSELoss(), True),
                                ("mse_convt", nn.MSELoss(), False)):
        torch.manual_seed(0)
        model = ref_unet.FrameInterpolationUNet(bilinear=bilinear).train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)
        f0, f1, gt = batch(21, 2, 32, 32)
        opt.zero_grad()
        y = model(f0, f1)
        l = crit(y, gt)
        l.backward()
        grads = {k: p.grad.clone() for k, p in model.named_parameters()}
        opt.step()
        out[f"{tag}_loss"] = np.float64(l.item())
        out[f"{tag}_output"] = y.detach().numpy()
        out[f"{tag}_grad_norms"] = np.array([grads[k].norm().item() for k in grads], dtype=np.float64)
        keys = ["unet.outc.conv.weight", "unet.outc.conv.bias", "unet.up4.conv.double_conv.4.weight",
                "unet.up4.conv.double_conv.4.bias", "unet.up4.conv.double_conv.3.weight"]
        if not bilinear:
            keys += ["unet.up4.up.weight", "unet.up4.up.bias"]
        for k in keys:
            out[f"{tag}_grad:{k}"] = grads[k].numpy()
        sd = model.state_dict()
        out[f"{tag}_after:unet.outc.conv.weight"] = sd["unet.outc.conv.weight"].numpy()
        out[f"{tag}_after:unet.inc.double_conv.1.running_mean"] = sd["unet.inc.double_conv.1.running_mean"].numpy()
        out[f"{tag}_after:unet.inc.double_conv.1.running_var"] = sd["unet.inc.double_conv.1.running_var"].numpy()
        out[f"{tag}_param_names"] = np.array([k for k, _ in model.named_parameters()])
    path = ROOT / "tests" / "golden" / "train_golden.npz"
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({path.stat().st_size} bytes)")


if __name__ == "__main__":
    main()
