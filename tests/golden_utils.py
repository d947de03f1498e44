"""Rebuilds the weights of a tests/golden/unet_golden.npz case WITHOUT the oracle: the product's own nn.Module skeleton
(model/unet.py) constructed under torch.manual_seed(0) draws the same default initialisation as the reference module
(same construction order), and the stressed cases overlay the BatchNorm / head tensors stored in the fixture. The
SHA-256 over the resulting state dict must equal the digest oracle/make_golden.py took from the reference's weights."""
import hashlib
from pathlib import Path

import numpy as np
import torch

GOLD = np.load(Path(__file__).parent / "golden" / "unet_golden.npz")
CASES = sorted({k.split("/")[0] for k in GOLD.files if k.endswith("/logits")})


def golden_case(name):
    """-> (module on CPU in eval mode, frames u8 [N,C,H,W], reference logits fp32 [N,n_classes,H,W])."""
    from model.unet import FrameInterpolationUNet, UNet
    n_ch, n_cls, bil, wrapper, stressed = (int(v) for v in GOLD[name + "/cfg"])
    torch.manual_seed(0)
    m = FrameInterpolationUNet(bilinear=bool(bil)) if wrapper else UNet(n_ch, n_cls, bool(bil))
    sd = m.state_dict()
    overlay = {k[len(name) + 4:]: torch.from_numpy(GOLD[k]) for k in GOLD.files if k.startswith(name + "/sd/")}
    assert bool(overlay) == bool(stressed)
    sd.update(overlay)
    m.load_state_dict(sd)
    h = hashlib.sha256()
    for k, v in m.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    assert h.digest() == GOLD[name + "/sd_sha256"].tobytes(), f"{name}: rebuilt weights differ from the reference's"
    return m.eval(), GOLD[name + "/frames"], GOLD[name + "/logits"]
