"""Generates tests/golden/unet_golden.npz by running the UNMODIFIED reference module (read-only import from
/root/reference/model/unet.py) in the build container. Also asserts that oracle.unet_oracle.init_state_dict reproduces
the reference's default initialisation bit for bit, so the fixtures need not store 124 MB of weights.

    python oracle/make_golden.py            # rewrites tests/golden/unet_golden.npz

/root/reference does not exist on the GPU box: nothing at test or bench time reads it — only this script does.
"""
import hashlib
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference/model")

import unet as ref_unet  # noqa: E402  (the reference, unmodified)
from oracle import unet_oracle as O  # noqa: E402

CASES = [
    # name, n_channels, n_classes, bilinear, wrapper, N, H, W, stressed
    ("convt_64", 2, 1, False, True, 1, 64, 64, False),
    ("bilinear_64", 2, 1, True, True, 1, 64, 64, False),
    ("convt_odd_70x54", 2, 1, False, True, 1, 70, 54, False),      # F.pad path (reference model/unet.py:49-53)
    ("bilinear_odd_70x54", 2, 1, True, True, 2, 70, 54, False),
    ("convt_stressed_48x80", 2, 1, False, True, 1, 48, 80, True),
    ("bilinear_stressed_48x80", 2, 1, True, True, 1, 48, 80, True),
    ("rgb_6_3_40x56", 6, 3, False, False, 1, 40, 56, False),
]


def sd_digest(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def main():
    out = {}
    for name, n_ch, n_cls, bil, wrapper, n, h, w, stressed in CASES:
        seed = 0
        torch.manual_seed(seed)
        m = (ref_unet.FrameInterpolationUNet(bilinear=bil) if wrapper else ref_unet.UNet(n_ch, n_cls, bil)).eval()
        sd = O.init_state_dict(seed, n_ch, n_cls, bil, prefix="unet." if wrapper else "")
        ref_sd = m.state_dict()
        assert list(sd) == list(ref_sd), name
        assert all(torch.equal(sd[k], ref_sd[k]) for k in sd), f"{name}: oracle init differs from the reference init"
        rs = np.random.RandomState(hash(name) % (2 ** 31) if False else len(name) * 7919 + h * 31 + w)
        frames = rs.randint(0, 256, size=(n, n_ch, h, w)).astype(np.uint8)
        x = O.preprocess_u8(frames)
        if stressed:
            sd = O.calibrate_head(O.stress_state_dict(sd, seed=1), x)
            m.load_state_dict(sd)
        with torch.no_grad():
            y = m(x[:, :1], x[:, 1:]) if wrapper else m(x)
        out[name + "/frames"] = frames
        if stressed:
            # the non-default tensors of the stressed fixture (BatchNorm vectors + calibrated head), so that a test can
            # rebuild the exact weights from a seeded module without calling the oracle
            base = O.init_state_dict(seed, n_ch, n_cls, bil, prefix="unet." if wrapper else "")
            for k, v in sd.items():
                if v.dtype.is_floating_point and not torch.equal(v, base[k]):
                    out[name + "/sd/" + k] = v.numpy()
        out[name + "/logits"] = y.numpy()
        out[name + "/sd_sha256"] = np.frombuffer(bytes.fromhex(sd_digest(sd)), dtype=np.uint8)
        out[name + "/cfg"] = np.array([n_ch, n_cls, int(bil), int(wrapper), int(stressed)], dtype=np.int32)
        print(f"{name}: logits range [{y.min():.4f}, {y.max():.4f}] std {y.std():.4f}")
    # pre/post-processing vectors (reference model/inference.py:32-35, 54-61), restated by hand here because
    # model/inference.py does not import without imageio; the arithmetic is numpy/torch one-liners
    u8 = np.arange(256, dtype=np.uint8)
    out["pre/u8"] = u8
    out["pre/norm"] = (2.0 * (u8.astype(np.float32) / 255.0) - 1.0).astype(np.float32)
    t = torch.linspace(-1.5, 1.5, 4001)
    img = torch.clamp((t + 1.0) / 2.0, 0.0, 1.0)
    out["post/in"] = t.numpy()
    out["post/out"] = (img.numpy() * 255).astype(np.uint8)
    dst = ROOT / "tests" / "golden" / "unet_golden.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, dst.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
