"""CPU oracle for the UNet frame-synthesis path — TEST INFRASTRUCTURE ONLY.

This file restates, in plain fp32 torch functional ops on the CPU, the algorithm of the reference
(daultanigaurav/AI-BASED-FRAME-INTERPOLATION). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it; the product path (ai-based-frame-interpolation_b200/) never does.

Pinning: the reference ships no golden vectors for this path (its only test asserts an output shape,
test_basic.py:72-77). The restatement is therefore pinned against the reference module itself, imported read-only from
/root/reference in the build container by oracle/make_golden.py, which writes tests/golden/unet_golden.npz; the oracle
is checked against those vectors in tests/test_oracle.py.  torch (unpinned in the reference's requirements.txt:1;
2.11.0 here) is the third-party dependency that holds the conv / batch-norm arithmetic.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default, reference model/unet.py:13,16


# ----------------------------------------------------------------------------------------------- state dict
def _double_conv_keys(prefix):
    return [f"{prefix}.double_conv.0", f"{prefix}.double_conv.1", f"{prefix}.double_conv.3", f"{prefix}.double_conv.4"]


def layer_table(n_channels=2, n_classes=1, bilinear=False):
    """(name, cin, mid, cout) of the nine DoubleConv blocks, reference model/unet.py:72-82 and :35-45."""
    f = 2 if bilinear else 1
    t = [("inc", n_channels, 64, 64)]
    enc = [64, 128, 256, 512, 1024 // f]
    for i in range(1, 5):
        t.append((f"down{i}.maxpool_conv.1", enc[i - 1], enc[i], enc[i]))
    up_in = [1024, 512, 256, 128]
    up_out = [512 // f, 256 // f, 128 // f, 64]
    for i in range(4):
        mid = up_in[i] // 2 if bilinear else up_out[i]
        t.append((f"up{i + 1}.conv", up_in[i], mid, up_out[i]))
    return t


def _kaiming_bound(fan_in):
    # torch.nn.init.kaiming_uniform_(a=math.sqrt(5)), same floating-point expression order as torch
    gain = math.sqrt(2.0 / (1 + math.sqrt(5) ** 2))
    std = gain / math.sqrt(fan_in)
    return math.sqrt(3.0) * std


def init_state_dict(seed=0, n_channels=2, n_classes=1, bilinear=False, prefix="unet."):
    """Default PyTorch initialisation consumed in the reference's construction order (model/unet.py:66-82) under
    torch.manual_seed(seed) — bit-identical to building the reference module after the same seed (checked in
    oracle/make_golden.py)."""
    g = torch.Generator().manual_seed(seed)

    def conv_w(cout, cin, k):
        w = torch.empty(cout, cin, k, k)
        # nn.Conv2d.reset_parameters: kaiming_uniform_(a=sqrt(5)) -> U(-bound, bound), bound ~ 1/sqrt(fan_in)
        return w.uniform_(-_kaiming_bound(cin * k * k), _kaiming_bound(cin * k * k), generator=g)

    sd = OrderedDict()

    def bn(key, c):
        sd[f"{key}.weight"] = torch.ones(c)
        sd[f"{key}.bias"] = torch.zeros(c)
        sd[f"{key}.running_mean"] = torch.zeros(c)
        sd[f"{key}.running_var"] = torch.ones(c)
        sd[f"{key}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    def double_conv(name, cin, mid, cout):
        k = _double_conv_keys(name)
        sd[f"{k[0]}.weight"] = conv_w(mid, cin, 3)
        bn(k[1], mid)
        sd[f"{k[2]}.weight"] = conv_w(cout, mid, 3)
        bn(k[3], cout)

    table = layer_table(n_channels, n_classes, bilinear)
    for name, cin, mid, cout in table[:5]:
        double_conv(name, cin, mid, cout)
    for i, (name, cin, mid, cout) in enumerate(table[5:]):
        if not bilinear:
            # nn.ConvTranspose2d(cin, cin//2, 2, 2): weight [cin, cin//2, 2, 2]; fan_in is computed from dim 1
            fan_in = (cin // 2) * 4
            wb, bb = _kaiming_bound(fan_in), 1.0 / math.sqrt(fan_in)
            sd[f"up{i + 1}.up.weight"] = torch.empty(cin, cin // 2, 2, 2).uniform_(-wb, wb, generator=g)
            sd[f"up{i + 1}.up.bias"] = torch.empty(cin // 2).uniform_(-bb, bb, generator=g)
        double_conv(name, cin, mid, cout)
    wb, bb = _kaiming_bound(64), 1.0 / math.sqrt(64)
    sd["outc.conv.weight"] = torch.empty(n_classes, 64, 1, 1).uniform_(-wb, wb, generator=g)
    sd["outc.conv.bias"] = torch.empty(n_classes).uniform_(-bb, bb, generator=g)
    return OrderedDict((prefix + k, v) for k, v in sd.items())


def stress_state_dict(sd, seed=1, out_std=0.5):
    """randomise every BatchNorm (gamma~U(.5,1.5), beta,mu~N(0,.1^2), var~U(.5,1.5)); the head is
    rescaled by calibrate_head() so the output spans both clamps of postprocess_image."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for k, v in sd.items():
        out[k] = v.clone()
    for k in list(out.keys()):
        if k.endswith("running_var"):
            base = k[: -len("running_var")]
            c = out[k].numel()
            out[base + "weight"] = torch.rand(c, generator=g) + 0.5
            out[base + "bias"] = torch.randn(c, generator=g) * 0.1
            out[base + "running_mean"] = torch.randn(c, generator=g) * 0.1
            out[base + "running_var"] = torch.rand(c, generator=g) + 0.5
    return out


def calibrate_head(sd, x, out_std=0.5):
    """Rescale outc so the oracle output on x has mean 0 and the given std."""
    y = unet_forward(sd, x)
    key_w = [k for k in sd if k.endswith("outc.conv.weight")][0]
    key_b = [k for k in sd if k.endswith("outc.conv.bias")][0]
    s = out_std / float(y.std())
    sd[key_w] = sd[key_w] * s                      # y' = s*(y - mean(y))
    sd[key_b] = (sd[key_b] - float(y.mean())) * s
    return sd


# ----------------------------------------------------------------------------------------------- forward
def _get(sd, key):
    if key in sd:
        return sd[key]
    return sd["unet." + key]


def _has(sd, key):
    return key in sd or ("unet." + key) in sd


def conv_bn_relu(sd, conv_key, bn_key, x):
    """conv3x3(pad 1, no bias) -> BatchNorm2d(eval) -> ReLU, reference model/unet.py:12-17."""
    x = F.conv2d(x, _get(sd, conv_key + ".weight"), None, stride=1, padding=1)
    x = F.batch_norm(x, _get(sd, bn_key + ".running_mean"), _get(sd, bn_key + ".running_var"),
                     _get(sd, bn_key + ".weight"), _get(sd, bn_key + ".bias"), False, 0.0, BN_EPS)
    return F.relu(x)


def double_conv(sd, name, x, taps=None):
    k = _double_conv_keys(name)
    x = conv_bn_relu(sd, k[0], k[1], x)
    if taps is not None:
        taps[name + ".mid"] = x
    return conv_bn_relu(sd, k[2], k[3], x)


def up_block(sd, i, x1, x2, taps=None):
    """Up.forward, reference model/unet.py:46-55."""
    if _has(sd, f"up{i}.up.weight"):
        x1 = F.conv_transpose2d(x1, _get(sd, f"up{i}.up.weight"), _get(sd, f"up{i}.up.bias"), stride=2)
    else:
        x1 = F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=True)
    if taps is not None:
        taps[f"up{i}.up"] = x1
    dy = x2.shape[2] - x1.shape[2]
    dx = x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    return double_conv(sd, f"up{i}.conv", torch.cat([x2, x1], dim=1), taps)


def unet_forward(sd, x, taps=None):
    """UNet.forward, reference model/unet.py:84-95. x: fp32 [N, n_channels, H, W] -> [N, n_classes, H, W]."""
    with torch.no_grad():
        x1 = double_conv(sd, "inc", x, taps)
        feats = [x1]
        for i in range(1, 5):
            feats.append(double_conv(sd, f"down{i}.maxpool_conv.1", F.max_pool2d(feats[-1], 2), taps))
        y = feats[4]
        for i in range(1, 5):
            y = up_block(sd, i, y, feats[4 - i], taps)
            if taps is not None:
                taps[f"up{i}"] = y
        if taps is not None:
            taps["inc"] = feats[0]
            for i in range(1, 5):
                taps[f"down{i}"] = feats[i]
        return F.conv2d(y, _get(sd, "outc.conv.weight"), _get(sd, "outc.conv.bias"))


def frame_interp_forward(sd, frame1, frame2):
    """FrameInterpolationUNet.forward, reference model/unet.py:105-112."""
    return unet_forward(sd, torch.cat([frame1, frame2], dim=1))


# ----------------------------------------------------------------------------------------------- pre / post
def preprocess_u8(img_u8):
    """Normalisation half of preprocess_image, reference model/inference.py:32-39 (resize/imread stay in cv2)."""
    a = img_u8.astype(np.float32) / 255.0
    a = 2.0 * a - 1.0
    return torch.from_numpy(np.ascontiguousarray(a))


def postprocess(t):
    """postprocess_image, reference model/inference.py:43-63 (without the squeeze/cpu plumbing)."""
    image = (t + 1.0) / 2.0
    image = torch.clamp(image, 0.0, 1.0)
    return (image.numpy() * 255).astype(np.uint8)


def flops_per_forward(n, h, w, n_channels=2, n_classes=1, bilinear=False):
    """2*MACs of the convolutions."""
    hs, ws = [h], [w]
    for _ in range(4):
        hs.append(hs[-1] // 2)
        ws.append(ws[-1] // 2)
    table = layer_table(n_channels, n_classes, bilinear)
    total = 0
    lvl = [0, 1, 2, 3, 4, 3, 2, 1, 0]
    for (name, cin, mid, cout), l in zip(table, lvl):
        total += 2 * n * hs[l] * ws[l] * 9 * (cin * mid + mid * cout)
    if not bilinear:
        for i, cin in enumerate([1024, 512, 256, 128]):
            l = 4 - i
            total += 2 * n * hs[l] * ws[l] * cin * (cin // 2) * 4
    total += 2 * n * h * w * 64 * n_classes
    return total
