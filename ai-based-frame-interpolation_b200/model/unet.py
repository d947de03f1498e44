"""Drop-in for the reference's model/unet.py: same class names, constructor signatures, state-dict keys and tensor
contracts, but forward() runs on the B200-native library (libfi_b200.so) instead of eager torch ops.

The nn.Module objects below are parameter containers only: they exist so that `.state_dict()`,
`.load_state_dict()`, `.parameters()`, `.to(device)` and `.eval()` behave exactly like the reference
(state-dict schema: SURVEY.md A.5 / reference model/unet.py:5-112). The arithmetic lives in csrc/:
  DoubleConv / Down / Up / OutConv   -> conv_gemm.cu (tcgen05 implicit GEMM, BN folded, pool / concat / head fused)
  first conv of `inc`                -> aux_kernels.cu stem kernel (frame-pair cat + normalisation fused)
There is no CPU fallback and no training-mode forward (train-mode BatchNorm is a later row of the scope table).
"""
from __future__ import annotations

import torch
import torch.nn as nn

try:  # imported as `model.unet` (main.py) or as top-level `unet` with model/ on sys.path (reference scripts)
    from . import _engine as _E
except ImportError:  # pragma: no cover - depends on how the caller set sys.path
    import _engine as _E


def _conv_bn_relu_slots(cin, cout):
    """Parameter slots of one conv3x3(no bias) + BatchNorm2d + ReLU stage (indices 0,1,2 / 3,4,5 of double_conv)."""
    return [nn.Conv2d(cin, cout, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True)]


def _fused_only(name):
    raise _E.FiError(f"{name} runs as part of the fused UNet schedule: call UNet / FrameInterpolationUNet (its kernel "
                     "is reachable on its own through fiConvGemm / fiStemConv / fiUpsample2x of the C ABI)")


class _EngineBacked(nn.Module):
    """Mixin: lazily mirrors this module's parameters into a fiNet handle and re-uploads when they change.

    `module.precision` selects the arithmetic: "bf16" (default; bf16 operands, fp32 accumulate, <= 2e-2 pixel error) or
    "fp32" (hi/lo-split bf16 operands, three products per MAC: fp32-grade, <= 1e-3; about 3x slower)."""

    precision = "bf16"

    def _engine_spec(self):  # (n_channels, n_classes, bilinear)
        raise NotImplementedError

    def _fingerprint(self):
        return tuple((t.data_ptr(), t._version) for t in self.state_dict(keep_vars=True).values())

    def _engine(self, device):
        if self.training:
            raise _E.FiError("training-mode forward (batch-statistics BatchNorm) is not part of the B200 inference "
                             "path; call .eval() first")
        net = self.__dict__.get("_fi_net")
        if net is None or net.device != device or net.precision != self.precision:
            if net is not None:
                net.close()
            n_ch, n_cls, bil = self._engine_spec()
            net = _E.Net(device, n_ch, n_cls, bil, self.precision)
            self.__dict__["_fi_net"] = net
            self.__dict__["_fi_print"] = None
        fp = self._fingerprint()
        if self.__dict__.get("_fi_print") != fp:
            net.load_state_dict(self.state_dict())
            self.__dict__["_fi_print"] = fp
        return net

    def _device_of(self, x):
        dev = _E.require_cuda(x.device)
        p = next(self.parameters())
        if p.device != x.device:
            raise _E.FiError(f"module parameters are on {p.device} but the input is on {x.device}")
        return torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())


class DoubleConv(nn.Module):
    """(conv3x3 => BatchNorm => ReLU) * 2 — reference model/unet.py:5-21."""

    def __init__(self, in_channels, out_channels, mid_channels=None):
        super().__init__()
        mid = mid_channels if mid_channels else out_channels
        self.double_conv = nn.Sequential(*(_conv_bn_relu_slots(in_channels, mid) + _conv_bn_relu_slots(mid, out_channels)))

    def forward(self, x):
        _fused_only("DoubleConv")


class Down(nn.Module):
    """MaxPool2d(2) then DoubleConv — reference model/unet.py:23-33 (the pool is fused into the producer's epilogue)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))

    def forward(self, x):
        _fused_only("Down")


class Up(nn.Module):
    """Upscale, pad to the skip size, concat [skip, up], DoubleConv — reference model/unet.py:35-55."""

    def __init__(self, in_channels, out_channels, bilinear=True):
        super().__init__()
        if bilinear:
            self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
            self.conv = DoubleConv(in_channels, out_channels, in_channels // 2)
        else:
            self.up = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2)
            self.conv = DoubleConv(in_channels, out_channels)

    def forward(self, x1, x2):
        _fused_only("Up")


class OutConv(nn.Module):
    """1x1 output head — reference model/unet.py:57-63 (fused into the last conv's epilogue)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1)

    def forward(self, x):
        _fused_only("OutConv")


class UNet(_EngineBacked):
    """reference model/unet.py:65-95: widths 64-128-256-512-1024, four Down / four Up stages, 1x1 head."""

    def __init__(self, n_channels=2, n_classes=1, bilinear=False):
        super().__init__()
        self.n_channels, self.n_classes, self.bilinear = n_channels, n_classes, bilinear
        factor = 2 if bilinear else 1
        self.inc = DoubleConv(n_channels, 64)
        self.down1 = Down(64, 128)
        self.down2 = Down(128, 256)
        self.down3 = Down(256, 512)
        self.down4 = Down(512, 1024 // factor)
        self.up1 = Up(1024, 512 // factor, bilinear)
        self.up2 = Up(512, 256 // factor, bilinear)
        self.up3 = Up(256, 128 // factor, bilinear)
        self.up4 = Up(128, 64, bilinear)
        self.outc = OutConv(64, n_classes)

    def _engine_spec(self):
        return self.n_channels, self.n_classes, self.bilinear

    @torch.no_grad()
    def forward(self, x):
        """x: fp32 (normalised) or uint8 (raw) NCHW CUDA tensor -> fp32 logits [N, n_classes, H, W]."""
        net = self._engine(self._device_of(x))
        out, _ = net.forward(x, None, want_f32=True)
        return out

    @torch.no_grad()
    def forward_u8(self, x0, x1=None):
        """Raw uint8 planes in, post-processed uint8 frame out (normalisation and postprocess_image fused)."""
        net = self._engine(self._device_of(x0))
        _, out = net.forward(x0, x1, want_f32=False, want_u8=True)
        return out


class FrameInterpolationUNet(_EngineBacked):
    """reference model/unet.py:97-112: cat([frame1, frame2], 1) -> UNet(2, 1). The cat is fused into the stem loader."""

    def __init__(self, bilinear=False):
        super().__init__()
        self.unet = UNet(n_channels=2, n_classes=1, bilinear=bilinear)

    def _engine_spec(self):
        return 2, 1, self.unet.bilinear

    @torch.no_grad()
    def forward(self, frame1, frame2):
        """frame1, frame2: [B,1,H,W] fp32 in [-1,1] (or raw uint8) on the GPU -> [B,1,H,W] fp32."""
        net = self._engine(self._device_of(frame1))
        out, _ = net.forward(frame1, frame2, want_f32=True)
        return out

    @torch.no_grad()
    def forward_u8(self, frame1, frame2):
        net = self._engine(self._device_of(frame1))
        _, out = net.forward(frame1, frame2, want_f32=False, want_u8=True)
        return out


def count_parameters(model):
    """Number of trainable parameters — reference model/unet.py:114-116."""
    return sum(p.numel() for p in model.parameters() if p.requires_grad)
