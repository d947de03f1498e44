"""Helpers shared by the GPU layer-parity tests: NHWC/bf16 conversions and the weight packing of include/fi_b200.h."""
import ctypes as C

import torch
import torch.nn.functional as F

from model import _engine as E


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def to_nhwc_bf16(t, device):
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(device)


def from_nhwc(t):
    return t.to(torch.float32).cpu().permute(0, 3, 1, 2).contiguous()


def pack_conv3x3(w):  # [cout, cin, 3, 3] -> bf16 [cout, 9*cin], K = tap*cin + ci
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous().to(torch.bfloat16)


def pack_conv3x3_dual(w, c0):  # concat order: [skip(c0) | up]
    return pack_conv3x3(w)


def pack_convT(w):  # [cin, cout, 2, 2] -> bf16 [4*cout, cin], row = (a*2+b)*cout + co
    cin, cout = w.shape[0], w.shape[1]
    return w.permute(2, 3, 1, 0).reshape(4 * cout, cin).contiguous().to(torch.bfloat16)


def split_bf16(t):
    """fp32 -> (hi, lo) bf16 pair with hi + lo ~ t to 16 mantissa bits (the precise-mode activation format)."""
    hi = t.to(torch.bfloat16)
    lo = (t - hi.to(torch.float32)).to(torch.bfloat16)
    return hi, lo


def pack3(wp):
    """[rows, taps, c] fp32 -> bf16 [rows, taps*3*c] as [w_hi | w_lo | w_hi] per tap."""
    hi, lo = split_bf16(wp)
    return torch.cat([hi, lo, hi], dim=2)


def run_conv_precise(device, x, w, bias, *, relu=True, mode=E.EPI_STORE, x1=None, off=(0, 0), head_w=None,
                     head_b=None):
    """Precise (bf16 hi/lo split, 3 products) variant of run_conv: fp32 in, fp32 (hi+lo) out."""
    n, c0, h, wd = x.shape
    keep = []

    def nhwc_pair(t):
        hi, lo = split_bf16(t.permute(0, 2, 3, 1).contiguous())
        hi, lo = hi.to(device), lo.to(device)
        keep.extend([hi, lo])
        return hi, lo

    d = E.ConvDesc()
    d.precise = 1
    s0h, s0l = nhwc_pair(x)
    d.src0, d.src0_lo, d.c0 = s0h.data_ptr(), s0l.data_ptr(), c0
    if mode == E.EPI_CONVT:
        cin, cout = w.shape[0], w.shape[1]
        wp = w.permute(2, 3, 1, 0).reshape(4 * cout, 1, cin)              # rows (a,b,co), K = ci
        n_total, b, d.taps = 4 * cout, bias.repeat(4).contiguous().to(device), 1
        packed = pack3(wp).reshape(n_total, -1)
    else:
        cout = w.shape[0]
        wp = w.permute(0, 2, 3, 1).reshape(cout, 9, w.shape[1])           # [co, tap, ci]
        parts = [pack3(wp[:, :, :c0])]
        if x1 is not None:
            s1h, s1l = nhwc_pair(x1)
            d.src1, d.src1_lo, d.c1, d.h1, d.w1 = s1h.data_ptr(), s1l.data_ptr(), x1.shape[1], x1.shape[2], x1.shape[3]
            d.off_y, d.off_x = off
            parts.append(pack3(wp[:, :, c0:]))
        packed = torch.cat(parts, dim=2).reshape(cout, -1)
        n_total, b, d.taps = cout, bias.contiguous().to(device), 9
    packed = packed.contiguous().to(device)
    keep += [packed, b]
    d.wpack, d.bias, d.n_total, d.mode, d.relu = packed.data_ptr(), b.data_ptr(), n_total, mode, int(relu)
    d.N, d.H, d.W = n, h, wd
    out = {}

    def alloc(shape):
        a = torch.full(shape, float("nan"), dtype=torch.bfloat16, device=device)
        b2 = torch.full(shape, float("nan"), dtype=torch.bfloat16, device=device)
        return a, b2
    if mode in (E.EPI_STORE, E.EPI_STORE_POOL):
        out["dst"] = alloc((n, h, wd, n_total))
        d.dst, d.dst_lo = out["dst"][0].data_ptr(), out["dst"][1].data_ptr()
        if mode == E.EPI_STORE_POOL:
            out["pool"] = alloc((n, h // 2, wd // 2, n_total))
            d.dst_pool, d.dst_pool_lo = out["pool"][0].data_ptr(), out["pool"][1].data_ptr()
    elif mode == E.EPI_CONVT:
        out["dst"] = alloc((n, 2 * h, 2 * wd, n_total // 4))
        d.dst, d.dst_lo = out["dst"][0].data_ptr(), out["dst"][1].data_ptr()
    else:
        hw = head_w.contiguous().to(device)
        hb = head_b.contiguous().to(device)
        keep += [hw, hb]
        of = torch.full((n, head_w.shape[0], h, wd), float("nan"), dtype=torch.float32, device=device)
        d.head_w, d.head_b, d.n_classes, d.out_f32 = hw.data_ptr(), hb.data_ptr(), head_w.shape[0], of.data_ptr()
        out["f32"] = of
    with torch.cuda.device(device):
        E.check(E.lib().fiConvGemm(C.byref(d), E.current_stream()))
        torch.cuda.synchronize()
    res = {}
    for k, v in out.items():
        res[k] = (from_nhwc(v[0]) + from_nhwc(v[1])) if isinstance(v, tuple) else v.cpu()
    return res


def run_conv(device, x, w, bias, *, relu=True, mode=E.EPI_STORE, x1=None, off=(0, 0), head_w=None, head_b=None,
             want_u8=False):
    """x: fp32 NCHW (CPU), w: conv weight fp32 (already BN-folded), returns dict of CPU fp32 NCHW tensors."""
    n, c0, h, wd = x.shape
    xs = to_nhwc_bf16(x, device)
    d = E.ConvDesc()
    d.src0, d.c0 = xs.data_ptr(), c0
    keep = [xs]
    if x1 is not None:
        x1s = to_nhwc_bf16(x1, device)
        keep.append(x1s)
        d.src1, d.c1, d.h1, d.w1 = x1s.data_ptr(), x1.shape[1], x1.shape[2], x1.shape[3]
        d.off_y, d.off_x = off
    if mode == E.EPI_CONVT:
        wp = pack_convT(w).to(device)
        n_total = 4 * w.shape[1]
        b = bias.repeat(4).contiguous().to(device)
        d.taps = 1
    else:
        wp = pack_conv3x3(w).to(device)
        n_total = w.shape[0]
        b = bias.contiguous().to(device)
        d.taps = 9
    keep += [wp, b]
    d.wpack, d.bias, d.n_total, d.mode, d.relu = wp.data_ptr(), b.data_ptr(), n_total, mode, int(relu)
    d.N, d.H, d.W = n, h, wd
    out = {}
    if mode in (E.EPI_STORE, E.EPI_STORE_POOL):
        dst = torch.full((n, h, wd, n_total), float("nan"), dtype=torch.bfloat16, device=device)
        d.dst = dst.data_ptr()
        out["dst"] = dst
        if mode == E.EPI_STORE_POOL:
            pool = torch.full((n, h // 2, wd // 2, n_total), float("nan"), dtype=torch.bfloat16, device=device)
            d.dst_pool = pool.data_ptr()
            out["pool"] = pool
    elif mode == E.EPI_CONVT:
        dst = torch.full((n, 2 * h, 2 * wd, n_total // 4), float("nan"), dtype=torch.bfloat16, device=device)
        d.dst = dst.data_ptr()
        out["dst"] = dst
    else:
        hw = head_w.contiguous().to(device)
        hb = head_b.contiguous().to(device)
        keep += [hw, hb]
        ncls = head_w.shape[0]
        of = torch.full((n, ncls, h, wd), float("nan"), dtype=torch.float32, device=device)
        d.head_w, d.head_b, d.n_classes, d.out_f32 = hw.data_ptr(), hb.data_ptr(), ncls, of.data_ptr()
        out["f32"] = of
        if want_u8:
            ou = torch.zeros((n, ncls, h, wd), dtype=torch.uint8, device=device)
            d.out_u8 = ou.data_ptr()
            out["u8"] = ou
    with torch.cuda.device(device):
        E.check(E.lib().fiConvGemm(C.byref(d), E.current_stream()))
        torch.cuda.synchronize()
    res = {}
    for k, v in out.items():
        res[k] = from_nhwc(v) if v.dtype == torch.bfloat16 else v.cpu()
    return res


def ref_conv3x3(x, w, bias, relu=True, x1=None, off=(0, 0)):
    """fp32 CPU reference on bf16-rounded operands (the kernel's exact inputs)."""
    xin = bf16_round(x)
    if x1 is not None:
        h, wd = x.shape[2], x.shape[3]
        p = F.pad(bf16_round(x1), [off[1], wd - x1.shape[3] - off[1], off[0], h - x1.shape[2] - off[0]])
        xin = torch.cat([xin, p], dim=1)
    y = F.conv2d(xin.double(), bf16_round(w).double(), bias.double(), padding=1).float()
    return F.relu(y) if relu else y
