"""Drop-in for the metric functions of the reference's model/evaluation.py:194-218 (identical in
model/evaluation_simple.py:103-109): compute_psnr / compute_ssim on uint8 2-D arrays, scikit-image semantics
(data_range=255, 7x7 uniform window, sample covariance), evaluated by the fused SSIM+PSNR CUDA kernel.

The report / plotting half of the reference file (matplotlib, seaborn, pandas; :520-1093) is outside the hot path and
is not reproduced; `evaluate_triplets` below is the per-triplet metric loop (:287-340) in batched form.
"""
from __future__ import annotations

import numpy as np
import torch

try:
    from . import _engine as _E
except ImportError:
    import _engine as _E


def _device(device=None):
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cuda"
    return _E.require_cuda(device)


def _as_u8_batch(a):
    a = np.asarray(a)
    if a.dtype != np.uint8:
        raise ValueError("expected uint8 images (the reference passes postprocess_image output / cv2 frames)")
    if a.ndim == 2:
        a = a[None]
    if a.ndim != 3:
        raise ValueError("expected a 2-D image or a batch [N,H,W] of 2-D images")
    return np.ascontiguousarray(a)


def compute_metrics(pred, target, device=None):
    """(psnr, ssim) float64 arrays of shape [N] for uint8 batches [N,H,W] (or single 2-D images)."""
    dev = _device(device)
    p, t = _as_u8_batch(pred), _as_u8_batch(target)
    if p.shape != t.shape:
        raise ValueError("Input images must have the same dimensions.")
    out = _E.ssim_psnr_u8(torch.from_numpy(p).to(dev), torch.from_numpy(t).to(dev)).cpu().numpy()
    return out[:, 0], out[:, 1]


def compute_psnr(pred, target):
    """skimage.metrics.peak_signal_noise_ratio(target, pred, data_range=255) — reference model/evaluation.py:194-205."""
    return float(compute_metrics(pred, target)[0][0])


def compute_ssim(pred, target):
    """skimage.metrics.structural_similarity(target, pred, data_range=255) — reference model/evaluation.py:207-218."""
    return float(compute_metrics(pred, target)[1][0])


def evaluate_triplets(interpolator, triplets, batch=8):
    """Per-triplet loop of reference model/evaluation.py:287-340 / evaluation_simple.py:146-200 in batched form:
    triplets = [(frame1_u8, ground_truth_u8, frame2_u8), ...] of equal-sized grey images. Returns the
    evaluation_simple.py:220-242 result schema for the 'unet' and 'linear' methods (the Farneback optical-flow
    baseline of :76-101 is CPU cv2 code outside this path)."""
    res = {"unet": {"psnr": [], "ssim": []}, "linear": {"psnr": [], "ssim": []}}
    for i in range(0, len(triplets), batch):
        chunk = triplets[i:i + batch]
        f1 = [c[0] for c in chunk]
        gt = np.stack([c[1] for c in chunk])
        f2 = [c[2] for c in chunk]
        pred = np.stack(interpolator._forward_pairs(f1, f2))
        lin = ((np.stack(f1).astype(np.float32) + np.stack(f2).astype(np.float32)) / 2).astype(np.uint8)
        for name, img in (("unet", pred), ("linear", lin)):
            ps, ss = compute_metrics(img, gt, interpolator.device)
            res[name]["psnr"] += [float(v) for v in ps]
            res[name]["ssim"] += [float(v) for v in ss]
    out = {"methods": {}, "num_triplets": len(triplets)}
    for name, r in res.items():
        finite = [v for v in r["psnr"] if np.isfinite(v)]
        out["methods"][name] = {
            "avg_psnr": float(np.mean(finite)) if finite else float("inf"),
            "std_psnr": float(np.std(finite)) if finite else 0.0,
            "avg_ssim": float(np.mean(r["ssim"])) if r["ssim"] else 0.0,
            "std_ssim": float(np.std(r["ssim"])) if r["ssim"] else 0.0,
            "psnr_values": r["psnr"], "ssim_values": r["ssim"],
        }
    return out
