"""CPU oracle for compute_psnr / compute_ssim — TEST INFRASTRUCTURE ONLY (see oracle/unet_oracle.py for the rules).

PARITY: pinned by closed-form vectors, NOT by a run of the real dependency. The reference
(model/evaluation.py:194-218, model/evaluation_simple.py:103-109) delegates to scikit-image
(`skimage.metrics.peak_signal_noise_ratio` / `structural_similarity`, unpinned in requirements.txt:8). scikit-image is
neither vendored in /root/reference nor installed in this image, and the reference holds no golden SSIM/PSNR value, so
"parity unpinned" still holds in the strict sense (no output of skimage itself was ever compared). This file restates the
published scikit-image algorithm (metrics/_structural_similarity.py and metrics/simple_metrics.py, defaults: win_size=7,
uniform window via scipy.ndimage.uniform_filter, sample covariance, K1=0.01, K2=0.03, crop (win_size-1)//2) for uint8
2-D inputs with data_range=255. What pins it (tests/test_oracle.py):
  * tests/golden/metrics_golden.json — eight images whose SSIM is an exact rational number derived by hand from that
    definition (oracle/make_metrics_golden.py: 7-periodic impulse lattice, a single centred impulse, correlated and
    anti-correlated two-level stripes with S < 0, a pure brightness shift, a 0/255 checkerboard with two window classes,
    constants, identical images -> 1.0 / +inf; non-square sizes, widths that are not multiples of 4, a single-window
    7x7 image, and the H < 7 / W < 7 error). The oracle agrees with every one to 1e-12;
  * an independent exact-integer window implementation (ssim_u8_integer below) on random images to 1e-12.
The worked SSIM values quoted in scikit-image's own documentation/tests could not be reproduced from memory and are not used.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import uniform_filter


def psnr_u8(pred, target):
    """skimage.metrics.peak_signal_noise_ratio(target, pred, data_range=255) for uint8 arrays."""
    t = np.asarray(target, dtype=np.float64)
    p = np.asarray(pred, dtype=np.float64)
    mse = np.mean((t - p) ** 2, dtype=np.float64)
    if mse == 0:
        return float("inf")
    return float(10 * np.log10((255.0 ** 2) / mse))


def ssim_u8(pred, target):
    """skimage.metrics.structural_similarity(target, pred, data_range=255), 2-D uint8 inputs, default arguments."""
    x = np.asarray(target, dtype=np.float64)
    y = np.asarray(pred, dtype=np.float64)
    if x.ndim != 2 or x.shape != y.shape:
        raise ValueError("expected two 2-D arrays of the same shape")
    win = 7
    if min(x.shape) < win:
        raise ValueError("win_size exceeds image extent")
    npx = win * win
    cov_norm = npx / (npx - 1)  # sample covariance
    ux = uniform_filter(x, size=win)
    uy = uniform_filter(y, size=win)
    uxx = uniform_filter(x * x, size=win)
    uyy = uniform_filter(y * y, size=win)
    uxy = uniform_filter(x * y, size=win)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    c1 = (0.01 * 255) ** 2
    c2 = (0.03 * 255) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
    pad = (win - 1) // 2
    return float(np.mean(s[pad:-pad, pad:-pad], dtype=np.float64))


def ssim_u8_integer(pred, target):
    """Independent implementation: exact integer 7x7 window sums via 2-D cumulative sums, rational form of S."""
    x = np.asarray(target, dtype=np.int64)
    y = np.asarray(pred, dtype=np.int64)

    def box(a):
        c = np.zeros((a.shape[0] + 1, a.shape[1] + 1), dtype=np.int64)
        c[1:, 1:] = a.cumsum(0).cumsum(1)
        return c[7:, 7:] - c[:-7, 7:] - c[7:, :-7] + c[:-7, :-7]

    sx, sy, sxx, syy, sxy = box(x), box(y), box(x * x), box(y * y), box(x * y)
    c1 = (0.01 * 255) ** 2
    c2 = (0.03 * 255) ** 2
    a1 = 2.0 * (sx * sy) + c1 * 2401
    b1 = (sx * sx + sy * sy) + c1 * 2401
    a2 = 2.0 * (49 * sxy - sx * sy) + c2 * 2352
    b2 = (49 * (sxx + syy) - sx * sx - sy * sy) + c2 * 2352
    return float(np.mean((a1 * a2) / (b1 * b2), dtype=np.float64))
