"""Closed-form golden vectors for compute_psnr / compute_ssim (reference model/evaluation.py:194-218 =
evaluation_simple.py:103-109, which delegate to scikit-image) -> tests/golden/metrics_golden.json.

scikit-image is not installed here, is not vendored by the reference, and the reference holds no golden SSIM/PSNR value,
so the metrics oracle cannot be pinned against a run of the real dependency. It is pinned instead against images whose
SSIM has an exact rational value DERIVED BY HAND from scikit-image's published definition
(metrics/_structural_similarity.py, defaults: 7x7 uniform window, sample covariance, K1=0.01, K2=0.03, data_range=255,
mean over the image cropped by 3 pixels):

    ux = Sx/49, vx = (49/48)(Sxx/49 - ux^2) = (49 Sxx - Sx^2)/2352, vxy = (49 Sxy - Sx Sy)/2352
    S  = (2 ux uy + C1)(2 vxy + C2) / ((ux^2 + uy^2 + C1)(vx + vy + C2)),  C1 = 6.5025, C2 = 58.5225

The trick: an image that is periodic with period 7 in both directions has the SAME five window sums in every 7x7
window (each window holds exactly one period), so its SSIM is the single value S of those sums; a two-level window
with n_a pixels at level a and n_b at level b has vx = n_a n_b (a-b)^2 / 2352. The derivations below are carried out
with integers (asserted), the final value with exact fractions. Neither scipy.ndimage.uniform_filter (what the oracle
uses) nor cumulative sums (what its cross-check uses) appear here.

    python oracle/make_metrics_golden.py        # rewrites tests/golden/metrics_golden.json
"""
import json
import math
from fractions import Fraction as Fr
from pathlib import Path

import numpy as np

C1 = Fr(65025, 10000)      # (0.01 * 255)^2
C2 = Fr(585225, 10000)     # (0.03 * 255)^2


def S(ux, uy, vx, vy, vxy):
    return (2 * ux * uy + C1) * (2 * vxy + C2) / ((ux * ux + uy * uy + C1) * (vx + vy + C2))


def tile7(tile, h, w):
    """h x w image, periodic with period 7 in both directions."""
    t = np.asarray(tile, dtype=np.uint8)
    assert t.shape == (7, 7)
    return np.tile(t, (h // 7 + 1, w // 7 + 1))[:h, :w]


def rows():
    out = []

    def add(name, pred, target, ssim, psnr, note):
        out.append({"name": name, "shape": list(pred.shape), "pred": pred.tolist(), "target": target.tolist(),
                    "ssim": None if ssim is None else float(ssim),
                    "ssim_fraction": None if ssim is None else f"{Fr(ssim).numerator}/{Fr(ssim).denominator}",
                    "psnr": psnr, "note": note})

    # A. impulse lattice against a constant. x tile: 100 everywhere, one pixel 149; y = 100.
    #    Sx = 48*100 + 149 = 4949 -> ux = 101; Sxx = 48*10^4 + 149^2 = 502201 -> 49 Sxx - Sx^2 = 24607849 - 24492601
    #    = 115248 -> vx = 115248/2352 = 49. y constant: uy = 100, vy = 0, Sxy = 100 Sx -> vxy = 0.
    assert 48 * 100 + 149 == 4949 and 4949 == 49 * 101
    assert 49 * (48 * 10 ** 4 + 149 ** 2) - 4949 ** 2 == 115248 and 115248 == 2352 * 49
    tA = np.full((7, 7), 100, np.uint8)
    tA[2, 4] = 149
    sA = S(Fr(101), Fr(100), Fr(49), Fr(0), Fr(0))
    # MSE = 49^2/49 per period; on a 21x28 image (whole periods) exactly 49 -> PSNR = 10 log10(65025/49)
    add("impulse_lattice_vs_constant_21x28", tile7(tA, 21, 28), np.full((21, 28), 100, np.uint8), sA,
        10 * math.log10(65025 / 49), "every window: ux=101, vx=49, uy=100, vy=vxy=0")

    # A'. the same statistics from ONE impulse: a 13x13 constant image with a single pixel 149 at the centre — each of
    #     the 7x7 valid windows contains row 6 and column 6, hence the impulse. MSE = 49^2/169.
    xA1 = np.full((13, 13), 100, np.uint8)
    xA1[6, 6] = 149
    add("single_impulse_13x13", xA1, np.full((13, 13), 100, np.uint8), sA, 10 * math.log10(65025 * 169 / 2401),
        "49 windows, all containing the impulse: same S as the lattice")

    # B. two-level stripes, period 7: 3 columns at level a, 4 at level b -> every window has 21 a's and 28 b's.
    #    x: (50, 200), y: (60, 180): ux = 6650/49 = 950/7, uy = 6300/49 = 900/7,
    #    vx = 21*28*150^2/2352 = 5625, vy = 21*28*120^2/2352 = 3600, vxy = 21*28*(-150)(-120)/2352 = 4500.
    assert 21 * 50 + 28 * 200 == 6650 and 21 * 60 + 28 * 180 == 6300
    assert 21 * 28 * 150 ** 2 == 2352 * 5625 and 21 * 28 * 120 ** 2 == 2352 * 3600 and 21 * 28 * 150 * 120 == 2352 * 4500
    row_x = [50, 50, 50, 200, 200, 200, 200]
    row_y = [60, 60, 60, 180, 180, 180, 180]
    sB = S(Fr(950, 7), Fr(900, 7), Fr(5625), Fr(3600), Fr(4500))
    # per period (7 px of a row): 3*(10)^2 + 4*(20)^2 = 1900 -> MSE = 1900/7 on whole periods (width 35)
    add("stripes_correlated_16x35", tile7([row_x] * 7, 16, 35), tile7([row_y] * 7, 16, 35), sB,
        10 * math.log10(65025 * 7 / 1900), "21/28 two-level windows, positively correlated")

    # C. the same stripes against the INVERTED pattern y: (180, 60): uy = (21*180 + 28*60)/49 = 5460/49 = 780/7,
    #    vy = 3600, vxy = 21*28*(-150)(+120)/2352 = -4500 -> negative structure term, S < 0.
    assert 21 * 180 + 28 * 60 == 5460 and 5460 * 7 == 780 * 49
    row_yi = [180, 180, 180, 60, 60, 60, 60]
    sC = S(Fr(950, 7), Fr(780, 7), Fr(5625), Fr(3600), Fr(-4500))
    assert sC < 0
    # per period: 3*(130)^2 + 4*(140)^2 = 129100 -> MSE = 129100/7
    add("stripes_anticorrelated_9x21_w_not_multiple_of_4", tile7([row_x] * 7, 9, 21), tile7([row_yi] * 7, 9, 21), sC,
        10 * math.log10(65025 * 7 / 129100), "negative SSIM")

    # D. brightness shift: y = x + 10 on the impulse lattice. vx = vy = vxy = 49 -> the contrast/structure factor is 1:
    #    S = (2*101*111 + C1)/(101^2 + 111^2 + C1). MSE = 100 -> PSNR = 20 log10(25.5).
    sD = (2 * Fr(101) * 111 + C1) / (Fr(101) ** 2 + Fr(111) ** 2 + C1)
    assert sD == S(Fr(101), Fr(111), Fr(49), Fr(49), Fr(49))
    xD = tile7(tA, 23, 11)
    add("shift_by_10_23x11_non_square", xD + 10, xD, sD, 20 * math.log10(25.5), "pure luminance term")

    # E. 2-periodic checkerboard (0 / 255) against the constant 128 on 13x14: 7x8 = 56 windows; a window whose corner is
    #    colour p holds 25 p's and 24 q's, and the two classes alternate -> 28 windows each.
    #    class 1 (25 zeros): ux = 24*255/49; class 2: ux = 25*255/49; both: vx = 25*24*255^2/2352; uy = 128, vy = vxy = 0.
    vxE = Fr(25 * 24 * 255 ** 2, 2352)
    s1 = S(Fr(24 * 255, 49), Fr(128), vxE, Fr(0), Fr(0))
    s2 = S(Fr(25 * 255, 49), Fr(128), vxE, Fr(0), Fr(0))
    yy, xx = np.mgrid[0:13, 0:14]
    board = (((yy + xx) % 2) * 255).astype(np.uint8)
    # 91 zeros (|0-128|^2 = 16384) and 91 pixels of 255 (127^2 = 16129) -> MSE = (16384 + 16129)/2
    add("checkerboard_vs_128_13x14", board, np.full((13, 14), 128, np.uint8), (s1 + s2) / 2,
        10 * math.log10(65025 * 2 / (16384 + 16129)), "two window classes, 28 windows each")

    # F. constants (variances vanish): S = (2ab + C1)/(a^2 + b^2 + C1); single window (7x7).
    add("constants_100_110_7x7_single_window", np.full((7, 7), 100, np.uint8), np.full((7, 7), 110, np.uint8),
        (2 * Fr(100) * 110 + C1) / (Fr(100) ** 2 + Fr(110) ** 2 + C1), 20 * math.log10(25.5), "one window")

    # G. identical images: SSIM exactly 1, PSNR +inf (scikit-image returns inf for a zero MSE).
    add("identical_12x13", board[:12, :13].repeat(1, 0), board[:12, :13].copy(), Fr(1), float("inf"), "degenerate")
    return out


def main():
    dst = Path(__file__).resolve().parent.parent / "tests" / "golden" / "metrics_golden.json"
    data = {"comment": "closed-form SSIM/PSNR vectors, see oracle/make_metrics_golden.py for the derivations",
            "errors": [{"shape": [6, 20], "raises": "win_size exceeds image extent (H < 7)"},
                       {"shape": [20, 5], "raises": "win_size exceeds image extent (W < 7)"}],
            "cases": rows()}
    dst.write_text(json.dumps(data).replace("Infinity", '"inf"'))
    for r in data["cases"]:
        print(f"{r['name']:55s} ssim {r['ssim']:+.15f}  psnr {r['psnr']}")
    print("wrote", dst, dst.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
