"""CPU tier: the oracle against the golden vectors produced by the unmodified reference (oracle/make_golden.py), and
the restated scikit-image metrics against an independent exact-integer implementation."""
import hashlib
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as M
from oracle import unet_oracle as O

GOLD = np.load(Path(__file__).parent / "golden" / "unet_golden.npz")
CASES = sorted({k.split("/")[0] for k in GOLD.files if k.endswith("/logits")})


def golden_state_dict(name):
    n_ch, n_cls, bil, wrapper, stressed = (int(v) for v in GOLD[name + "/cfg"])
    sd = O.init_state_dict(0, n_ch, n_cls, bool(bil), prefix="unet." if wrapper else "")
    frames = GOLD[name + "/frames"]
    x = O.preprocess_u8(frames)
    if stressed:
        sd = O.calibrate_head(O.stress_state_dict(sd, seed=1), x)
    return sd, x


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    sd, x = golden_state_dict(name)
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.contiguous().numpy().tobytes())
    assert h.digest() == GOLD[name + "/sd_sha256"].tobytes(), "oracle weights differ from the reference's"
    y = O.unet_forward(sd, x).numpy()
    ref = GOLD[name + "/logits"]
    assert y.shape == ref.shape
    # same torch build, same op sequence -> identical; allow fp32 reassociation noise for other torch builds
    assert np.abs(y - ref).max() <= 1e-5


def test_preprocess_postprocess_golden():
    assert np.array_equal(O.preprocess_u8(GOLD["pre/u8"]).numpy(), GOLD["pre/norm"])
    assert np.array_equal(O.postprocess(torch.from_numpy(GOLD["post/in"])), GOLD["post/out"])
    # truncation, not rounding
    assert O.postprocess(torch.tensor([0.999 * 2 - 1]))[0] == 254


def test_flops_table():
    # SURVEY.md §8d / BASELINE.md §3
    assert abs(O.flops_per_forward(1, 256, 256) / 1e9 - 96.26) < 0.01
    assert abs(O.flops_per_forward(1, 1080, 1920) / 1e9 - 3043.75) < 0.01
    assert abs(O.flops_per_forward(1, 256, 256, bilinear=True) / 1e9 - 79.88) < 0.01
    assert abs(O.flops_per_forward(1, 2160, 3840, 6, 3) / 1e9 - 12223.16) < 0.02


@pytest.mark.parametrize("shape,seed", [((7, 7), 0), ((16, 23), 1), ((64, 64), 2), ((135, 240), 3)])
def test_ssim_restatement_vs_integer_windows(shape, seed):
    rs = np.random.RandomState(seed)
    a = rs.randint(0, 256, size=shape).astype(np.uint8)
    b = np.clip(a.astype(np.int32) + rs.randint(-20, 21, size=shape), 0, 255).astype(np.uint8)
    assert abs(M.ssim_u8(a, b) - M.ssim_u8_integer(a, b)) < 1e-12
    assert M.ssim_u8(a, a) == pytest.approx(1.0, abs=1e-15)
    assert M.psnr_u8(a, a) == float("inf")
    mse = np.mean((a.astype(np.float64) - b) ** 2)
    assert M.psnr_u8(a, b) == pytest.approx(10 * np.log10(65025 / mse))


def test_ssim_known_values():
    # constant images: S = (2 ux uy + C1) / (ux^2 + uy^2 + C1) exactly (variances vanish)
    a = np.full((9, 9), 100, np.uint8)
    b = np.full((9, 9), 110, np.uint8)
    c1 = (0.01 * 255) ** 2
    assert M.ssim_u8(a, b) == pytest.approx((2 * 100 * 110 + c1) / (100 ** 2 + 110 ** 2 + c1), rel=1e-12)
    with pytest.raises(ValueError):
        M.ssim_u8(np.zeros((6, 20), np.uint8), np.zeros((6, 20), np.uint8))


# ---------------------------------------------------------------------- closed-form metric vectors (hand-derived)
import json  # noqa: E402

METRICS_GOLD = json.loads((Path(__file__).parent / "golden" / "metrics_golden.json").read_text())


@pytest.mark.parametrize("row", METRICS_GOLD["cases"], ids=lambda r: r["name"])
def test_metrics_oracle_against_closed_forms(row):
    """The restated scikit-image metrics against exact rational SSIM / closed-form PSNR values derived by hand in
    oracle/make_metrics_golden.py (reference call sites model/evaluation.py:194-218)."""
    from fractions import Fraction
    pred, target = np.array(row["pred"], np.uint8), np.array(row["target"], np.uint8)
    assert list(pred.shape) == row["shape"]
    num, den = row["ssim_fraction"].split("/")
    exact = float(Fraction(int(num), int(den)))
    assert exact == row["ssim"]
    assert abs(M.ssim_u8(pred, target) - exact) <= 1e-12
    assert abs(M.ssim_u8_integer(pred, target) - exact) <= 1e-12
    want = float("inf") if row["psnr"] == "inf" else row["psnr"]
    got = M.psnr_u8(pred, target)
    assert got == want if want == float("inf") else abs(got - want) <= 1e-10


def test_metrics_oracle_error_rows():
    for row in METRICS_GOLD["errors"]:
        z = np.zeros(row["shape"], np.uint8)
        with pytest.raises(ValueError):
            M.ssim_u8(z, z)
