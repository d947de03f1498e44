"""Parity at the sizes BASELINE.json names and bench.py measures (reference model/unet.py:84-95, pad path :49-53):

  config 1   1 x 256x256 pair              config 2   32 x 256x256 pairs
  config 3   4 x 1080x1920 pairs — exactly the step bench.py times (synthetic clip frames, u8 in -> u8 out, ConvT, bf16)
  config 4   1 x 2160x3840 pair

each on the default-init fixture (what BASELINE words as "random-init") and on the stressed fixture (SURVEY.md A.6 ii:
randomised BatchNorm statistics + rescaled head, output spans both clamps), against the fp32 CPU oracle, plus the CUDA path
directly against the bytes the UNMODIFIED reference produced (tests/golden/unet_golden.npz) without the oracle in between.

Stated tolerances (BASELINE.json north_star, bf16 path): max |err| <= 2e-2 in [0,1] pixel units (logits are in [-1,1]
units, hence the /2), PSNR >= 45 dB, u8 frames within 6 grey levels (2e-2*255 = 5.1, +1 for the truncating cast). The
default-init output is nearly constant, so relative L2 bounds on the output and on tapped layers are added.
"""
import numpy as np
import pytest
import torch

from golden_utils import CASES as GOLDEN_CASES
from golden_utils import golden_case
from model import _engine as E
from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu

TOL_PIXEL = 2e-2     # max abs error, [0,1] pixel units
TOL_PSNR = 45.0      # dB, [0,1] pixel units
TOL_U8 = 6           # grey levels
TOL_REL = 2e-2       # relative L2, output and taps


def psnr_unit(a, b):
    mse = float(((a - b) ** 2).mean())
    return 99.0 if mse == 0 else 10 * np.log10(1.0 / mse)


def fixture_weights(kind, bilinear, x_first):
    sd = O.init_state_dict(0, 2, 1, bilinear)
    if kind == "stressed":
        # the head is rescaled so that the oracle output on the (first pair of the) actual input has
        # mean 0 and std 0.5 — on a crop the scale would be off by the crop's statistics
        sd = O.calibrate_head(O.stress_state_dict(sd, seed=1), x_first)
    return sd


def clip_frames(n_pairs, h, w, like_bench):
    if like_bench:
        import bench
        assert (bench.H, bench.W) == (h, w)
        return bench.synthetic_frames(n_pairs + 1)          # [F,1,H,W] u8: the frames bench.py's step consumes
    rs = np.random.RandomState(h * 7 + w + n_pairs)
    return rs.randint(0, 256, size=(n_pairs + 1, 1, h, w)).astype(np.uint8)


def oracle_forward(sd, f1, f2, taps_first=None):
    """One oracle forward per pair (bounded host memory at 1080p / 4K); taps of the first pair only."""
    outs = []
    for i in range(f1.shape[0]):
        x = torch.cat([O.preprocess_u8(f1[i:i + 1]), O.preprocess_u8(f2[i:i + 1])], 1)
        outs.append(O.unet_forward(sd, x, taps_first if i == 0 else None))
    return torch.cat(outs, 0)


# (name, pairs, H, W, frames, weights, max-abs tolerance). Frames: "bench" = the synthetic clip bench.py feeds (smooth
# gradient + disc + mild noise), "uniform" = uniform random u8. The BASELINE-stated 2e-2 bar applies to
# the reference's random-init fixture and to the SURVEY-defined stressed fixture (uniform inputs). The last row is an
# extra, harsher combination: stressed weights whose head is calibrated to std 0.5 on the LOW-CONTRAST bench clip, which
# amplifies the bf16 noise relative to the signal (measured: rel L2 1.4e-2, PSNR 49.1 dB, max 2.3e-2 = a 6.5 sigma tail
# over 33 M pixels, u8 within 5 grey levels); its max-abs bound is 3e-2, PSNR / u8 / relative bounds unchanged.
CASES = [
    ("config1_1x256", 1, 256, 256, "uniform", "default", TOL_PIXEL),
    ("config1_1x256", 1, 256, 256, "uniform", "stressed", TOL_PIXEL),
    ("config2_32x256", 32, 256, 256, "uniform", "default", TOL_PIXEL),
    ("config2_32x256", 32, 256, 256, "uniform", "stressed", TOL_PIXEL),
    ("config3_4x1080p", 4, 1080, 1920, "bench", "default", TOL_PIXEL),      # exactly the step bench.py times
    ("config3_4x1080p", 4, 1080, 1920, "uniform", "stressed", TOL_PIXEL),
    ("config4_1x4k", 1, 2160, 3840, "uniform", "default", TOL_PIXEL),
    ("config4_1x4k", 1, 2160, 3840, "uniform", "stressed", TOL_PIXEL),
    ("config3_4x1080p", 4, 1080, 1920, "bench", "stressed", 3e-2),
]


@pytest.mark.parametrize("name,n,h,w,frames_kind,kind,tol_pixel", CASES,
                         ids=[f"{c[0]}-{c[4]}-{c[5]}" for c in CASES])
def test_parity_at_baseline_sizes(cuda_device, name, n, h, w, frames_kind, kind, tol_pixel):
    fr = clip_frames(n, h, w, frames_kind == "bench")
    f1, f2 = fr[:-1], fr[1:]
    sd = fixture_weights(kind, False, torch.cat([O.preprocess_u8(f1[:1]), O.preprocess_u8(f2[:1])], 1))
    taps = {}
    ref = oracle_forward(sd, f1, f2, taps)
    if kind == "stressed":
        assert ref.min() < -1.05 and ref.max() > 1.05, "stressed fixture must exercise both clamps"

    net = E.Net(cuda_device, 2, 1, False)
    net.load_state_dict(sd)
    d1, d2 = torch.from_numpy(f1).to(cuda_device), torch.from_numpy(f2).to(cuda_device)
    got_f, got_u = net.forward(d1, d2, want_f32=True, want_u8=True)     # the bench step: u8 frames in, u8 frame out
    got_f, got_u = got_f.cpu(), got_u.cpu().numpy()

    assert got_f.shape == ref.shape == (n, 1, h, w)
    err = (got_f - ref).abs().max().item() / 2
    psnr = psnr_unit(got_f / 2, ref / 2)
    rel = ((got_f - ref).norm() / ref.norm()).item()
    du8 = np.abs(got_u.astype(np.int32) - O.postprocess(ref).astype(np.int32))
    print(f"{name}/{kind}: max|err| {err:.2e} px, PSNR {psnr:.1f} dB, rel L2 {rel:.2e}, u8 max diff {du8.max()}")
    assert err <= tol_pixel, f"max abs pixel error {err}"
    assert psnr >= TOL_PSNR, f"PSNR {psnr:.2f} dB"
    assert rel < TOL_REL, f"output relative L2 error {rel:.4f}"
    assert du8.max() <= TOL_U8, f"u8 frames differ by {du8.max()} grey levels"
    # the fused u8 head is exactly postprocess_image of the fp32 logits the same kernel wrote
    assert np.array_equal(got_u, O.postprocess(got_f))

    # per-layer taps of the first pair (a batch-1 forward gives bit-identical per-image results: test_gpu_unet.py)
    one = net.forward(d1[:1], d2[:1], want_f32=True)[0].cpu()
    assert torch.equal(one[0], got_f[0])
    for layer in ("inc", "down4", "up1.up", "up1", "up3"):
        b = taps[layer]
        a = net.read_activation(layer, 1, max_elems=b.numel())
        assert a.shape == b.shape, (layer, a.shape, b.shape)
        r = ((a - b).norm() / (b.norm() + 1e-12)).item()
        assert r < TOL_REL, f"{name}/{kind} layer {layer}: relative L2 error {r:.4f}"
    net.close()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_cuda_path_against_reference_golden_bytes(cuda_device, name):
    """The CUDA path against outputs of the unmodified reference module (oracle/make_golden.py ran it in the build
    container): weights are rebuilt from the product's own nn.Module skeleton + the fixture (digest-checked against the
    reference's), the forward is the library's — the CPU port is not involved."""
    m, frames, ref = golden_case(name)
    m = m.to(cuda_device)
    u8 = torch.from_numpy(frames).to(cuda_device)
    if hasattr(m, "unet"):
        got = m(u8[:, :1], u8[:, 1:]).cpu().numpy()            # raw u8 frames: normalisation fused in the stem
    else:
        got = m(u8).cpu().numpy()
    assert got.shape == ref.shape
    err = np.abs(got - ref).max() / 2
    rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    print(f"golden {name}: max|err| {err:.2e} px, rel L2 {rel:.2e}")
    assert err <= TOL_PIXEL and rel < TOL_REL
    assert psnr_unit(got / 2, ref / 2) >= TOL_PSNR
    if "stressed" in name:
        m.precision = "fp32"                                     # the <= 1e-3 bar of the fp32-grade path
        x = u8[:, :1], u8[:, 1:]
        got32 = m(*x).cpu().numpy()
        assert np.abs(got32 - ref).max() / 2 <= 1e-3
