"""CPU tier: the loss / dataset / schedule mirrors of reference model/train.py against the golden vectors generated
from the unmodified reference (oracle/make_train_golden.py)."""
import os

import numpy as np
import pytest
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

from model import train as T


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "train_golden.npz"))


def test_combined_loss_value_and_gradient_match_reference(golden):
    pred = torch.from_numpy(golden["loss_pred"]).requires_grad_(True)
    target = torch.from_numpy(golden["loss_target"])
    loss = T.CombinedLoss()(pred, target)
    loss.backward()
    assert abs(loss.item() - float(golden["loss_value"])) < 1e-6
    assert np.allclose(pred.grad.numpy(), golden["loss_grad"], rtol=1e-4, atol=1e-8)
    assert abs(T.SSIMLoss()(pred.detach(), target).item() - float(golden["ssim_loss_value"])) < 1e-6


def test_ssim_loss_identical_images_is_zero_and_multichannel():
    x = torch.rand(1, 3, 24, 24)
    assert abs(T.SSIMLoss()(x, x).item()) < 1e-6
    per_image = T.SSIMLoss(size_average=False)(torch.rand(4, 1, 16, 16), torch.rand(4, 1, 16, 16))
    assert per_image.shape == (4,)


def test_dataset_builds_triplets(tmp_path):
    cv2 = pytest.importorskip("cv2")
    for video, count in (("a", 4), ("b", 2), ("c", 3)):
        d = tmp_path / video
        d.mkdir()
        for i in range(count):
            cv2.imwrite(str(d / f"f{i:03d}.png"), np.full((20, 30), 40 * i, np.uint8))
    (tmp_path / "stray.txt").write_text("x")
    ds = T.FrameTripletDataset(str(tmp_path))
    assert len(ds) == 2 + 0 + 1   # len-2 triplets per directory (reference model/train.py:108)
    trip = next(t for t in ds.triplets if t["video_dir"].endswith("a") and t["frame_t0"] == "f000.png")
    assert (trip["frame_t1"], trip["ground_truth"]) == ("f002.png", "f001.png")
    f0, f1, gt = ds[ds.triplets.index(trip)]
    assert f0.shape == (1, 256, 256) and f0.dtype == torch.float32
    assert abs(gt.mean().item() - 40 / 255) < 1e-6 and abs(f1.mean().item() - 80 / 255) < 1e-6


def test_plateau_schedule_halves_after_patience():
    class S:
        lr = 1e-4
    s = S()
    sched = T._PlateauSchedule(s)
    ref_opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1e-4)
    ref = torch.optim.lr_scheduler.ReduceLROnPlateau(ref_opt, mode="min", factor=0.5, patience=10)
    for i, m in enumerate([1.0, 0.9] + [0.95] * 25 + [0.5] + [0.6] * 12):
        sched.step(m)
        ref.step(m)
        assert abs(s.lr - ref_opt.param_groups[0]["lr"]) < 1e-12, i
