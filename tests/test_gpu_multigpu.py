"""GPU tier, needs >= 2 devices (gpurun --gpus 2): the product's multi-GPU split — FrameInterpolator(gpus=N) /
`main.py video --gpus N` (reference call sites main.py:118-129; north_star: "video interpolation shards naturally by
frame pair ... no collective on the inference path"). The sharded result must equal the single-GPU result byte for
byte, a failing worker's range must be re-queued, and a second device in one process must work (the shared-memory
opt-in of every kernel is per device)."""
import numpy as np
import pytest
import torch
import cv2

from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    return 2


def clip(n, h=72, w=104, color=False):
    rs = np.random.RandomState(7)
    yy, xx = np.mgrid[0:h, 0:w]
    out = []
    for i in range(n):
        img = 40 + 60 * (xx / w) + 150 * (((xx - (20 + 5 * i)) ** 2 + (yy - h // 2) ** 2) < 14 ** 2) + rs.randint(0, 9, (h, w))
        img = np.clip(img, 0, 255).astype(np.uint8)
        out.append(np.stack([img, np.roll(img, 3, 0), np.roll(img, 5, 1)], -1) if color else img)
    return np.stack(out)


@pytest.fixture(scope="module")
def checkpoint(tmp_path_factory):
    fr = clip(3)
    x = torch.cat([O.preprocess_u8(fr[0][None, None]), O.preprocess_u8(fr[2][None, None])], 1)
    sd = O.calibrate_head(O.stress_state_dict(O.init_state_dict(0, 2, 1, False), seed=1), x, out_std=0.4)
    p = tmp_path_factory.mktemp("ckpt") / "model.pth"
    torch.save(sd, p)
    return str(p), sd


def test_second_device_in_one_process(two_gpus, checkpoint):
    """Regression for the per-process `configured` flags: every kernel needs its > 48 KB dynamic shared-memory opt-in on
    EACH device it runs on."""
    from model import _engine as E
    _, sd = checkpoint
    fr = clip(4)[:, None]
    outs = []
    for dev in (0, 1, 0):
        net = E.Net(f"cuda:{dev}", 2, 1, False)
        net.load_state_dict(sd)
        outs.append(net.interpolate_clip_host_u8(fr, 2))
        net.close()
        # the library restores the caller's current device on every exit path (a leaked cudaSetDevice would re-target
        # the caller's next "cuda" tensor / FrameInterpolator(..., "cuda"))
        assert torch.cuda.current_device() == 0
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    ref = O.postprocess(O.unet_forward(sd, torch.cat([O.preprocess_u8(fr[:-1]), O.preprocess_u8(fr[1:])], 1)))
    assert np.abs(outs[1].astype(int) - ref.astype(int)).max() <= 6


@pytest.mark.parametrize("color", [False, True])
@pytest.mark.parametrize("factor", [2, 4, 3])
def test_two_gpus_equal_one_gpu_byte_for_byte(two_gpus, checkpoint, color, factor):
    from model.inference import FrameInterpolator
    path, _ = checkpoint
    frames = clip(11, color=color)
    one = FrameInterpolator(path, "cuda", pairs_per_batch=2, gpus=1)
    two = FrameInterpolator(path, "cuda", pairs_per_batch=2, gpus=2)
    assert (one.gpus, two.gpus) == (1, 2)
    a = one.interpolate_sequence(frames, factor)
    b = two.interpolate_sequence(list(frames), factor)
    assert len(a) == len(b) == (len(frames) - 1) * factor + 1
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    pool = two._gpu_pool()
    assert pool.n_alive == 2 and not pool.errors
    assert {r.device.index for r in pool.runners} == {0, 1}
    two.close()


def test_failed_worker_range_is_requeued(two_gpus, checkpoint):
    from model.inference import FrameInterpolator
    path, _ = checkpoint
    frames = clip(9)
    ref = FrameInterpolator(path, "cuda", pairs_per_batch=2).interpolate_clip(frames)
    fi = FrameInterpolator(path, "cuda", pairs_per_batch=2, gpus=[0, 1])
    pool = fi._gpu_pool()

    def fault(widx, first, n):
        if widx == 1:
            raise RuntimeError("injected GPU failure")

    pool.fault_hook = fault
    got = fi.interpolate_clip(frames)
    assert np.array_equal(got, ref)
    assert pool.alive == [True, False] and "injected" in str(pool.errors[0][2])
    assert np.array_equal(fi.interpolate_clip(frames), ref)      # keeps working on the survivor
    # an invalid request (frames below 16x16) is the caller's error: raised as is, nobody is retired
    from model._engine import FiError
    with pytest.raises(FiError, match="smaller than 16x16"):
        fi.interpolate_clip(np.zeros((5, 8, 8), np.uint8))
    assert pool.alive == [True, False]
    pool.fault_hook = lambda widx, first, n: (_ for _ in ()).throw(RuntimeError("injected GPU failure"))
    with pytest.raises(RuntimeError, match="all GPU workers failed"):
        fi.interpolate_clip(frames)
    fi.close()


def test_video_cli_two_gpus_matches_one(two_gpus, checkpoint, tmp_path):
    import main as cli
    path, _ = checkpoint
    frames = clip(14, 96, 128, color=True)
    src = str(tmp_path / "in.mp4")
    wr = cv2.VideoWriter(src, cv2.VideoWriter_fourcc(*"mp4v"), 10.0, (128, 96), True)
    for f in frames:
        wr.write(f)
    wr.release()
    outs = []
    for g in (1, 2):
        dst = str(tmp_path / f"out{g}.mp4")
        assert cli.main(["video", "--input", src, "--output", dst, "--factor", "2", "--model", path, "--gpus", str(g)]) == 0
        cap, got = cv2.VideoCapture(dst), []
        while True:
            ok, fr = cap.read()
            if not ok:
                break
            got.append(fr)
        outs.append(np.stack(got))
    assert outs[0].shape == (27, 96, 128, 3)
    assert np.array_equal(outs[0], outs[1])
