"""CPU tier: the drop-in keeps the reference's Python surface (names, signatures, CLI flags, HTTP routes and their
validation) — SURVEY.md §8b. Nothing here computes on a GPU."""
import inspect

import pytest


def test_unet_module_surface_and_state_dict_schema():
    from model import unet
    for name in ("DoubleConv", "Down", "Up", "OutConv", "UNet", "FrameInterpolationUNet", "count_parameters"):
        assert hasattr(unet, name)
    assert str(inspect.signature(unet.UNet.__init__)) == "(self, n_channels=2, n_classes=1, bilinear=False)"
    assert str(inspect.signature(unet.FrameInterpolationUNet.__init__)) == "(self, bilinear=False)"
    assert list(inspect.signature(unet.FrameInterpolationUNet.forward).parameters) == ["self", "frame1", "frame2"]
    assert str(inspect.signature(unet.DoubleConv.__init__)) == "(self, in_channels, out_channels, mid_channels=None)"
    assert str(inspect.signature(unet.Up.__init__)) == "(self, in_channels, out_channels, bilinear=True)"
    from oracle import unet_oracle as O
    for bilinear, n_params, n_tensors in ((False, 31037057, 118), (True, 17262401, 110)):
        m = unet.FrameInterpolationUNet(bilinear=bilinear)
        assert unet.count_parameters(m) == n_params
        sd = m.state_dict()
        assert len(sd) == n_tensors
        assert list(sd.keys()) == list(O.init_state_dict(0, 2, 1, bilinear).keys())
        m.load_state_dict(O.init_state_dict(0, 2, 1, bilinear))  # strict
    assert unet.count_parameters(unet.UNet(6, 3)) == 31039491


def test_inference_module_surface():
    from model import inference as inf
    sigs = {
        "preprocess_image": ["image_path", "target_size"],
        "postprocess_image": ["tensor"],
        "load_model": ["model_path", "device"],
        "interpolate_frames": ["model", "frame1", "frame2", "device"],
        "generate_multiple_intermediate_frames": ["model", "frame1", "frame2", "num_intermediate", "device"],
        "create_smooth_transition_frames": ["frame1", "frame2", "num_intermediate"],
        "save_frames_as_video": ["frames", "output_path", "fps"],
    }
    for fn, params in sigs.items():
        assert list(inspect.signature(getattr(inf, fn)).parameters) == params, fn
    assert inspect.signature(inf.preprocess_image).parameters["target_size"].default == (256, 256)
    assert inspect.signature(inf.save_frames_as_video).parameters["fps"].default == 30
    fi = inf.FrameInterpolator
    assert list(inspect.signature(fi.interpolate_frames).parameters) == ["self", "frame1", "frame2"]
    assert list(inspect.signature(fi.interpolate_video).parameters)[:4] == ["self", "input_path", "output_path", "factor"]


def test_inference_cli_flags_and_error_exit_code(tmp_path, capsys):
    from model import inference as inf
    with pytest.raises(SystemExit):
        inf.main([])  # --frame1/--frame2 are required, like the reference
    rc = inf.main(["--frame1", str(tmp_path / "a.png"), "--frame2", str(tmp_path / "b.png"), "--model",
                   str(tmp_path / "none.pth"), "--num-intermediate", "3", "--fps", "24", "--save-comparison",
                   "--device", "auto", "--output", str(tmp_path / "o.png")])
    assert rc == 1 and "Error during inference" in capsys.readouterr().out


def test_linear_baseline_matches_reference_formula():
    import torch
    from model.inference import create_smooth_transition_frames
    a, b = torch.zeros(1, 1, 2, 2), torch.ones(1, 1, 2, 2)
    fr = create_smooth_transition_frames(a, b, 3)
    assert [float(f.mean()) for f in fr] == [0.25, 0.5, 0.75]


def test_main_cli_surface(capsys, tmp_path):
    import shlex
    import torch
    import main as cli
    p = cli.build_parser()
    # the command lines the reference documents (README.md:75-111), verbatim after `python main.py`
    readme = {
        "train --data-dir data/train --epochs 100 --batch-size 16": dict(command="train", batch_size=16, lr=0.001),
        "train --data-dir data/train --epochs 200 --lr 0.0001": dict(command="train", epochs=200, lr=0.0001, batch_size=8),
        "infer --frame1 frame1.jpg --frame2 frame2.jpg --output result.jpg": dict(model="best_model.pth", device="auto"),
        "infer --frame1 frame1.jpg --frame2 frame2.jpg --output result.jpg --model my_model.pth": dict(model="my_model.pth"),
        "video --input video.mp4 --output interpolated.mp4 --factor 2": dict(model="best_model.pth", factor=2, gpus=None),
        "video --input video.mp4 --output interpolated.mp4 --factor 4": dict(factor=4, device="auto"),
        "serve --host 0.0.0.0 --port 8000": dict(port=8000, reload=False),
        "serve --host 0.0.0.0 --port 8000 --reload": dict(reload=True),
        "info --model best_model.pth": dict(model="best_model.pth"),
        "serve": dict(host="0.0.0.0", port=8000),
    }
    for line, expect in readme.items():
        a = p.parse_args(shlex.split(line))
        for k, v in expect.items():
            assert getattr(a, k) == v, (line, k)
    with pytest.raises(SystemExit):  # reference main.py:52: --output is required on `infer`
        p.parse_args(["infer", "--frame1", "a", "--frame2", "b"])
    assert p.parse_args(["video", "--input", "i", "--output", "o", "--gpus", "8"]).gpus == 8
    # `info`: reference main.py:139-158 (it raises AttributeError before getting there: no --device on `info`)
    assert cli.main(["info", "--model", str(tmp_path / "missing.pth")]) == 0
    assert "Model file not found" in capsys.readouterr().out
    from model.unet import FrameInterpolationUNet
    ck = tmp_path / "best_model.pth"
    torch.save({"epoch": 7, "model_state_dict": FrameInterpolationUNet(bilinear=True).state_dict(), "train_loss": 0.125,
                "val_loss": 0.25}, ck)
    assert cli.main(["info", "--model", str(ck)]) == 0
    out = capsys.readouterr().out
    assert "Epoch: 7" in out and "Training Loss: 0.125000" in out and "Validation Loss: 0.250000" in out
    assert "17,262,401" in out
    torch.save(FrameInterpolationUNet(bilinear=False).state_dict(), ck)     # bare state dict, ConvT decoder
    assert cli.main(["info", "--model", str(ck)]) == 0
    assert "31,037,057" in capsys.readouterr().out


def test_http_routes_and_validation():
    from fastapi.testclient import TestClient
    from api.app import app
    c = TestClient(app)
    assert c.get("/health").json()["status"] == "healthy"
    assert "POST /interpolate" in c.get("/").json()["endpoints"]
    png = b"\x89PNG\r\n\x1a\n" + b"0" * 16
    files = {"frame1": ("a.png", png, "image/png"), "frame2": ("b.png", png, "image/png")}
    assert c.post("/interpolate", files=files, data={"num_intermediate": "11", "fps": "30"}).status_code == 400
    assert c.post("/interpolate", files=files, data={"num_intermediate": "3", "fps": "5"}).status_code == 400
    bad = {"frame1": ("a.txt", b"x", "text/plain"), "frame2": ("b.png", png, "image/png")}
    assert c.post("/interpolate", files=bad, data={"num_intermediate": "3", "fps": "30"}).status_code == 400
    assert c.post("/interpolate", data={"num_intermediate": "3"}).status_code == 422  # missing files


def test_metric_functions_exist_with_reference_argument_order():
    from model import evaluation, evaluation_simple
    for mod in (evaluation, evaluation_simple):
        assert list(inspect.signature(mod.compute_psnr).parameters) == ["pred", "target"]
        assert list(inspect.signature(mod.compute_ssim).parameters) == ["pred", "target"]
        assert list(inspect.signature(mod.linear_interpolation_baseline).parameters) == ["frame1", "frame2"]
        assert list(inspect.signature(mod.optical_flow_interpolation_baseline).parameters) == ["frame1_np", "frame2_np"]
        assert list(inspect.signature(mod.load_test_triplets).parameters) == ["test_dir"]
    # reference model/evaluation.py:264 / evaluation_simple.py:134
    lead = ["model", "test_triplets", "device", "save_results", "output_dir"]
    assert list(inspect.signature(evaluation.evaluate_model).parameters)[:5] == lead
    assert list(inspect.signature(evaluation_simple.evaluate_model_simple).parameters)[:5] == lead


def test_load_test_triplets_and_host_baseline(tmp_path):
    import cv2
    import numpy as np
    from model import evaluation
    for video, count in (("clip_b", 5), ("clip_a", 3), ("short", 2)):
        d = tmp_path / video
        d.mkdir()
        for i in range(count):
            img = np.zeros((40, 48), np.uint8)
            cv2.circle(img, (10 + 4 * i, 20), 6, 255, -1)
            cv2.imwrite(str(d / f"frame_{i:03d}.png"), img)
    (tmp_path / "notes.txt").write_text("x")
    trips = evaluation.load_test_triplets(str(tmp_path))
    assert len(trips) == 3 + 1 + 0
    t = next(t for t in trips if t["video_name"] == "clip_b" and t["triplet_id"] == 1)
    assert (t["frame_t0"], t["ground_truth"], t["frame_t1"]) == ("frame_001.png", "frame_002.png", "frame_003.png")
    # Farneback baseline: host cv2 code; like the reference it samples frame 1 at x + flow/2 (a half-flow warp)
    a = cv2.imread(str(tmp_path / "clip_b" / "frame_000.png"), 0)
    b = cv2.imread(str(tmp_path / "clip_b" / "frame_002.png"), 0)
    mid = evaluation.optical_flow_interpolation_baseline(a, b)
    assert mid.shape == a.shape and mid.dtype == np.uint8
    cx = lambda im: float((im.astype(np.float64) * np.arange(im.shape[1])).sum() / im.sum())  # noqa: E731
    assert 0.1 < abs(cx(mid) - cx(a)) < abs(cx(b) - cx(a))
    # summary / JSON writers accept the result schema
    res = evaluation._summarise(2, {"linear": {"psnr": [30.0, 32.0], "ssim": [0.9, 0.8]}},
                                {"linear": [{"triplet_id": 0}, {"triplet_id": 1}]}, ["linear"])
    assert res["metrics_by_method"]["linear"]["average_psnr"] == 31.0 and res["successful_evaluations"] == 2
    evaluation.print_evaluation_summary(res)
    evaluation.save_evaluation_results(res, str(tmp_path / "r.json"))
    import json
    assert json.load(open(tmp_path / "r.json"))["metrics_by_method"]["linear"]["max_ssim"] == 0.9
