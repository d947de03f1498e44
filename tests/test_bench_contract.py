"""CPU tier: the driver-facing contract of bench.py that can be checked without a GPU — the reference arm prints exactly
ONE JSON line on stdout with the keys the driver reads, on the same `config` object as the b200 arm."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line_on_the_b200_arms_config():
    # one bounded step of ONE 1080p pair through the reference's CPU path (about 10 s on 8 cores)
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--pairs", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, str(ROOT))
    import argparse
    import bench
    assert d["metric"] == bench.METRIC
    assert d["config"] == bench.config_1080p(argparse.Namespace(bilinear=False, pairs=1))
    # when the build container installed the unmodified reference (baseline/install_ref.py), the arm runs it
    if bench.REF_UNET.exists():
        assert d["cpu_baseline"]["kind"] == "reference"


def test_other_workloads_of_the_reference_arm_say_unavailable():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "train"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and "unavailable" in d
