import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "ai-based-frame-interpolation_b200"
for p in (str(PKG), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)


def _ensure_library():
    """The C-ABI library is a build artefact (git-ignored). Build it when it is missing so that a fresh checkout can
    run the CPU tier directly (nvcc cross-compiles sm_100a without a GPU); a GPU box receives the prebuilt .so."""
    import shutil
    import subprocess
    lib = PKG / "libfi_b200.so"
    if not lib.exists() and shutil.which("nvcc") or (not lib.exists() and Path("/usr/local/cuda/bin/nvcc").exists()):
        subprocess.run(["make", "-C", str(PKG / "csrc"), "-j4"], check=True, stdout=subprocess.DEVNULL)


_ensure_library()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
