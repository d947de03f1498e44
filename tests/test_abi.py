"""CPU tier: the C-ABI library builds, loads and exports every symbol include/fi_b200.h declares; compute calls fail
loudly (no fallback) when there is no CUDA device."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

from model import _engine as E

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "fi_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fi[A-Z]\w*)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = E.lib()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in fi_b200.h but not exported"
    assert set(names) == set(E.EXPORTED_SYMBOLS), "ctypes binding and header disagree"
    assert lib.fiVersion() >= 100


def test_struct_layouts_match_header():
    # field order of the ctypes mirrors follows the header (a reorder would silently corrupt calls)
    text = (ROOT / "include" / "fi_b200.h").read_text()
    body = re.search(r"typedef struct fiConvDesc \{(.*?)\} fiConvDesc;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(",")
        fields.append(names[0].split()[-1].lstrip("*"))
        fields += [n.strip().lstrip("*") for n in names[1:]]
    assert fields == [f[0] for f in E.ConvDesc._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    h = C.c_void_p()
    rc = E.lib().fiNetCreate(C.byref(h), 0, 2, 1, 0)
    assert rc == -2 and b"no CPU fallback" in E.lib().fiLastError()
    from model.unet import FrameInterpolationUNet
    m = FrameInterpolationUNet().eval()
    with pytest.raises(E.FiError):
        m(torch.zeros(1, 1, 32, 32), torch.zeros(1, 1, 32, 32))
    with pytest.raises(E.FiError):
        E.ssim_psnr_u8(torch.zeros(8, 8, dtype=torch.uint8), torch.zeros(8, 8, dtype=torch.uint8))


def test_argument_validation_without_gpu():
    assert E.lib().fiNetCreate(None, 0, 2, 1, 0) == -1
    h = C.c_void_p()
    assert E.lib().fiNetCreate(C.byref(h), 0, 99, 1, 0) == -1
    assert E.lib().fiSsimPsnrWorkspaceBytes(0, 10, 10) == 0
    assert E.lib().fiSsimPsnrWorkspaceBytes(2, 2160, 3840) == 2 * 8 * 60 * 16


def test_product_code_never_touches_the_oracle_or_the_reference():
    """oracle/ is test infrastructure: the package, the tools and the measured arm of bench.py must not import it, and
    nothing that runs on the GPU box may read /root/reference."""
    import ast
    import pathlib
    root = pathlib.Path(__file__).resolve().parent.parent

    def oracle_imports(path, skip_functions=()):
        tree = ast.parse(path.read_text())
        hits = []
        for node in ast.walk(tree):
            if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef)) and node.name in skip_functions:
                for sub in ast.walk(node):
                    sub._skipped = True
            mod = None
            if isinstance(node, ast.ImportFrom):
                mod = node.module or ""
            elif isinstance(node, ast.Import):
                mod = ",".join(a.name for a in node.names)
            if mod is not None and "oracle" in mod and not getattr(node, "_skipped", False):
                hits.append((path.name, node.lineno))
        return hits

    files = list((root / "ai-based-frame-interpolation_b200").rglob("*.py")) + list((root / "tools").rglob("*.py"))
    bad = [h for f in files for h in oracle_imports(f)]
    # bench.py: only the CPU-baseline / reference arm may run the oracle
    bad += oracle_imports(root / "bench.py", skip_functions=("cpu_step_fn",))
    assert not bad, bad
    for f in files + [root / "bench.py", root / "__graft_entry__.py"]:
        assert "/root/reference" not in f.read_text(), f
