"""fiWgrad (tensor-core weight gradient, operands read straight from NHWC) against an fp64 reference on the same bf16
operands: both operand roles (M = cout / M = cin), every tile width, concat sources, odd and tiny shapes."""
import pytest
import torch
import torch.nn.functional as F

from model import _engine as E

pytestmark = pytest.mark.gpu

CASES = [
    # n, h, w, c0, c1, cout
    (1, 8, 8, 64, 0, 64),        # 64x64: half-filled M tile
    (2, 16, 24, 128, 0, 64),     # swapped roles (M = cin), W not a multiple of 16
    (2, 12, 20, 64, 0, 128),
    (1, 16, 16, 256, 0, 256),    # N tile 256
    (3, 9, 13, 512, 0, 128),     # odd shape: out-of-bounds pixels in every slab
    (2, 32, 32, 64, 0, 64),
    (2, 16, 16, 64, 64, 64),     # concat [skip | up] -> swapped roles, two sources
    (1, 8, 8, 256, 256, 256),    # concat, N tile 256 spanning both sources
    (2, 2, 2, 512, 0, 512),      # deepest level of a 32x32 training image
    (1, 1, 3, 64, 0, 192),       # m_total not a multiple of 128
]


@pytest.mark.parametrize("n,h,w,c0,c1,cout", CASES)
def test_wgrad_matches_reference(cuda_device, n, h, w, c0, c1, cout):
    cin = c0 + c1
    g = torch.Generator().manual_seed(cin + cout + h)
    x = torch.randn(n, cin, h, w, generator=g).to(torch.bfloat16)
    dz = (torch.randn(n, cout, h, w, generator=g) * 0.1).to(torch.bfloat16)
    lib, st = E.lib(), E.current_stream()
    xs = x.permute(0, 2, 3, 1).contiguous().to(cuda_device)
    x0 = xs[..., :c0].contiguous()
    x1 = xs[..., c0:].contiguous() if c1 else None
    ds = dz.permute(0, 2, 3, 1).contiguous().to(cuda_device)
    prior = torch.randn(9, cout, cin, generator=g).to(cuda_device)   # the call accumulates into dW
    dW = prior.clone()
    E.check(lib.fiWgrad(ds.data_ptr(), x0.data_ptr(), c0, x1.data_ptr() if c1 else None, c1, n, h, w, cout,
                        dW.data_ptr(), st))
    torch.cuda.synchronize()
    wt = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), wt, padding=1).backward(dz.double())
    ref = wt.grad.permute(2, 3, 0, 1).reshape(9, cout, cin).float()
    err = ((dW - prior).cpu() - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-4, (err, ref.abs().max().item())


def test_wgrad_rejects_bad_arguments(cuda_device):
    lib, st = E.lib(), E.current_stream()
    t = torch.zeros(1, 4, 4, 64, dtype=torch.bfloat16, device=cuda_device)
    dW = torch.zeros(9, 64, 64, device=cuda_device)
    assert lib.fiWgrad(t.data_ptr(), t.data_ptr(), 48, None, 0, 1, 4, 4, 64, dW.data_ptr(), st) != 0
    assert lib.fiWgrad(t.data_ptr(), t.data_ptr(), 64, None, 64, 1, 4, 4, 64, dW.data_ptr(), st) != 0
    assert lib.fiWgrad(None, t.data_ptr(), 64, None, 0, 1, 4, 4, 64, dW.data_ptr(), st) != 0
