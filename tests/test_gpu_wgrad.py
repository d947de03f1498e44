"""fiTransposePad + fiWgrad (tensor-core weight gradient) against an fp64 reference on the same bf16 operands."""
import pytest
import torch
import torch.nn.functional as F

from model import _engine as E

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 8, 8, 64, 64), (2, 16, 24, 128, 64), (2, 12, 20, 64, 128),
                                            (1, 16, 16, 256, 256), (3, 9, 13, 512, 128), (2, 32, 32, 64, 64)])
def test_wgrad_matches_reference(cuda_device, n, h, w, cin, cout):
    g = torch.Generator().manual_seed(cin + cout + h)
    x = torch.randn(n, cin, h, w, generator=g).to(torch.bfloat16)
    dz = (torch.randn(n, cout, h, w, generator=g) * 0.1).to(torch.bfloat16)
    lib, st = E.lib(), E.current_stream()
    xs = x.permute(0, 2, 3, 1).contiguous().to(cuda_device)
    ds = dz.permute(0, 2, 3, 1).contiguous().to(cuda_device)
    kp = lib.fiTransposePadK(n, h, w)
    wp8 = lib.fiTransposePadRow(w)
    xT = torch.full((3, cin, kp), float("nan"), dtype=torch.bfloat16, device=cuda_device)
    dT = torch.full((cout, kp), float("nan"), dtype=torch.bfloat16, device=cuda_device)
    E.check(lib.fiTransposePad(xs.data_ptr(), xT.data_ptr(), n, h, w, cin, 3, st))
    E.check(lib.fiTransposePad(ds.data_ptr(), dT.data_ptr(), n, h, w, cout, 1, st))
    torch.cuda.synchronize()
    # layout check: interior = transposed pixels, border / tail = zeros; copies 0 and 2 are the row shifted by -1 / +1
    ref_T = F.pad(x.float(), [1, wp8 - w - 1, 1, 1]).permute(1, 0, 2, 3).reshape(cin, -1)
    got = xT.float().cpu()
    assert torch.equal(got[1, :, :ref_T.shape[1]], ref_T) and (got[1, :, ref_T.shape[1]:] == 0).all()
    assert torch.equal(got[0, :, 1:ref_T.shape[1]], ref_T[:, :-1]) and torch.equal(got[2, :, :ref_T.shape[1] - 1], ref_T[:, 1:])
    dW = torch.zeros((9, cout, cin), dtype=torch.float32, device=cuda_device)
    E.check(lib.fiWgrad(dT.data_ptr(), xT.data_ptr(), cout, cin, kp, wp8, dW.data_ptr(), st))
    torch.cuda.synchronize()
    # reference: grad of conv2d w.r.t. its weight
    wt = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), wt, padding=1).backward(dz.double())
    ref = wt.grad.permute(2, 3, 0, 1).reshape(9, cout, cin).float()
    err = (dW.cpu() - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-4, (err, ref.abs().max().item())
