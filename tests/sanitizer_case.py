"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck), one tool per gpurun call:
    compute-sanitizer --tool memcheck python tests/sanitizer_case.py
Covers both conv kernels (all epilogue modes), the stem, the ConvT scatter store, the bilinear decoder, the fp32x3 path,
the clip pipeline and the metric kernel on odd sizes (partial tiles + F.pad path)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "ai-based-frame-interpolation_b200"))
sys.path.insert(0, str(ROOT))
from model import _engine as E  # noqa: E402
from oracle import unet_oracle as O  # noqa: E402

dev = torch.device("cuda", 0)
rs = np.random.RandomState(0)
f = rs.randint(0, 256, size=(4, 1, 38, 54)).astype(np.uint8)
for bilinear in (False, True):
    for precision in ("bf16", "fp32"):
        sd = O.init_state_dict(0, 2, 1, bilinear)
        net = E.Net(dev, 2, 1, bilinear, precision)
        net.load_state_dict(sd)
        out = net.interpolate_clip_host_u8(f, pairs_per_batch=2)
        ref = O.postprocess(O.unet_forward(sd, torch.cat([O.preprocess_u8(f[:-1]), O.preprocess_u8(f[1:])], 1)))
        d = np.abs(out.astype(int) - ref.astype(int)).max()
        m = E.ssim_psnr_u8(torch.from_numpy(out[:, 0]).to(dev), torch.from_numpy(f[:-1, 0]).to(dev)).cpu().numpy()
        print(f"bilinear={bilinear} precision={precision}: max u8 diff {d}, psnr {m[0, 0]:.2f}")
        assert d <= 6
        net.close()
torch.cuda.synchronize()
print("sanitizer case ok")
