"""Kernel timeline of the training step (torch.profiler / CUPTI): per-kernel time, gaps and overlap (diagnostic)."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ai-based-frame-interpolation_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from model.train import TrainStep
from model.unet import FrameInterpolationUNet

graph = "--graph" in sys.argv
overlap = "--no-overlap" not in sys.argv
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = FrameInterpolationUNet(bilinear=True).to(dev).train()
step = TrainStep(model, lr=1e-4, cuda_graph=graph, overlap_wgrad=overlap)
f0, f1 = (torch.rand(16, 1, 256, 256, device=dev) for _ in range(2))
gt = (f0 + f1) / 2
for _ in range(6):
    step(f0, f1, gt)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step(f0, f1, gt)
    torch.cuda.synchronize()
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", f"train_trace_{'graph' if graph else 'eager'}{'' if overlap else '_noov'}.json")
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
# last step: split at adam kernels
adam = [i for i, e in enumerate(ev) if "adam" in e["name"]]
last = ev[adam[-2] + 1: adam[-1] + 1]
t0, t1 = last[0]["ts"], last[-1]["ts"] + last[-1]["dur"]
busy = sum(e["dur"] for e in last)
# union of intervals
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in last)
union, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
for s, e in iv[1:]:
    if s > cur_e:
        union += cur_e - cur_s; cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
union += cur_e - cur_s
print(f"{'graph' if graph else 'eager'} overlap={overlap}: span {t1 - t0:.0f} us, sum of kernels {busy:.0f} us, union {union:.0f} us, idle {t1 - t0 - union:.0f} us, {len(last)} kernels")
import collections, re
agg = collections.defaultdict(lambda: [0, 0.0])
for e in last:
    n = re.sub(r"\(anonymous namespace\)::", "", e["name"])
    n = re.sub(r"\(.*", "", n).replace("void ", "")[:60]
    agg[n][0] += 1; agg[n][1] += e["dur"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {v[1]:8.1f} us {v[0]:4d} {k}")
