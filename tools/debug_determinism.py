"""Two runs of the same training step on identical copies: per-tensor gradient difference (diagnostic)."""
import copy, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "ai-based-frame-interpolation_b200"))
import torch
from test_gpu_train_step import make_model
from model.train import TrainStep

dev = torch.device("cuda:0")
for overlap in (False, True):
    base = make_model(7).to(dev).train()
    g = torch.Generator().manual_seed(2)
    f1, f2 = torch.rand(2, 1, 32, 32, generator=g).to(dev), torch.rand(2, 1, 32, 32, generator=g).to(dev)
    tgt = (f1 + f2) / 2
    grads = []
    for run in range(3):
        m = copy.deepcopy(base)
        s = TrainStep(m, lr=0.0, overlap_wgrad=overlap)
        s.keep_activations = True
        s(f1, f2, tgt)
        torch.cuda.synchronize()
        grads.append({n: s.grad_view[p].clone() for n, p in m.named_parameters()})
        if run == 0:
            acts0 = {k: v.float().clone() for k, v in s.last_activations.items()}
        elif run == 1:
            print(f"overlap={overlap}: activation differences between two runs")
            for k, v in s.last_activations.items():
                d = (v.float() - acts0[k])
                print(f"   {k:10s} rel {(d.norm() / (acts0[k].norm() + 1e-20)).item():.3e}  changed {int((d != 0).sum())} of {d.numel()}")
    worst = []
    for n in grads[0]:
        a, b, c = grads[0][n], grads[1][n], grads[2][n]
        rel = max(((a - b).norm() / (a.norm() + 1e-20)).item(), ((a - c).norm() / (a.norm() + 1e-20)).item())
        worst.append((rel, n))
    worst.sort(reverse=True)
    print(f"overlap={overlap}: worst run-to-run relative gradient differences")
    for rel, n in worst[:6]:
        print(f"   {rel:.3e}  {n}")
