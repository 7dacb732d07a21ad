"""Per-parameter gradient cosine / loss error of the CUDA path vs the CPU oracle on random cases."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import las_oracle as O
from tests.util import cosine, pkg, rel_err
from tests.test_gpu_supervised import _random_case

cfgs = [dict(seed=3, B=5, T=61, D=40, H=64, sub=[2, 2, 2], V=20, E=32, A=48, C=5, K=7, ls=0.05),
        dict(seed=4, B=9, T=48, D=249, H=32, sub=[1, 2, 2], V=34, E=16, A=32, C=10, K=20, ls=0.0),
        dict(seed=5, B=8, T=200, D=249, H=320, sub=[2, 2, 2], V=34, E=128, A=320, C=10, K=100, ls=0.05)]
which = [int(a) for a in sys.argv[1:]] or [0, 1]
for ci in which:
    cfg = cfgs[ci]
    m, P, x, lens, ys, labeldist = _random_case(**cfg)
    t0 = time.time()
    loss_o, grads_o, norm_o, _ = O.supervised_step(torch.from_numpy(x), lens, ys, P, {}, cfg["sub"], cfg["ls"], labeldist, fast=True)
    t_cpu = time.time() - t0
    m.train()
    _, logp, _, _ = m(torch.from_numpy(x).cuda(), lens, [torch.from_numpy(y).cuda() for y in ys])
    loss = -torch.mean(logp)
    m.zero_grad(); loss.backward(); torch.cuda.synchronize()
    print(f"cfg{ci}: loss {float(loss):.6f} oracle {loss_o:.6f} rel {abs(float(loss)-loss_o)/abs(loss_o):.2e} cpu_step {t_cpu:.2f}s")
    a, b = [], []
    for k, p in m.named_parameters():
        c = cosine(p.grad, grads_o[k])
        nr = float(p.grad.norm()) / (float(grads_o[k].norm()) + 1e-30)
        print(f"   {k:50s} cos {c:.6f} norm_ratio {nr:.4f} |g| {float(grads_o[k].norm()):.3e}")
        a.append(p.grad.detach().cpu().flatten()); b.append(grads_o[k].flatten())
    print("   WHOLE cos", cosine(torch.cat(a), torch.cat(b)))
