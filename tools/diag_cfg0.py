"""Per-parameter gradient cosines of the failing random case (cfg0) against the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import las_oracle as O
from tests.util import cosine, pkg
from tests.test_gpu_supervised import _random_case
cfg = dict(seed=3, B=5, T=61, D=40, H=64, sub=[2, 2, 2], V=20, E=32, A=48, C=5, K=7, ls=0.05)
if os.environ.get("LAS_BIG"):
    cfg = dict(seed=5, B=4, T=320, D=249, H=320, sub=[2, 2, 2], V=34, E=128, A=320, C=10, K=100, ls=0.05)
m, P, x, lens, ys, labeldist = _random_case(**cfg)
loss_o, grads_o, norm_o, _ = O.supervised_step(torch.from_numpy(x), lens, ys, P, {}, cfg["sub"], cfg["ls"], labeldist, fast=True)
m.train()
_, logp, _, _ = m(torch.from_numpy(x).cuda(), lens, [torch.from_numpy(y).cuda() for y in ys])
loss = -torch.mean(logp)
m.zero_grad(); loss.backward()
print("loss", float(loss), loss_o)
for k, p in m.named_parameters():
    c = cosine(p.grad, grads_o[k])
    print(f"{k:50s} cos {c:.5f}  |g| {float(p.grad.norm()):.3e} oracle {float(grads_o[k].norm()):.3e}")
