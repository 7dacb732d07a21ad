"""Worst per-tensor gradient cosines of the three random cases of tests/test_gpu_supervised.py::test_random_case_against_oracle
(the numbers behind the per-tensor bar in DESIGN.md section 2). LAS_FAST_ACT selects the BLSTM gate-activation form."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import las_oracle as O
from tests.test_gpu_supervised import _random_case, cosine

CASES = [
    dict(seed=3, B=5, T=61, D=40, H=64, sub=[2, 2, 2], V=20, E=32, A=48, C=5, K=7, ls=0.05),
    dict(seed=4, B=9, T=48, D=249, H=32, sub=[1, 2, 2], V=34, E=16, A=32, C=10, K=20, ls=0.0),
    dict(seed=5, B=4, T=320, D=249, H=320, sub=[2, 2, 2], V=34, E=128, A=320, C=10, K=100, ls=0.05),
]
for cfg in CASES:
    m, P, x, lens, ys, labeldist = _random_case(**cfg)
    loss_o, grads_o, _, _ = O.supervised_step(torch.from_numpy(x), lens, ys, P, {}, cfg["sub"], cfg["ls"], labeldist, fast=True)
    m.train()
    _, logp, _, _ = m(torch.as_tensor(x).cuda(), lens, [torch.from_numpy(y).cuda() for y in ys])
    loss = -torch.mean(logp)
    m.zero_grad()
    loss.backward()
    total = float(torch.cat([g.flatten() for g in grads_o.values()]).double().norm())
    worst = sorted((cosine(p.grad, grads_o[k]), k, float(grads_o[k].double().norm()) / total) for k, p in m.named_parameters())
    print(f"seed {cfg['seed']} LAS_FAST_ACT={os.environ.get('LAS_FAST_ACT')} loss rel {abs(float(loss) - loss_o) / abs(loss_o):.2e}")
    for c, k, share in worst[:4]:
        print(f"   {c:.6f}  {k}  (norm share {share:.2e})")
