"""Loss / gradient parity of the CUDA path against the CPU oracle at BASELINE config-2 size (B=32, Tmax=1000): relative
loss error, whole-model gradient cosine and the per-tensor cosines (the numbers behind tests/test_gpu_fullsize.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import las_oracle as O
from tests.test_gpu_supervised import _random_case, cosine
from tests.test_gpu_fullsize import CFG2

m, P, x, lens, ys, labeldist = _random_case(**CFG2)
loss_o, grads_o, _, _ = O.supervised_step(torch.from_numpy(x), lens, ys, P, {}, CFG2["sub"], CFG2["ls"], labeldist, fast=True)
m.train()
_, logp, _, _ = m(torch.as_tensor(x).cuda(), lens, [torch.from_numpy(y).cuda() for y in ys])
loss = -torch.mean(logp)
m.zero_grad()
loss.backward()
total = float(torch.cat([g.flatten() for g in grads_o.values()]).double().norm())
a, b, worst = [], [], []
for k, p in m.named_parameters():
    a.append(p.grad.detach().cpu().flatten()); b.append(grads_o[k].flatten())
    worst.append((cosine(p.grad, grads_o[k]), k, float(grads_o[k].double().norm()) / total))
whole = cosine(torch.cat(a), torch.cat(b))
print(f"LAS_FAST_ACT={os.environ.get('LAS_FAST_ACT')} loss {float(loss):.6f} oracle {loss_o:.6f} rel {abs(float(loss) - loss_o) / abs(loss_o):.2e}; whole-model gradient cosine {whole:.6f}")
for c, k, share in sorted(worst)[:8]:
    print(f"   {c:.6f}  {k}  (norm share {share:.2e})")
