"""GPU parity of the LSTM recurrence kernels (cluster-persistent and per-timestep variants) against
the CPU oracle: one BLSTM layer forward (zeros past the length, replicate row) and its BPTT."""
import numpy as np
import pytest
import torch

from oracle import las_oracle as O
from tests.util import cosine, pkg, rel_err

pytestmark = pytest.mark.gpu


def _layer_case(seed, B, T, D, H):
    torch.manual_seed(seed)
    rng = np.random.RandomState(seed)
    lens = sorted([T] + [int(rng.randint(1, T + 1)) for _ in range(B - 1)], reverse=True)
    x = torch.zeros(B, T, D)
    for b, l in enumerate(lens):
        x[b, :l] = torch.randn(l, D)
    lstm = torch.nn.LSTM(D, H, num_layers=1, bidirectional=True, batch_first=True)
    P = {"L." + k: v.detach().clone() for k, v in lstm.state_dict().items()}
    return x, lens, P


def _run_gpu(x, lens, P, H, persistent):
    Fn = pkg("functional")
    L = pkg("_lib").lib()
    prev = L.las_set_persistent(1 if persistent else 0)
    try:
        B, T, D = x.shape
        dev = torch.device("cuda")
        names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
        w = [P["L." + n].to(dev).requires_grad_(True) for n in names] + \
            [P["L." + n + "_reverse"].to(dev).requires_grad_(True) for n in names]
        # identity-like projection so that EncoderFn exposes the BLSTM output: proj = [I; 0] picks y[:, :H'] ...
        # simpler: use a random projection and compare through the oracle's encoder_forward
        torch.manual_seed(99)
        # (a 0.2-scaled random projection saturates at the wide sizes: its 4H-long rows get 1/sqrt(4H)-like entries there)
        proj_w = (torch.randn(H, 4 * H) * (0.2 if H <= 320 else 0.03)).to(dev).requires_grad_(True)
        proj_b = (torch.randn(H) * 0.1).to(dev).requires_grad_(True)
        lens_dev = Fn.lens_tensor(lens, dev)
        out = Fn.EncoderFn.apply(x.to(dev), lens_dev, (2,), 0.0, *w, proj_w, proj_b)
        torch.manual_seed(7)
        gout = torch.randn(out.shape).to(dev)
        (out * gout).sum().backward()
        return out.detach().cpu(), [t.grad.detach().cpu() for t in w + [proj_w, proj_b]], gout.cpu(), \
            proj_w.detach().cpu(), proj_b.detach().cpu()
    finally:
        L.las_set_persistent(prev)


CASES = [dict(seed=1, B=3, T=9, D=16, H=8), dict(seed=2, B=11, T=22, D=24, H=64), dict(seed=3, B=8, T=17, D=40, H=48),
         dict(seed=4, B=5, T=31, D=32, H=320), dict(seed=5, B=17, T=12, D=16, H=24),
         # "wide" geometry of the persistent kernels (16-CTA clusters, part of the recurrent weights resident in shared
         # memory): the LM judge's hidden size 640 (model.py:466) and a size in between
         dict(seed=6, B=9, T=24, D=32, H=640), dict(seed=7, B=4, T=11, D=24, H=400),
         # three clusters per direction with an odd T: the later (shorter) groups stop at their own longest utterance
         # and fill the rest of y / hprev / dG (and the replicated row) in the tail loop
         dict(seed=8, B=19, T=15, D=16, H=32)]


@pytest.mark.parametrize("persistent", [False, True])
@pytest.mark.parametrize("cfg", CASES)
def test_blstm_layer_against_oracle(cfg, persistent):
    x, lens, P = _layer_case(**cfg)
    H = cfg["H"]
    out, grads, gout, proj_w, proj_b = _run_gpu(x, lens, P, H, persistent)
    PP = {k.replace("L.", "encoder.enc2.layers.0."): v.clone().requires_grad_(True) for k, v in P.items()}
    PP["encoder.enc2.project_layers.0.weight"] = proj_w.clone().requires_grad_(True)
    PP["encoder.enc2.project_layers.0.bias"] = proj_b.clone().requires_grad_(True)
    ref, ref_lens = O.encoder_forward(x, lens, PP, [2])
    assert ref_lens == [(l + 1) // 2 for l in lens]
    assert rel_err(out, ref) < 3e-2
    (ref * gout).sum().backward()
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    keys = ["encoder.enc2.layers.0." + n for n in names] + ["encoder.enc2.layers.0." + n + "_reverse" for n in names] + \
        ["encoder.enc2.project_layers.0.weight", "encoder.enc2.project_layers.0.bias"]
    for k, gv in zip(keys, grads):
        assert cosine(gv, PP[k].grad) > 0.999, (k, cosine(gv, PP[k].grad))


@pytest.mark.parametrize("cfg", CASES)
def test_persistent_equals_stepwise(cfg):
    """Both kernel families implement the same arithmetic (bf16 state, f32 accumulate): outputs agree
    to accumulation-order noise."""
    x, lens, P = _layer_case(**cfg)
    a = _run_gpu(x, lens, P, cfg["H"], False)
    b = _run_gpu(x, lens, P, cfg["H"], True)
    assert rel_err(a[0], b[0]) < 1e-2
    for ga, gb in zip(a[1], b[1]):
        assert cosine(ga, gb) > 0.9995
