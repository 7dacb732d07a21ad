"""Per-tensor cosines of one tests/test_gpu_lstm.py case: persistent / per-timestep kernels against the oracle.
LAS_LSTM_WALK_ALL=1 makes every cluster of the persistent kernels walk all T timesteps (A/B for the per-group walk)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import las_oracle as O  # noqa: E402
from tests import test_gpu_lstm as TL  # noqa: E402
from tests.util import cosine, rel_err  # noqa: E402

idx = int(sys.argv[1]) if len(sys.argv) > 1 else len(TL.CASES) - 1
cfg = dict(TL.CASES[idx])
for kv in sys.argv[2:]:          # overrides: B=19 T=15 D=16 H=32 seed=8
    k, v = kv.split("=")
    cfg[k] = int(v)
x, lens, P = TL._layer_case(**cfg)
print(cfg, "lens", lens, "walk_all", os.environ.get("LAS_LSTM_WALK_ALL"))
names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
keys = ["encoder.enc2.layers.0." + n for n in names] + ["encoder.enc2.layers.0." + n + "_reverse" for n in names] + \
    ["encoder.enc2.project_layers.0.weight", "encoder.enc2.project_layers.0.bias"]
res = {}
for persistent in (False, True):
    out, grads, gout, proj_w, proj_b = TL._run_gpu(x, lens, P, cfg["H"], persistent)
    PP = {k.replace("L.", "encoder.enc2.layers.0."): v.clone().requires_grad_(True) for k, v in P.items()}
    PP["encoder.enc2.project_layers.0.weight"] = proj_w.clone().requires_grad_(True)
    PP["encoder.enc2.project_layers.0.bias"] = proj_b.clone().requires_grad_(True)
    ref, _ = O.encoder_forward(x, lens, PP, [2])
    (ref * gout).sum().backward()
    print("persistent" if persistent else "per-step  ", "out rel_err %.2e" % rel_err(out, ref),
          " ".join("%s %.5f" % (k.split(".")[-1], cosine(g, PP[k].grad)) for k, g in zip(keys, grads)))
    res[persistent] = (out, grads)
print("persistent vs per-step: out %.2e" % rel_err(res[True][0], res[False][0]),
      " ".join("%.5f" % cosine(a, b) for a, b in zip(res[True][1], res[False][1])))
