"""Greedy decoding (mode 1, eval, no_grad) decoder-only timing: cluster-persistent launch vs per-timestep kernels.
Prints one JSON line per (batch, frames) point."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as BN
from tools.bench_configs import make_e2e, timed

Fn = importlib.import_module(BN.PKG + ".functional")
dev = torch.device("cuda")
steps = int(os.environ.get("STEPS", 230))
for sub, T in (([2, 2, 2], 1000), ([1, 2, 2, 2], 2000)):
    cfg = dict(BN.CFG)
    cfg.update(enc_n_layers=len(sub), subsample=sub)
    rng = np.random.RandomState(7)
    x, lens, ys = BN.synth_batch(rng, 32, T, cfg["input_dim"], cfg["V"])
    torch.manual_seed(7)
    m = make_e2e(cfg, BN.labeldist_of(ys, cfg["V"]), 0.0, dev).eval()
    xd = torch.from_numpy(x).to(dev)
    with torch.no_grad():
        for B in (1, 8, 32):
            enc_h, enc_l = m.encoder(xd[:B], lens[:B])
            row = {"B": B, "T": T, "Te": int(enc_h.shape[1]), "steps": steps}
            for name, flag in (("persistent", True), ("per_step", False)):
                Fn.DEC_PERSISTENT = flag
                ms = timed(lambda: m.decoder(enc_h, enc_l, ys=None, max_dec_timesteps=steps), 5, 2)
                row[name + "_ms"] = ms
                row[name + "_us_per_step"] = ms * 1e3 / steps
            Fn.DEC_PERSISTENT = True
            print(json.dumps(row), flush=True)
