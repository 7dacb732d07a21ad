import importlib, os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench as BN
PKG = BN.PKG
M = importlib.import_module(PKG + ".model")
dev = torch.device("cuda")
ux, ulens, ys = BN.synth_batch(np.random.RandomState(2234), 32, 1000, 249, 34)
ld = BN.labeldist_of(ys, 34)
torch.manual_seed(1234)
m = M.E2E(**BN._cfg_kwargs(ld, 0.3)).to(dev).train()
x = torch.from_numpy(ux).to(dev)
for it in range(2):
    if it == 1:
        torch.cuda.profiler.start()
    _, logp, pred, _ = m(x, ulens, ys=None, label_smoothing=False, max_dec_timesteps=12, smooth=True, scaling=3.0)
    (-logp.mean()).backward()
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
