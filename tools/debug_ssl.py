import os, sys, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as BN
PKG = BN.PKG
M = importlib.import_module(PKG + ".model"); E = importlib.import_module(PKG + ".engine"); OPT = importlib.import_module(PKG + ".optim")
dev = torch.device("cuda")
cfg = dict(BN.CFG)
rng = np.random.RandomState(1234)
x, lens, ys = BN.synth_batch(rng, 32, 1000, 249, 34)
ux, ulens, _ = BN.synth_batch(np.random.RandomState(2234), 32, 1000, 249, 34)
ld = BN.labeldist_of(ys, 34)
torch.manual_seed(1234)
drop = float(os.environ.get("DROP", 0.3))
m = M.E2E(input_dim=249, enc_hidden_dim=320, enc_n_layers=3, subsample=[2, 2, 2], dropout_rate=drop, dec_hidden_dim=320, att_dim=320,
          conv_channels=10, conv_kernel_size=100, att_odim=320, embedding_dim=128, output_dim=34, ls_weight=0.05, labeldist=ld).to(dev)
m.train()
nj = int(os.environ.get("JUDGE", 0))
if nj:
    lm = M.LM(output_dim=34, embedding_dim=256, hidden_dim=640, dropout_rate=0.5 if drop > 0 else 0.0, n_layers=2, bos=1, eos=2, pad=0,
              ls_weight=0.05, labeldist=ld).to(dev)
    jt = E.JudgeTrainer(lm, OPT.FusedAdam(lm.parameters(), lr=2e-4), max_grad_norm=5.0)
    text = sorted([torch.from_numpy(y).to(dev) for y in ys], key=len, reverse=True)
    for i in range(nj):
        jl, ja, jn = jt.step(text)
    torch.cuda.synchronize()
    print("judge ok, loss", float(jl), "norm", float(jn), flush=True)
pre = int(os.environ.get("PRE", 0))
if pre:
    opt = OPT.FusedAdam(m.parameters(), lr=1e-4, weight_decay=1e-6, amsgrad=True)
    tr = E.SupervisedTrainer(m, opt, max_grad_norm=5.0, use_graph=os.environ.get("GRAPH", "1") == "1")
    pinned = (torch.from_numpy(x).pin_memory(), lens, [torch.from_numpy(y) for y in ys])
    for i in range(pre):
        l, n = tr.step(*pinned)
    torch.cuda.synchronize()
    print("pretrain ok, loss", float(l), "norm", float(n), flush=True)
    print("params finite:", all(bool(torch.isfinite(p).all()) for p in m.parameters()), flush=True)
uxd = torch.from_numpy(ux).to(dev)
with torch.no_grad():
    enc_h, enc_lens = m.encoder(uxd, ulens)
    torch.cuda.synchronize()
    print("encoder ok", bool(torch.isfinite(enc_h).all()), float(enc_h.abs().max()), flush=True)
    logits, logp, pred, ws = m(uxd, ulens, ys=None, sample=False, label_smoothing=False, max_dec_timesteps=125, smooth=True, scaling=3.0)
    torch.cuda.synchronize()
    print("free-run ok; logits finite", bool(torch.isfinite(logits).all()), "logp finite", bool(torch.isfinite(logp).all()),
          "pred range", int(pred.min()), int(pred.max()), "non-eos", int((pred != 2).sum()), "of", pred.numel(), flush=True)
    bad = (~torch.isfinite(logits)).nonzero()
    if len(bad):
        print("first non-finite logits at", bad[:5].tolist())
if os.environ.get("SSL"):
    lm = M.LM(output_dim=34, embedding_dim=256, hidden_dim=640, dropout_rate=0.5 if drop > 0 else 0.0, n_layers=2, bos=1, eos=2, pad=0,
              ls_weight=0.05, labeldist=ld).to(dev)
    gen_opt = OPT.FusedAdam(m.parameters(), lr=1e-4, weight_decay=1e-6, amsgrad=True)
    ssl = E.SSLTrainer(m, lm, gen_opt, max_grad_norm=5.0, unsup_weight=0.001, proportion=0.125, smooth=True, scaling=3.0)
    lab = (torch.from_numpy(x).to(dev), lens, [torch.from_numpy(y).to(dev) for y in ys])
    unlab = (uxd, ulens)
    for i in range(int(os.environ["SSL"])):
        m.train(); lm.train()
        loss, sup, unsup, (u_logp, u_pred, lm_probs) = ssl.losses(lab, unlab)
        gen_opt.zero_grad()
        loss.backward()
        torch.cuda.synchronize()
        gn = {k: float(p.grad.norm()) for k, p in m.named_parameters()}
        badk = [k for k, v in gn.items() if not np.isfinite(v)]
        tot = float(torch.cat([p.grad.flatten() for p in m.parameters()]).double().norm())
        print(f"step {i}: loss {float(loss):.5f} sup {float(sup):.5f} unsup {float(unsup):.5f} non-eos {int((u_pred != 2).sum())} "
              f"gradnorm {tot:.4f} nonfinite grads: {badk[:6]}", flush=True)
        if i == 0:
            print("  top grad norms:", sorted(gn.items(), key=lambda kv: -kv[1] if np.isfinite(kv[1]) else -1e30)[:5])
        gen_opt.clip_and_step(5.0)
