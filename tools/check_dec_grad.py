"""Gradients of the cluster-persistent decoder against the per-step kernels (same GPU inputs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import pkg, cosine
Fn = pkg("functional"); M = pkg("model")

def run(B, Te, Hd, A, C, K, E, V, L, seed=0):
    torch.manual_seed(seed)
    dev = torch.device("cuda")
    att = M.AttLoc(Hd, Hd, A, C, K, Hd)
    dec = M.Decoder(V, E, Hd, att, Hd, 0.0, 1, 2, 0).to(dev)
    enc0 = torch.relu(torch.randn(B, Te, Hd, device=dev))
    lens = torch.tensor(sorted([Te] + [int(torch.randint(max(1, Te // 2), Te + 1, (1,))) for _ in range(B - 1)], reverse=True), dtype=torch.int32, device=dev)
    ys_in = torch.randint(3, V, (B, L + 1), device=dev)
    ys_out = torch.randint(3, V, (B, L), device=dev)
    res = []
    for flag in (False, True):
        Fn.DEC_PERSISTENT = flag
        enc_h = enc0.clone().requires_grad_(True)
        dec.zero_grad()
        for it in range(2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dec.zero_grad(); enc_h.grad = None
            logits, logp, pred, ws = dec.forward_dev(enc_h, lens, ys_in, ys_out, L, 0)
            loss = -logp.mean()
            e0.record()
            loss.backward()
            e1.record()
            torch.cuda.synchronize()
        g = {k: v.grad.detach().clone() for k, v in dec.named_parameters()}
        g["enc_h"] = enc_h.grad.detach().clone()
        res.append((float(loss), g, e0.elapsed_time(e1)))
    (l0, g0, t0), (l1, g1, t1) = res
    worst = min((cosine(g0[k], g1[k]), k) for k in g0)
    print(f"B={B} Te={Te} Hd={Hd} A={A} L={L}: loss {l0:.6f} / {l1:.6f}; bwd stepwise {t0:.2f} ms persistent {t1:.2f} ms; worst grad cosine {worst[0]:.6f} ({worst[1]})")
    for k in g0:
        c = cosine(g0[k], g1[k])
        if c < 0.999:
            print(f"   {k}: cos {c:.5f} |g0| {float(g0[k].norm()):.3e} |g1| {float(g1[k].norm()):.3e}")

if os.environ.get("LAS_ONLY_BIG"):
    run(32, 125, 320, 320, 10, 100, 128, 34, 126)
else:
    run(5, 8, 64, 48, 5, 7, 32, 20, 9)
    run(3, 40, 64, 48, 5, 7, 32, 20, 12)
    run(9, 37, 128, 64, 10, 20, 16, 34, 11)
    run(32, 125, 320, 320, 10, 100, 128, 34, 126)

if os.environ.get("LAS_TRACE"):
    LIB = pkg("_lib")
    torch.manual_seed(0)
    dev = torch.device("cuda")
    B, Te, Hd, A, C, K, E, V, L = 32, 125, 320, 320, 10, 100, 128, 34, 126
    att = M.AttLoc(Hd, Hd, A, C, K, Hd)
    dec = M.Decoder(V, E, Hd, att, Hd, 0.0, 1, 2, 0).to(dev)
    enc_h = torch.relu(torch.randn(B, Te, Hd, device=dev)).requires_grad_(True)
    lens = torch.full((B,), Te, dtype=torch.int32, device=dev)
    ys_in = torch.randint(3, V, (B, L + 1), device=dev)
    ys_out = torch.randint(3, V, (B, L), device=dev)
    Fn.DEC_PERSISTENT = True
    dbg = torch.zeros(2048, device=dev, dtype=torch.int64)
    LIB.lib().las_set_debug_buffer(dbg.data_ptr())
    logits, logp, pred, ws = dec.forward_dev(enc_h, lens, ys_in, ys_out, L, 0)
    (-logp.mean()).backward()
    torch.cuda.synchronize()
    LIB.lib().las_set_debug_buffer(None)
    d = dbg.cpu()[1024:1024 + 12 * 64].view(12, 4, 16)
    names = ["A:mma+epi", "wait dc", "B1-2:dw,de", "B3:energy", "B4:ddz,dwn", "wait ddz", "C:cell", "stores"]
    print("per-warp phase durations (cycles), second traced step")
    print("warp " + " ".join(f"{n:>11s}" for n in names) + " |     step")
    for w in range(12):
        row, nxt = d[w, 1], d[w, 2]
        if int(row[0]) == 0:
            continue
        print(f"{w:4d} " + " ".join(f"{int(row[i + 1] - row[i]):11d}" for i in range(8)) + f" | {int(nxt[0] - row[0]):8d}"
              + f" | A: wait-dg {int(row[9] - row[0])} mma {int(row[10] - row[9])} sync {int(row[11] - row[10])} epi {int(row[1] - row[11])}"
              + f" | B4: ddz+dconv {int(row[12] - row[4])} G-passes {int(row[13] - row[12])} sync {int(row[14] - row[13])} dwn-out {int(row[5] - row[14])}")
