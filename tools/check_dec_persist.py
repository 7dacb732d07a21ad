"""Forward of the cluster-persistent decoder against the per-step kernels on the same GPU inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import pkg
Fn = pkg("functional"); M = pkg("model")

def run(B, Te, Hd, A, C, K, E, V, L, seed=0):
    torch.manual_seed(seed)
    dev = torch.device("cuda")
    att = M.AttLoc(Hd, Hd, A, C, K, Hd)
    dec = M.Decoder(V, E, Hd, att, Hd, 0.0, 1, 2, 0).to(dev)
    enc_h = torch.relu(torch.randn(B, Te, Hd, device=dev))
    lens = torch.tensor(sorted([Te] + [int(torch.randint(max(1, Te // 2), Te + 1, (1,))) for _ in range(B - 1)], reverse=True), dtype=torch.int32, device=dev)
    ys_in = torch.randint(3, V, (B, L + 1), device=dev)
    ys_out = torch.randint(3, V, (B, L), device=dev)
    outs = []
    for flag in (False, True):
        Fn.DEC_PERSISTENT = flag
        with torch.no_grad():
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for it in range(2):
                e0.record()
                logits, logp, pred, ws = dec.forward_dev(enc_h, lens, ys_in, ys_out, L, 0)
                e1.record()
                torch.cuda.synchronize()
            outs.append((logits.float().clone(), ws.clone(), e0.elapsed_time(e1)))
    (l0, w0, t0), (l1, w1, t1) = outs
    print(f"B={B} Te={Te} Hd={Hd} A={A} L={L}: stepwise {t0:.3f} ms  persistent {t1:.3f} ms | "
          f"logits max|d| {float((l0 - l1).abs().max()):.3e} (max {float(l0.abs().max()):.2f})  ws max|d| {float((w0 - w1).abs().max()):.3e}  "
          f"ws rowsum err {float((w1.sum(-1) - 1).abs().max()):.2e}")

if os.environ.get("LAS_ONLY_BIG"):
    run(32, 125, 320, 320, 10, 100, 128, 34, 126)
else:
    run(5, 8, 64, 48, 5, 7, 32, 20, 9)
    run(3, 40, 64, 48, 5, 7, 32, 20, 12)
    run(9, 37, 128, 64, 10, 20, 16, 34, 11)
    run(32, 125, 320, 320, 10, 100, 128, 34, 126)
    run(64, 250, 320, 320, 10, 100, 128, 34, 60)

if os.environ.get("LAS_PROF") or os.environ.get("LAS_TRACE"):
    from torch.profiler import profile, ProfilerActivity
    torch.manual_seed(0)
    dev = torch.device("cuda")
    B, Te, Hd, A, C, K, E, V, L = 32, 125, 320, 320, 10, 100, 128, 34, 126
    att = M.AttLoc(Hd, Hd, A, C, K, Hd)
    dec = M.Decoder(V, E, Hd, att, Hd, 0.0, 1, 2, 0).to(dev)
    enc_h = torch.relu(torch.randn(B, Te, Hd, device=dev))
    lens = torch.full((B,), Te, dtype=torch.int32, device=dev)
    ys_in = torch.randint(3, V, (B, L + 1), device=dev)
    ys_out = torch.randint(3, V, (B, L), device=dev)
    Fn.DEC_PERSISTENT = True
    with torch.no_grad():
        dec.forward_dev(enc_h, lens, ys_in, ys_out, L, 0)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            dec.forward_dev(enc_h, lens, ys_in, ys_out, L, 0)
            torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=60))

if os.environ.get("LAS_TRACE"):
    LIB = pkg("_lib")
    dbg = torch.zeros(2048, device="cuda", dtype=torch.int64)
    LIB.lib().las_set_debug_buffer(dbg.data_ptr())
    Fn.DEC_PERSISTENT = True
    with torch.no_grad():
        dec.forward_dev(enc_h, lens, ys_in, ys_out, L, 0)
    torch.cuda.synchronize()
    LIB.lib().las_set_debug_buffer(None)
    d = dbg.cpu()[:1024].view(16, 4, 16)
    names = ["wait z,c", "mma+sync", "epi+send", "conv", "wait z+dz", "dz-send", "wait dz+E", "sync", "e-send", "wait e", "smax+ctx"]
    print("per-warp phase durations (cycles), step 9; slots 12/13 = epilogue math done / sends done (relative to slot 2)")
    print("warp " + " ".join(f"{n:>9s}" for n in names) + " |     step   epi-math  epi-sends")
    for w in range(16):
        row, nxt = d[w, 1], d[w, 2]
        if int(row[0]) == 0:
            continue
        print(f"{w:4d} " + " ".join(f"{int(row[i + 1] - row[i]):9d}" for i in range(11)) + f" | {int(nxt[0] - row[0]):8d}"
              + (f" {int(row[12] - row[2]):10d} {int(row[13] - row[2]):10d}" if int(row[12]) else " " * 22)
              + f" | softmax {int(row[14] - row[10]):6d}" + (f" ctx-mma {int(row[15] - row[14]):6d} tail {int(row[11] - row[15]):6d}" if int(row[15]) else ""))
    t0 = int(d[0, 1, 0])
    print("absolute slot times of step 9 relative to warp 0 slot 0:")
    for w in range(16):
        if int(d[w, 1, 0]):
            print(f"{w:4d} " + " ".join(f"{int(d[w, 1, i]) - t0:7d}" for i in range(12)))
