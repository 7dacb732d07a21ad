"""Kernel timeline of one captured train step (torch profiler over a graph replay): per-stream busy time, the
main-stream kernels in launch order with gaps, for finding what sits on the critical path."""
import json, os, sys, importlib, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as BN
from torch.profiler import profile, ProfilerActivity
PKG = BN.PKG
M = importlib.import_module(PKG + ".model"); E = importlib.import_module(PKG + ".engine"); OPT = importlib.import_module(PKG + ".optim")
dev = torch.device("cuda")
rng = np.random.RandomState(1234)
B, T = int(os.environ.get("B", 32)), int(os.environ.get("T", 1000))
x, lens, ys = BN.synth_batch(rng, B, T, 249, 34)
ld = BN.labeldist_of(ys, 34)
C = dict(BN.CFG)
if os.environ.get("SUB"):
    C["subsample"] = [int(v) for v in os.environ["SUB"].split(",")]
    C["enc_n_layers"] = len(C["subsample"])
torch.manual_seed(1234)
m = M.E2E(input_dim=C["input_dim"], enc_hidden_dim=C["enc_hidden_dim"], enc_n_layers=C["enc_n_layers"], subsample=C["subsample"],
          dropout_rate=0.3, dec_hidden_dim=C["dec_hidden_dim"], att_dim=C["att_dim"], conv_channels=C["conv_channels"],
          conv_kernel_size=C["conv_kernel_size"], att_odim=C["att_odim"], embedding_dim=C["embedding_dim"], output_dim=C["V"],
          ls_weight=C["ls_weight"], labeldist=ld).to(dev)
opt = OPT.FusedAdam(m.parameters(), lr=C["lr"], weight_decay=C["weight_decay"], amsgrad=True)
tr = E.SupervisedTrainer(m, opt, max_grad_norm=5.0)
batch = (torch.from_numpy(x).pin_memory(), lens, [torch.from_numpy(y) for y in ys])
for _ in range(4):
    tr.step(*batch)
torch.cuda.synchronize()
key = tr.stage(*batch)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.run(key)
    torch.cuda.synchronize()
f = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(f)
ev = [e for e in json.load(open(f))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in ev)
print(f"step span {(t1 - t0) / 1e3:.3f} ms, {len(ev)} kernels")
streams = {}
for e in ev:
    streams.setdefault(e["args"]["stream"], []).append(e)
for s, L in streams.items():
    print(f"stream {s}: {len(L)} kernels, busy {sum(e['dur'] for e in L) / 1e3:.3f} ms")
main = max(streams.values(), key=len)
big = float(os.environ.get("MIN_US", 15))
prev_end = t0
agg_small, n_small, gap_tot = 0.0, 0, 0.0
for e in main:
    gap = e["ts"] - prev_end
    gap_tot += max(gap, 0)
    if e["dur"] >= big or gap > 10:
        if n_small:
            print(f"      ... {n_small} small kernels, {agg_small:.0f} us")
            agg_small, n_small = 0.0, 0
        print(f"{(e['ts'] - t0) / 1e3:8.3f} ms  +gap {gap:6.1f} us  {e['dur']:8.1f} us  {e['name'][:90]}")
    else:
        agg_small += e["dur"]; n_small += 1
    prev_end = max(prev_end, e["ts"] + e["dur"])
print(f"main-stream gaps total {gap_tot / 1e3:.3f} ms")
print("---- other streams")
for s, L in streams.items():
    if L is main:
        continue
    for e in L:
        print(f"{(e['ts'] - t0) / 1e3:8.3f} ms  {e['dur']:8.1f} us  [s{s}] {e['name'][:90]}")
head_us = float(os.environ.get("HEAD_US", 0))
if head_us > 0:
    print(f"---- every kernel of the first {head_us:.0f} us, all streams")
    for e in ev:
        if e["ts"] - t0 > head_us:
            break
        print(f"{(e['ts'] - t0):8.1f} us  {e['dur']:7.1f} us  [s{e['args']['stream']}] {e['name'][:100]}")
