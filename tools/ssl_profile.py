"""Kernel-time breakdown of one captured semi-supervised generator step (config 4) with the torch profiler."""
import collections, importlib, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as BN
from torch.profiler import profile, ProfilerActivity
PKG = BN.PKG
M = importlib.import_module(PKG + ".model"); E = importlib.import_module(PKG + ".engine"); OPT = importlib.import_module(PKG + ".optim")
dev = torch.device("cuda")
cfg = dict(BN.CFG)
x, lens, ys = BN.synth_batch(np.random.RandomState(1234), 32, 1000, 249, 34)
ux, ulens, _ = BN.synth_batch(np.random.RandomState(2234), 32, 1000, 249, 34)
ld = BN.labeldist_of(ys, 34)
torch.manual_seed(1234)
m = M.E2E(input_dim=249, enc_hidden_dim=320, enc_n_layers=3, subsample=[2, 2, 2], dropout_rate=0.3, dec_hidden_dim=320, att_dim=320,
          conv_channels=10, conv_kernel_size=100, att_odim=320, embedding_dim=128, output_dim=34, ls_weight=0.05, labeldist=ld).to(dev)
lm = M.LM(output_dim=34, embedding_dim=256, hidden_dim=640, dropout_rate=0.5, n_layers=2, bos=1, eos=2, pad=0, ls_weight=0.05, labeldist=ld).to(dev)
opt = OPT.FusedAdam(m.parameters(), lr=1e-4, weight_decay=1e-6, amsgrad=True)
tr = E.SSLTrainer(m, lm, opt, proportion=0.125, use_graph=True, guard_empty_mask=True)
lab = (torch.from_numpy(x).to(dev), lens, [torch.from_numpy(y).to(dev) for y in ys])
unlab = (torch.from_numpy(ux).to(dev), ulens)
for _ in range(3):
    tr.step(lab, unlab)
torch.cuda.synchronize()
key = tr.stage(lab, unlab)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.run(key)
    torch.cuda.synchronize()
f = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(f)
ev = [e for e in json.load(open(f))["traceEvents"] if e.get("cat") == "kernel"]
t0 = min(e["ts"] for e in ev); t1 = max(e["ts"] + e["dur"] for e in ev)
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    n = e["name"].replace("(anonymous namespace)::", "").split("(")[0][-70:]
    agg[n][0] += 1; agg[n][1] += e["dur"]
print(f"span {(t1 - t0) / 1e3:.2f} ms, {len(ev)} kernels, sum of kernel times {sum(v for _, v in agg.values()) / 1e3:.2f} ms")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{n:72s} {c:6d} {v / 1e3:9.3f} ms  {v / c:8.1f} us")

streams = collections.defaultdict(list)
for e in ev:
    streams[e["args"]["stream"]].append(e)
for sid, L in sorted(streams.items(), key=lambda kv: -len(kv[1])):
    big = sorted(L, key=lambda e: -e["dur"])[:6]
    print(f"stream {sid}: {len(L)} kernels, busy {sum(e['dur'] for e in L) / 1e3:.2f} ms, active {(min(e['ts'] for e in L) - t0) / 1e3:.2f} .. "
          f"{(max(e['ts'] + e['dur'] for e in L) - t0) / 1e3:.2f} ms")
    for e in sorted(big, key=lambda e: e["ts"]):
        print(f"      {(e['ts'] - t0) / 1e3:8.3f} ms  {e['dur']:8.1f} us  {e['name'].replace('(anonymous namespace)::', '')[:80]}")
