"""Where the end-to-end loop (SupervisedTrainer.steps over pinned host batches) spends its time between steps."""
import json, os, sys, importlib, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as BN
from torch.profiler import profile, ProfilerActivity
PKG = BN.PKG
M = importlib.import_module(PKG + ".model"); E = importlib.import_module(PKG + ".engine"); OPT = importlib.import_module(PKG + ".optim")
dev = torch.device("cuda")
rng = np.random.RandomState(1234)
C = BN.CFG
batches = [BN.synth_batch(rng, 32, 1000, 249, 34) for _ in range(2)]
ld = BN.labeldist_of([y for b in batches for y in b[2]], 34)
torch.manual_seed(1234)
m = M.E2E(input_dim=C["input_dim"], enc_hidden_dim=C["enc_hidden_dim"], enc_n_layers=C["enc_n_layers"], subsample=C["subsample"],
          dropout_rate=0.3, dec_hidden_dim=C["dec_hidden_dim"], att_dim=C["att_dim"], conv_channels=C["conv_channels"],
          conv_kernel_size=C["conv_kernel_size"], att_odim=C["att_odim"], embedding_dim=C["embedding_dim"], output_dim=C["V"],
          ls_weight=C["ls_weight"], labeldist=ld).to(dev)
opt = OPT.FusedAdam(m.parameters(), lr=C["lr"], weight_decay=C["weight_decay"], amsgrad=True)
tr = E.SupervisedTrainer(m, opt, max_grad_norm=5.0)
pinned = [(torch.from_numpy(x).pin_memory(), lens, [torch.from_numpy(y) for y in ys]) for x, lens, ys in batches]
for loss, _ in tr.steps(pinned[i % 2] for i in range(6)):
    loss.item()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for loss, _ in tr.steps(pinned[i % 2] for i in range(6)):
        loss.item()
    torch.cuda.synchronize()
f = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(f)
tr_ev = json.load(open(f))["traceEvents"]
ev = [e for e in tr_ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
# step boundaries: adam_step_kernel marks the end of a step
ends = [e for e in ev if "adam_step" in e["name"]]
firsts = [e for e in ev if "cvt_pad_bf16" in e["name"]]
print("adam ends (ms):", [round((e["ts"] + e["dur"] - t0) / 1e3, 3) for e in ends])
for a, b in zip(ends[:-1], ends[1:]):
    lo, hi = a["ts"] + a["dur"], b["ts"] + b["dur"]
    inside = [e for e in ev if lo <= e["ts"] < hi]
    k = [e for e in inside if e["cat"] == "kernel"]
    first_k = min(e["ts"] for e in k)
    print(f"step: {(hi - lo) / 1e3:.3f} ms; idle before first kernel {(first_k - lo):.1f} us; memcpy in window: "
          + ", ".join(f"{e['name'][:24]} {e['dur']:.0f}us@{(e['ts'] - lo) / 1e3:.2f}ms" for e in inside if e["cat"] == "gpu_memcpy" and e["dur"] > 5))
cpu = [e for e in tr_ev if e.get("cat") == "cuda_runtime" and "GraphLaunch" in e.get("name", "")]
print("cudaGraphLaunch host durations (us):", [round(e["dur"]) for e in cpu])
