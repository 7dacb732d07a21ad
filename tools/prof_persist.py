"""One BLSTM layer (B=32, T=1000, H=320) forward + BPTT through the C-ABI, for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import pkg
Fn = pkg("functional"); LIB = pkg("_lib")
B, T, H = int(os.environ.get('B', 32)), int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 320
dev = torch.device("cuda")
torch.manual_seed(0)
xproj = torch.randn(B * T, 8 * H, device=dev) * 0.1
w = [torch.randn(4 * H, H, device=dev) * 0.05 for _ in range(2)]
whh = torch.cat([Fn.pack_afrag(w[0], 3, H), Fn.pack_afrag(w[1], 3, H)])
wT = torch.cat([Fn.pack_whhT(w[0])[0], Fn.pack_whhT(w[1])[0]])
lens = torch.full((B,), T, device=dev, dtype=torch.int32)
y = torch.zeros(B, T, 2 * H, device=dev, dtype=torch.bfloat16)
hprev = torch.empty_like(y)
rec = torch.empty(2 * B * T * H, 4, device=dev, dtype=torch.int32)
dy = torch.randn(B, T, 2 * H, device=dev) * 0.01
dG = torch.empty(B * T, 8 * H, device=dev, dtype=torch.bfloat16)
def fwd():
    Fn.call("las_lstm_persist_fwd", Fn.ptr(xproj), Fn.ptr(whh), Fn.ptr(lens), B, T, H, 2, Fn.ptr(y), T * 2 * H, 2 * H, 0,
            Fn.ptr(hprev), T * 2 * H, 2 * H, Fn.ptr(rec))
def bwd():
    Fn.call("las_lstm_persist_bwd", Fn.ptr(dy), T * 2 * H, 2 * H, 0, Fn.ptr(wT), Fn.ptr(lens), B, T, H, 2, Fn.ptr(rec),
            Fn.ptr(dG), T * 8 * H, 8 * H)
for _ in range(2):
    fwd(); bwd()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); fwd(); e[1].record(); bwd(); e[2].record(); torch.cuda.synchronize()
print(f"fwd {e[0].elapsed_time(e[1])*1e3/T:.3f} us/step  bwd {e[1].elapsed_time(e[2])*1e3/T:.3f} us/step  (B={B}, T={T}, LAS_FWD_CS={os.environ.get('LAS_FWD_CS')}); "
      f"resident clusters: fwd default {LIB.lib().las_lstm_persist_max_clusters(0, H)}, fwd 7-CTA {LIB.lib().las_lstm_persist_max_clusters(2, H)}, "
      f"fwd 10-CTA {LIB.lib().las_lstm_persist_max_clusters(3, H)}, bwd {LIB.lib().las_lstm_persist_max_clusters(1, H)}")
if os.environ.get("LAS_TRACE"):
    dbg = torch.zeros(128, device=dev, dtype=torch.int64)
    LIB.lib().las_set_debug_buffer(dbg.data_ptr())
    fwd(); bwd(); torch.cuda.synchronize()
    LIB.lib().las_set_debug_buffer(None)
    d = dbg.cpu().view(2, 8, 8)
    for k, name in enumerate(["fwd", "bwd"]):
        print(name, "phase deltas (cycles) per step: slots 0..7, then step total")
        for s in range(7):
            row = d[k, s]
            deltas = [int(row[i + 1] - row[i]) if row[i + 1] and row[i] else -1 for i in range(7)]
            print("  ", deltas, "| step", int(d[k, s + 1, 0] - row[0]))
