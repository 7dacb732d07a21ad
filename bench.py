"""bench.py — train utterances/sec of the LAS supervised train step (BASELINE.json config 2:
si284-shaped, batch 32 per GPU, T~1000 frames, bf16 MMA operands / f32 state) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path (one rank per GPU)
  python bench.py --impl reference ...                         the reference's own CPU path (oracle/_ref: its unmodified
                                                               model.py + the step body of solver.py:375-385; the
                                                               oracle port only where oracle/_ref is absent)

A "step" = forward + loss + backward + clip + AMSGrad on one synthetic WSJ-shaped batch
(SURVEY.md §8(d)). `value`: inputs already resident in HBM (graph replay only); `e2e`: the same
step through SupervisedTrainer.step() with pinned-host inputs, H2D copies and a loss read-back
inside the timed region. Prints ONE JSON line on rank 0.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "semi-supervised-asr_b200"

CFG = dict(input_dim=249, enc_hidden_dim=320, enc_n_layers=3, subsample=[2, 2, 2], dec_hidden_dim=320, att_dim=320,
           conv_channels=10, conv_kernel_size=100, att_odim=320, embedding_dim=128, ls_weight=0.05, V=34,
           lr=5e-4, weight_decay=1e-6, max_grad_norm=5.0)
# algorithmic work per utterance-step, SURVEY.md §8(d) / BASELINE.md §4 (config 2)
FLOP_PER_UTT = 19.81e9
HBM_BYTES_PER_UTT = 51e6
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the serial kernels at config 2, from the committed
# `ncu --set full` captures (profiles/*_ncu_full_summary.txt); None = not captured for the current build
NCU_TRAFFIC = {
    "lstm_persist_fwd": (329358848 + 366706176, "profiles/r03_ncu_full_summary.txt, layer 0"),
    "lstm_persist_bwd": (None, None),
    "dec_persist_fwd": (None, None),
    "dec_persist_bwd": (None, None),
}
try:                      # refreshed by tools/ncu_summary.py from the round's capture
    NCU_TRAFFIC.update({k: tuple(v) for k, v in json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).items()})
except Exception:
    pass


def synth_batch(rng, B, Tmax, D, V):
    """T_b ~ U[0.4 Tmax, Tmax], T_0 = Tmax, sorted descending; x ~ N(0,1), zeros
    past T_b; L_b = clip(round(0.125 T_b), 2, 250), tokens ~ U[3, V)."""
    lens = sorted([Tmax] + [int(rng.randint(int(0.4 * Tmax), Tmax + 1)) for _ in range(B - 1)], reverse=True)
    x = np.zeros((B, Tmax, D), dtype=np.float32)
    ys = []
    for b, l in enumerate(lens):
        x[b, :l] = rng.standard_normal((l, D)).astype(np.float32)
        L = int(np.clip(round(0.125 * l), 2, 250))
        ys.append(rng.randint(3, V, size=L).astype(np.int64))
    return x, lens, ys


def synth_shard(seed, B, world, rank, Tmax, D, V):
    """One GLOBAL batch of B * world utterances under synth_batch's length law (the same on every rank), dealt over
    the ranks the way data.shard_items deals a real batch (sorted by frame count, round-robin), so the ranks of a
    data-parallel run get DIFFERENT (Tmax, Lmax) geometries and the straggler cost of uneven shards is in the
    measurement. Only this rank's utterances are materialised. -> (x [B, T_r, D], lens, ys, all_ys)"""
    rng = np.random.RandomState(seed)
    n = B * world
    lens = sorted([Tmax] + [int(rng.randint(int(0.4 * Tmax), Tmax + 1)) for _ in range(n - 1)], reverse=True)
    all_ys = [rng.randint(3, V, size=int(np.clip(round(0.125 * l), 2, 250))).astype(np.int64) for l in lens]
    mine = list(range(n))[rank::world]                   # data.shard_items on a batch already sorted descending
    T = lens[mine[0]]
    x = np.zeros((len(mine), T, D), dtype=np.float32)
    for b, i in enumerate(mine):
        x[b, :lens[i]] = np.random.RandomState(seed * 100003 + i).standard_normal((lens[i], D)).astype(np.float32)
    return x, [lens[i] for i in mine], [all_ys[i] for i in mine], all_ys


def labeldist_of(ys, V):
    cnt = np.zeros(V)
    for y in ys:
        for t in y:
            cnt[t] += 1
    cnt[2] += len(ys)
    cnt[0] = cnt[1] = 0
    return cnt / cnt.sum()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        if os.environ.get("LAS_BENCH_NO_SMI"):     # diagnosis only: is the sampler perturbing the host loop?
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def _cfg_kwargs(ld, dropout):
    return dict(input_dim=CFG["input_dim"], enc_hidden_dim=CFG["enc_hidden_dim"], enc_n_layers=CFG["enc_n_layers"],
                subsample=CFG["subsample"], dropout_rate=dropout, dec_hidden_dim=CFG["dec_hidden_dim"],
                att_dim=CFG["att_dim"], conv_channels=CFG["conv_channels"], conv_kernel_size=CFG["conv_kernel_size"],
                att_odim=CFG["att_odim"], embedding_dim=CFG["embedding_dim"], output_dim=CFG["V"],
                ls_weight=CFG["ls_weight"], labeldist=ld)


def cpu_reference_run(steps, warmup, threads, B=32, Tmax=1000, dropout=0.3):
    """The reference's CPU path for the same step on the same synthetic batch law (B=32, Tmax=1000, dropout 0.3).
    kind "reference": oracle/_ref -- the reference's UNMODIFIED model.py (E2E on torch CPU fp32, nn.LSTM over packed
    sequences) driven by the step body of solver.py:375-385 (forward, -mean(log_probs), backward, clip_grad_norm_,
    Adam(amsgrad)); kind "port": the oracle restatement, only where oracle/_ref is absent.
    -> (utt/s, s/step, sample description, kind)"""
    torch.set_num_threads(threads)
    rng = np.random.RandomState(1234)
    x, lens, ys = synth_batch(rng, B, Tmax, CFG["input_dim"], CFG["V"])
    ld = labeldist_of(ys, CFG["V"])
    torch.manual_seed(1234)
    xt = torch.from_numpy(x)
    from oracle import build_ref
    times = []
    if build_ref.available():
        kind = "reference"
        R = build_ref.import_reference()
        assert not torch.cuda.is_available(), "the reference moves itself to a visible GPU (utils.cc): hide it"
        m = R.E2E(**_cfg_kwargs(ld, dropout))
        m.train()
        opt = torch.optim.Adam(m.parameters(), lr=CFG["lr"], weight_decay=CFG["weight_decay"], amsgrad=True)  # solver.py:152
        yt = [torch.from_numpy(y) for y in ys]
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            _, log_probs, _, _ = m(xt, lens, yt, tf_rate=1.0, sample=False)      # solver.py:375
            loss = -torch.mean(log_probs)                                          # solver.py:377
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=CFG["max_grad_norm"])   # solver.py:384
            opt.step()
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    else:
        kind = "port"
        from oracle import las_oracle as O
        M = importlib.import_module(PKG + ".model")
        m = M.E2E(**_cfg_kwargs(ld, 0.0))
        P = {k: v.detach().clone() for k, v in m.state_dict().items()}
        state = {}
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            _, _, _, new = O.supervised_step(xt, lens, ys, P, state, CFG["subsample"], CFG["ls_weight"], ld,
                                             lr=CFG["lr"], weight_decay=CFG["weight_decay"],
                                             max_grad_norm=CFG["max_grad_norm"], fast=True)
            for k in new:
                P[k] = new[k]
                if k.startswith("attention."):
                    P["decoder." + k] = new[k]
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    sample = (f"{len(times)} step(s) of the full workload batch (B={B}, Tmax={Tmax}, dropout {dropout if kind == 'reference' else 0.0}), "
              f"torch {torch.__version__} CPU fp32, {threads} threads")
    return B * len(times) / sum(times), float(np.mean(times)), sample, kind


def cpu_baseline_subprocess(cores):
    """The CPU arm in a child process with the GPUs hidden (the reference's utils.cc moves every module to a visible
    GPU; CUDA is already initialised in this process). -> the child's JSON line."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    res = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         env=env, capture_output=True, text=True, timeout=1200)
    for line in reversed(res.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("CPU baseline child failed: " + res.stderr[-2000:])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--tmax", type=int, default=1000)
    ap.add_argument("--dropout", type=float, default=0.3, help="dropout_rate (config.yaml:30 of the reference: 0.3)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-step", action="store_true", help="run ONE eager step between cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)")
    args = ap.parse_args()
    if os.environ.get("LAS_BENCH_WATCHDOG"):       # stack dump + exit if a (multi-rank) run stops making progress
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["LAS_BENCH_WATCHDOG"]), exit=True)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    workload = f"supervised LAS train step, si284-shaped synthetic, B={args.batch}/GPU, Tmax={args.tmax}, 3xpBLSTM-320 [2,2,2], V=34"

    if args.impl == "reference":
        if rank != 0:
            return
        if torch.cuda.is_available() and os.environ.get("CUDA_VISIBLE_DEVICES", None) != "":
            # the reference's utils.cc() moves every module to a visible GPU: re-run with the GPUs hidden
            env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
            sys.exit(subprocess.call([sys.executable] + sys.argv, env=env))
        # bounded sample: the unmodified reference needs ~1-2 minutes per B=32 step on 8-16 cores (its decoder loop is
        # 126 Python iterations of Conv2d(1, 10, (1, 201)) + small matmuls, and autograd replays them)
        steps = max(1, min(args.steps, 2))
        warm = 0
        v, sec, sample, kind = cpu_reference_run(steps, warm, cores, B=args.batch, Tmax=args.tmax, dropout=args.dropout)
        print(json.dumps({
            "impl": "reference", "metric": "train utterances/sec", "value": v, "unit": "utt/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "global_batch": args.batch, "dropout": args.dropout if kind == "reference" else 0.0,
                       "sample": sample},
            "cpu_baseline": {"value": v, "unit": "utt/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the product path has no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    M = importlib.import_module(PKG + ".model")
    E = importlib.import_module(PKG + ".engine")
    OPT = importlib.import_module(PKG + ".optim")
    LIB = importlib.import_module(PKG + "._lib")

    nb = 4                                             # distinct synthetic global batches, cycled
    shards = [synth_shard(1234 + k, args.batch, world, rank, args.tmax, CFG["input_dim"], CFG["V"]) for k in range(nb)]
    batches = [s[:3] for s in shards]
    ld = labeldist_of([y for s in shards for y in s[3]], CFG["V"])
    torch.manual_seed(1234)
    m = M.E2E(**_cfg_kwargs(ld, args.dropout)).to(dev)
    if world > 1:                                      # identical initial weights on every rank
        for p in m.parameters():
            torch.distributed.broadcast(p.data, 0)
    opt = OPT.FusedAdam(m.parameters(), lr=CFG["lr"], weight_decay=CFG["weight_decay"], amsgrad=True)
    tr = E.SupervisedTrainer(m, opt, max_grad_norm=CFG["max_grad_norm"], use_graph=not args.no_graph)
    pinned = [(torch.from_numpy(x).pin_memory(), lens, [torch.from_numpy(y) for y in ys]) for x, lens, ys in batches]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (first step of a geometry is eager, second captures the graph)
    W = max(args.warmup, 3)
    # the ranks of a data-parallel run see different (Tmax, Lmax) per batch: every distinct geometry is met three times
    # before anything is timed (eager, capture, replay), so that no graph capture falls into a timed region
    # (the same count on EVERY rank -- each step ends in a collective -- so it must not depend on this rank's batches)
    if world > 1:
        W = max(W, 3 * nb)
    losses = []
    # through the same public API the end-to-end region uses (steps()): its upload slots, copy / read-back streams
    # and pinned staging are created here, not inside a timed region
    for loss, _ in tr.steps(pinned[i % nb] for i in range(W)):
        losses.append(float(loss))
    if args.profile_step:
        tr.use_graph = False
        key = tr.stage(*pinned[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        tr.run(key)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profile_step": "done", "loss": losses}))
        return
    c0 = LIB.lib().las_launch_count()
    key = tr.stage(*pinned[0])
    if args.no_graph:
        tr.run(key)
        launches = int(LIB.lib().las_launch_count() - c0)
    else:
        launches = None

    # ---- device-resident timing: inputs staged once per distinct batch, then replay only
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the staging arenas are shared by all batches: every batch keeps a device-resident copy, and a step starts with
    # the device-to-device copy of its inputs into the graph's static buffers (~10 us; no host traffic)
    keys, snaps = [], []
    for p in pinned:
        k = tr.stage(*p)
        keys.append(k)
        snaps.append(tr.snapshot(k))
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        tr.run(tr.restore(keys[i % nb], snaps[i % nb]))
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # ---- end-to-end timing: pinned host inputs, H2D + step + loss read-back every step
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    # public API: SupervisedTrainer.steps() over host batches -- the H2D copy of batch i+1 is issued while step i
    # runs; every step's (loss, grad-norm) pair is read back to the host (on the copy stream, handed out one step
    # late so the GPU never waits for the host)
    for loss, _ in tr.steps(pinned[i % nb] for i in range(args.steps)):
        losses.append(loss.item())
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    # ---- the same, starting one level further up: utterances in the reference's pickle item format (feature [T, D]
    # float32 array, token list) -> data.BatchLoader (sorting, zero-padding and collation on a background thread
    # straight into page-locked buffers) -> SupervisedTrainer.steps. One GPU only (every rank would need every item).
    ms_loader = None
    if world == 1:
        D = importlib.import_module(PKG + ".data")
        items = [(np.ascontiguousarray(x[b, :l]), ys[b].tolist()) for x, lens_b, ys in batches for b, l in enumerate(lens_b)]
        loader = D.BatchLoader(items, args.batch, shuffle=False, drop_last=True, prefetch=2)
        epochs = (args.steps + nb - 1) // nb

        def stream():
            for _ in range(epochs):
                yield from loader
        for _ in tr.steps(b for b in loader):          # one untimed pass: ring buffers allocated and pinned
            pass
        barrier()
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e4.record()
        n_l = 0
        for loss, _ in tr.steps(stream()):
            n_l += 1
        e5.record()
        barrier()
        ms_loader = e4.elapsed_time(e5) / n_l
    clocks = sampler.stop()
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    x0, _, ys0 = pinned[0]
    Lmax = max(len(y) for y in ys0) + 1
    h2d = x0.numel() * 4 + args.batch * 4 + args.batch * (2 * Lmax + 1) * 8
    utt = args.batch * world * args.steps
    value = utt / (ms * 1e-3)
    e2e = utt / (ms_e2e * 1e-3)

    # launches per step (ours): measured with graphs off on one extra eager step (every rank runs it: the
    # step contains the gradient all-reduce)
    if launches is None:
        tr2_graph = tr.use_graph
        tr.use_graph = False
        c0 = LIB.lib().las_launch_count()
        tr.run(tr.restore(keys[0], snaps[0]))
        torch.cuda.synchronize()
        launches = int(LIB.lib().las_launch_count() - c0)
        tr.use_graph = tr2_graph
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    step_us = ms / args.steps * 1e3
    # every rank takes part: the profiled step contains the gradient all-reduce
    tr.restore(keys[0], snaps[0])
    roofs = kernel_rooflines(tr, keys[0], args.batch, args.tmax, step_us, hbm_peak, peaks.get("bf16_tflops_sustained", 1400.0),
                             "measured" if peaks else "fallback")
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    roof = max((r for r in roofs if r["bound"] == "hbm"), key=lambda r: r["us_per_step"])
    out = {
        "metric": "train utterances/sec", "value": value, "unit": "utt/s", "n_gpus": world, "steps": args.steps,
        "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload, "global_batch": args.batch * world, "parallelism": f"dp{world}",
                   "precision": "bf16 MMA operands, f32 accumulate/state/master weights", "dropout": args.dropout,
                   "cuda_graph": not args.no_graph,
                   "batches": f"{nb} global batches of {args.batch * world} utterances cycled, each dealt over the ranks by "
                              "sorted round-robin (data.shard_items): per-rank Tmax/Lmax differ",
                   "l2": "working set per step (>1 GB of saved activations) exceeds the 126 MB L2; no flush"},
        "e2e": {"value": e2e, "unit": "utt/s", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": 8},
        "e2e_from_items": (None if ms_loader is None else
                           {"value": args.batch / (ms_loader * 1e-3), "unit": "utt/s", "ms_per_step": ms_loader,
                            "what": "pickle-format items -> data.BatchLoader(prefetch=2: background collation into pinned "
                                    "buffers) -> SupervisedTrainer.steps; H2D and loss read-back every step"}),
        "gpu_launches": launches * args.steps,
        "gpu_launches_per_step": launches,
        "clocks": clocks,
        "step_roofline": {"algorithmic_tflops": FLOP_PER_UTT * args.batch / (ms / args.steps * 1e-3) / 1e12,
                          "frac_of_bf16_sustained": FLOP_PER_UTT * args.batch / (ms / args.steps * 1e-3) / 1e12 / tf_peak,
                          "algorithmic_hbm_gbs": HBM_BYTES_PER_UTT * args.batch / (ms / args.steps * 1e-3) / 1e9},
        "roofline": roof,
        "roofline_by_kernel": roofs,
        "loss_first_last": [losses[0], losses[-1]],
    }
    if not args.no_cpu_baseline:
        child = cpu_baseline_subprocess(cores)
        out["cpu_baseline"] = dict(child["cpu_baseline"], s_per_step=child["ms_per_step"] * 1e-3)
    print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


def kernel_rooflines(tr, key, B, Tmax, step_us, hbm_peak, tf_peak, which):
    """Per-kernel-family roofline entries for the serial kernels and the GEMM family. Durations are measured HERE:
    one eager (graph-off) step of the same trainer with a CUDA-event pair recorded around every C-ABI call on the
    stream it launches on (`_lib.TIMER`); shares are relative to the graph-replay step time of the timed region.
    Algorithmic bytes follow SURVEY.md 8(d)'s bf16 "reserve-space" model: per (utterance, timestep,
    direction) a recurrence pass moves 20 H bytes -- forward: gate pre-activations 8H read, h 2H + gates/cell 10H
    written; backward: gates/cell 10H + dy 2H read, gate gradients 8H written; a decoder pass streams enc_h and P
    (Te x (H + A) bf16) once per step and utterance."""
    LIB = importlib.import_module(PKG + "._lib")
    H, A = CFG["enc_hidden_dim"], CFG["att_dim"]
    use_graph = tr.use_graph
    tr.use_graph = False
    tr.run(key)                                  # warm (allocator, lazy attributes) with the timer off
    torch.cuda.synchronize()
    LIB.TIMER = []
    try:
        tr.run(key)
        torch.cuda.synchronize()
        rows = [(n, e0.elapsed_time(e1) * 1e3) for n, e0, e1 in LIB.TIMER]
    finally:
        LIB.TIMER = None
        tr.use_graph = use_graph
    fam = {}
    for n, us in rows:
        f = {"las_lstm_persist_fwd": "lstm_persist_fwd", "las_lstm_persist_bwd": "lstm_persist_bwd",
             "las_dec_fwd": "dec_persist_fwd", "las_dec_bwd": "dec_persist_bwd", "las_gemm_bf16_ws": "gemm_tcgen05"}.get(n)
        if f:
            fam.setdefault(f, []).append(us)
    # timesteps per pass over the pyramid (config 2: 1000 + 500 + 250), decoder geometry
    T_l, t = [], Tmax
    for sub in CFG["subsample"]:
        T_l.append(t)
        t = (t + 1) // sub if sub > 1 else t
    Te, Lp1 = t, int(np.clip(round(0.125 * Tmax), 2, 250)) + 1
    rec_bytes = sum(B * tl * 2 * 20 * H for tl in T_l) + len(T_l) * 2 * 4 * H * H * 2
    dec_bytes = B * Lp1 * Te * (H + A) * 2
    alg = {"lstm_persist_fwd": rec_bytes, "lstm_persist_bwd": rec_bytes, "dec_persist_fwd": dec_bytes, "dec_persist_bwd": dec_bytes}
    names = {"lstm_persist_fwd": "lstm_persist_fwd_kernel (pBLSTM recurrence, all layers: one launch per layer)",
             "lstm_persist_bwd": "lstm_persist_bwd_kernel (pBLSTM BPTT, all layers: one launch per layer)",
             "dec_persist_fwd": "dec_persist_fwd_kernel (attention decoder, all L+1 steps in one launch)",
             "dec_persist_bwd": "dec_persist_bwd_kernel (attention decoder BPTT, one launch)"}
    out = []
    for f in ("lstm_persist_bwd", "lstm_persist_fwd", "dec_persist_bwd", "dec_persist_fwd"):
        if f not in fam:
            continue
        us = sum(fam[f])
        ach = alg[f] / (us * 1e-6) / 1e9
        traffic, src = NCU_TRAFFIC.get(f, (None, None))
        steps_serial = sum(T_l) if f.startswith("lstm") else Lp1
        out.append({"kernel": names[f], "bound": "hbm", "launches_per_step": len(fam[f]), "us_per_step": us,
                    "share_of_step": us / step_us, "achieved": ach, "peak": hbm_peak, "peak_source": which, "unit": "GB/s",
                    "frac": ach / hbm_peak, "algorithmic_bytes": alg[f], "traffic": traffic, "traffic_source": src,
                    "us_per_timestep": us / steps_serial,
                    "note": "latency-bound serial chain (dependent timesteps; DESIGN.md 4b/5): the HBM fraction is low by construction"})
    if "gemm_tcgen05" in fam:
        us = sum(fam["gemm_tcgen05"])
        flops = 0.488 * FLOP_PER_UTT * B          # dense GEMMs are 48.8 % of the step's algorithmic FLOPs
        ach = flops / (us * 1e-6) / 1e12
        out.append({"kernel": "gemm_kernel (tcgen05 + TMA; every dense contraction of the step)", "bound": "tensor",
                    "launches_per_step": len(fam["gemm_tcgen05"]), "us_per_step": us, "share_of_step": us / step_us,
                    "achieved": ach, "peak": tf_peak, "peak_source": which, "unit": "TFLOP/s", "frac": ach / tf_peak,
                    "traffic": None,
                    "note": "sum of per-launch durations incl. side-stream launches that run concurrently with the serial kernels"})
    return out


if __name__ == "__main__":
    main()
