"""Train-step engines behind `Solver.sup_train_one_epoch` (solver.py:360-393): forward, loss =
-mean(log_probs), backward, clip_grad_norm_(max_grad_norm), Adam(amsgrad) step — with the whole
step (several thousand per-timestep launches) captured into one CUDA graph per batch geometry
(B, Tmax, Lmax) and replayed.

Data parallelism (one process per GPU): every rank runs the same step on its shard; the flat
gradient buffer is summed with ONE NCCL all-reduce between backward and the optimiser, and the
update uses grad_scale = 1/world_size (DDP averaging). There is no other exchange on this path.
"""
import os

import numpy as np
import torch

from . import functional as Fn


def build_targets(ys, bos, eos, pad, L_min=0):
    """model.py:301-306: ys_in = [BOS, y], ys_out = [y, EOS], both EOS-padded; plus one guard
    column (PAD) so that every per-step buffer has L+1 rows. Returns (ys_in [B, L+1], ys_out [B, L]).
    L_min: pad to at least this many steps (the global Lmax + 1 of a data-parallel batch, global-exact mode)."""
    ys_host = [y.detach().cpu().numpy().astype(np.int64) if torch.is_tensor(y) else np.asarray(y, np.int64)
               for y in ys]
    B = len(ys_host)
    L = max(max(len(y) for y in ys_host) + 1, int(L_min))
    ys_in = np.full((B, L + 1), eos, dtype=np.int64)
    ys_out = np.full((B, L), eos, dtype=np.int64)
    for b, y in enumerate(ys_host):
        ys_in[b, 0] = bos
        ys_in[b, 1:1 + len(y)] = y
        ys_out[b, :len(y)] = y
    ys_in[:, L] = pad
    return ys_in, ys_out


class _Arena:
    """Named device / page-locked host buffers that only ever grow to the largest extent requested so far; every
    batch geometry gets VIEWS of them (all starting at the buffer base). Memory is therefore bounded by the largest
    batch, not by the number of distinct (B, Tmax, Lmax) seen (reference batches almost never repeat a geometry).
    Growing a buffer moves it: `gen` is bumped so that owners of captured graphs (which hold raw addresses) drop
    them. `reserve()` the configured maxima up front and nothing ever moves."""

    def __init__(self, dev):
        self.dev = dev
        self.bufs = {}
        self.gen = 0

    def get(self, name, shape, dtype, pinned=False):
        n = int(torch.Size(shape).numel())
        buf = self.bufs.get(name)
        if buf is None or buf.numel() < n:
            if buf is not None:
                torch.cuda.synchronize(self.dev)     # copies / kernels in flight may still use the old storage
                self.gen += 1
                n_alloc = max(n, int(buf.numel() * 1.25))
            else:
                n_alloc = n
            if pinned:
                buf = torch.zeros(max(n_alloc, 1), dtype=dtype).pin_memory()
            else:
                buf = torch.zeros(max(n_alloc, 1), device=self.dev, dtype=dtype)
            self.bufs[name] = buf
        assert buf.dtype == dtype
        return buf[:n].view(shape)

    def nbytes(self):
        return sum(b.numel() * b.element_size() for b in self.bufs.values() if b.is_cuda)


class _Views:
    """Views of one batch geometry into the arena: the graph's static inputs, their pinned host staging, and the
    two upload slots of the pipelined path."""

    def __init__(self, arena, B, T, D, L, prefix=""):
        g = lambda name, shape, dtype, pinned=False: arena.get(prefix + name, shape, dtype, pinned)
        i32, i64, f32 = torch.int32, torch.int64, torch.float32
        self.x = g("x", (B, T, D), f32)
        self.lens = g("lens", (B,), i32)
        self.ys_in = g("ys_in", (B, L + 1), i64)
        self.ys_out = g("ys_out", (B, L), i64)
        self.h_lens = g("h_lens", (B,), i32, True)
        self.h_ys_in = g("h_ys_in", (B, L + 1), i64, True)
        self.h_ys_out = g("h_ys_out", (B, L), i64, True)
        self.inv = g("inv", (1,), f32)          # global-exact data parallelism: 1 / (global B * (global Lmax + 1))


class _Slot:
    """State of one upload slot (its buffers are arena views made per batch)."""

    def __init__(self):
        self.ready = torch.cuda.Event()    # upload finished (recorded on the copy stream)
        self.consumed = None               # device copy into the static buffers finished (recorded on the main stream)
        self.used = False
        self.views = None


class _GraphCache:
    """LRU of captured step graphs, keyed by batch geometry, all captured into ONE shared memory pool (a private
    pool per graph would hold ~1 GB of activations each). A geometry is run eagerly on first sight, captured on
    the second, evicted when more than `max_graphs` others have been used since. Entries captured against an older
    arena generation or other optimiser hyper-parameters (they are baked into the captured launches by value) are
    dropped."""

    def __init__(self, max_graphs=8, max_seen=4096):
        import collections
        self.max_graphs, self.max_seen = int(max_graphs), int(max_seen)
        self.graphs = collections.OrderedDict()
        self.seen = collections.OrderedDict()
        self.pool = None
        self.captures, self.evictions = 0, 0

    def lookup(self, key, stamp):
        ent = self.graphs.get(key)
        if ent is not None and ent["stamp"] != stamp:
            del self.graphs[key]
            ent = None
        if ent is not None:
            self.graphs.move_to_end(key)
        return ent

    def sighting(self, key):
        """Number of earlier sightings of this geometry (bounded memory: oldest keys are forgotten)."""
        n = self.seen.pop(key, 0)
        self.seen[key] = n + 1
        while len(self.seen) > self.max_seen:
            self.seen.popitem(last=False)
        return n

    def store(self, key, ent):
        self.graphs[key] = ent
        self.captures += 1
        while len(self.graphs) > max(1, self.max_graphs):
            self.graphs.popitem(last=False)
            self.evictions += 1

    def clear(self):
        self.graphs.clear()

    def pool_handle(self):
        """The shared pool. torch releases a pool when the last graph captured into it dies (and asserts if the handle
        is used again), which evicting or invalidating every entry would do: a trivial anchor graph keeps it alive."""
        if self.pool is None:
            self.pool = torch.cuda.graph_pool_handle()
            self._anchor = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._anchor, pool=self.pool):
                self._anchor_out = torch.zeros(8, device="cuda")
        return self.pool


def _hyper_stamp(opt):
    g = opt.param_groups[0]
    return (float(g["lr"]), tuple(float(b) for b in g["betas"]), float(g["eps"]), float(g["weight_decay"]))


def _host_lens(ilens):
    return [int(l) for l in (ilens.tolist() if torch.is_tensor(ilens) else ilens)]


class SupervisedTrainer:
    def __init__(self, model, optimizer, max_grad_norm=5.0, use_graph=True, process_group=None, max_graphs=8,
                 global_exact=False, overlap_allreduce=False):
        """global_exact (data parallel only): every shard is padded to the GLOBAL Tmax / Lmax and its
        loss is -sum(log_probs) / (global B * (global Lmax + 1)); the summed gradients then equal the single-GPU
        gradient of the concatenated batch up to floating-point reassociation (the reference's loss couples the
        utterances of a batch through its padded extents). Costs one small host-synchronous
        all-reduce per step to agree on the extents. Default: standard DDP (each rank's loss is the reference loss
        of its shard, gradients averaged)."""
        self.global_exact = bool(global_exact)
        self.global_shape = None          # (T, L, B) override used instead of the collective (tests)
        # Data parallel, overlap_allreduce=True: the encoder's first layer gets an autograd graph (and a captured CUDA
        # graph) of its own, so that the gradients of everything else -- 80 % of the bytes -- are all-reduced on a
        # communication stream, outside any capture, WHILE that layer's BPTT (the last and longest kernel of the backward
        # pass) runs. Measured on 2 B200s it LOSES (8.31 vs 8.07 ms per step end to end): the graph boundary forces the
        # join of the weight-gradient stream before layer 0's BPTT, which puts layer 1's weight-gradient GEMMs (otherwise
        # hidden under that kernel) on the critical path, and costs one more graph launch -- more than the ~0.1 ms
        # exchange it hides. Off by default; LAS_OVERLAP=1 switches it on.
        self.overlap_allreduce = bool(overlap_allreduce) or os.environ.get("LAS_OVERLAP", "0") == "1"
        self.force_split = False          # tests: take the two-part step on one GPU (no collective is issued)
        self._stage_ev = None             # H2D copies of the last stage() (its pinned staging is shared by all geometries)
        self.comm_stream = None
        self._buckets = None
        self.model = model
        self.opt = optimizer
        self.max_grad_norm = max_grad_norm
        self.use_graph = use_graph and os.environ.get("LAS_NO_GRAPH", "0") != "1"   # debugging switch
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.arena = None
        self.cache = _GraphCache(max_graphs)
        self.slots = None
        self.launches_per_step = None
        self.update_graph, self.update_norm, self.update_stamp = None, None, None
        self.cap_stream = None
        self.copy_stream = None
        self.rb_dev, self.rb_host, self.rb_stream = None, None, None
        # solver.py:370-373 adds N(0, gaussian_std) to the padded features on the host (8 M numpy normals per batch,
        # ~10 step times); set this and the same noise is drawn on the device right after the upload
        self.input_noise_std = 0.0

    # ---- the step body: everything below runs on the current stream, no host sync
    def _fwd_bwd(self, st, L):
        m = self.model
        enc = m.encoder.enc2
        # bf16 copies / fragment packs of all weights: on the side stream, so that only layer 0's stay on the
        # critical path and the rest overlaps the first recurrence
        jobs = enc.prep_jobs() + m.decoder.prep_jobs(0)
        jobs.sort(key=lambda j: "bwd" in j[0])          # stable: forward operands first, in order of use
        Fn.prepare_ahead(jobs)
        enc_h = enc.forward_dev(st.x, st.lens)
        enc_lens = enc.out_lens_dev(st.lens)
        _, logp, _, _ = m.decoder.forward_dev(enc_h, enc_lens, st.ys_in, st.ys_out, L, 0)
        if self.global_exact:
            loss = -torch.sum(logp) * st.inv[0]                    # this shard's share of the global mean
        else:
            loss = -torch.mean(logp)                               # solver.py:377 (ALL B x (Lmax+1) positions)
        self.opt.zero_grad()
        with Fn.deferred_wgrad():      # weight-gradient contractions run on a side stream, joined on exit
            loss.backward()
        return loss.detach()

    # ---- data parallel with the exchange overlapped: the step in two parts
    def _split(self):
        return ((self.world > 1 and self.overlap_allreduce) or self.force_split) and len(self.model.encoder.enc2.layers) > 1

    def _part1(self, st, L):
        """Forward, loss, backward of everything except encoder layer 0. -> (loss, y0, y0_leaf)"""
        m = self.model
        enc = m.encoder.enc2
        jobs = enc.prep_jobs() + m.decoder.prep_jobs(0)
        jobs.sort(key=lambda j: "bwd" in j[0])
        Fn.prepare_ahead(jobs)
        enc_h, y0, y0_leaf = enc.forward_dev_split(st.x, st.lens)
        enc_lens = enc.out_lens_dev(st.lens)
        _, logp, _, _ = m.decoder.forward_dev(enc_h, enc_lens, st.ys_in, st.ys_out, L, 0)
        loss = -torch.sum(logp) * st.inv[0] if self.global_exact else -torch.mean(logp)
        self.opt.zero_grad()
        with Fn.deferred_wgrad():
            loss.backward()
        return loss.detach(), y0, y0_leaf

    def _part2(self, y0, y0_leaf):
        """Backward of encoder layer 0 (BPTT + its weight gradients)."""
        with Fn.deferred_wgrad():
            y0.backward(y0_leaf.grad)

    def _bucket_ranges(self):
        """(ranges exchanged during part 2, ranges exchanged after it) as views of the flat gradient: the second set is
        what part 2 writes -- encoder layer 0's LSTM and projection parameters."""
        if self._buckets is None:
            enc = self.model.encoder.enc2
            late = {id(p) for p in list(enc.layers[0].parameters()) + list(enc.project_layers[0].parameters())}
            a, b = [], []
            for p, o in zip(self.opt.params, self.opt.offsets):
                dst = b if id(p) in late else a
                if dst and dst[-1][1] == o:
                    dst[-1][1] = o + (p.numel() + 3) // 4 * 4
                else:
                    dst.append([o, o + (p.numel() + 3) // 4 * 4])
            g = self.opt.flat_grad
            self._buckets = ([g[s:e] for s, e in a], [g[s:e] for s, e in b])
        return self._buckets

    def _exchange_overlapped(self, run_part2):
        """all-reduce of the early bucket on the communication stream while `run_part2()` executes on the main stream,
        then the late bucket on the main stream."""
        dev = self.opt.flat_grad.device
        main = torch.cuda.current_stream(dev)
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=dev)
        early, late = self._bucket_ranges()
        self.comm_stream.wait_stream(main)                   # part 1 (and its weight gradients) is complete
        with torch.cuda.stream(self.comm_stream):
            for t in early:
                if self.world > 1:
                    torch.distributed.all_reduce(t, group=self.pg)
        run_part2()
        for t in late:
            if self.world > 1:
                torch.distributed.all_reduce(t, group=self.pg)
        main.wait_stream(self.comm_stream)

    def _update(self):
        return self.opt.clip_and_step(self.max_grad_norm, grad_scale=1.0 if self.global_exact else 1.0 / self.world)

    def _body(self, st, L):
        if self._split():
            loss, y0, y0_leaf = self._part1(st, L)
            self._exchange_overlapped(lambda: self._part2(y0, y0_leaf))
            return loss, self._update()
        loss = self._fwd_bwd(st, L)
        if self.world > 1:
            torch.distributed.all_reduce(self.opt.flat_grad, group=self.pg)
        return loss, self._update()

    # ---- geometry -> arena views
    def _dev(self):
        return next(self.model.parameters()).device

    def _arena(self):
        if self.arena is None:
            self.arena = _Arena(self._dev())
        return self.arena

    def reserve(self, B, T, D, L):
        """Size every staging buffer for the largest batch the run can produce (e.g. config batch_size,
        max_feature_length, input_dim, max_text_length + 1): no buffer then ever moves and no graph is dropped."""
        _Views(self._arena(), B, T, D, L)
        for k in range(2):
            _Views(self._arena(), B, T, D, L, prefix=f"slot{k}.")

    def _global_extents(self, T, L, B):
        """(global Tmax, global Lmax + 1, global B) of this step's data-parallel batch."""
        if self.global_shape is not None:
            return self.global_shape
        if self.world == 1:
            return T, L, B
        t = torch.tensor([T, L, B], device=self._dev(), dtype=torch.int64)
        mx, sm = t.clone(), t.clone()
        torch.distributed.all_reduce(mx, op=torch.distributed.ReduceOp.MAX, group=self.pg)
        torch.distributed.all_reduce(sm, op=torch.distributed.ReduceOp.SUM, group=self.pg)
        return int(mx[0]), int(mx[1]), int(sm[2])

    def _geometry(self, xs, ilens, ys):
        """-> (key, views, host_lens, T_local, ys_in, ys_out); key = (B, T, D, L) is the PADDED geometry the step runs
        at (the shard's own extents, or the global ones in global-exact mode)."""
        m = self.model
        host_lens = _host_lens(ilens)
        Tl = max(host_lens)
        T, L_min, inv = Tl, 0, None
        if self.global_exact:
            L_loc = max(len(y) for y in ys) + 1
            T, L_min, Bg = self._global_extents(Tl, L_loc, len(host_lens))
            inv = 1.0 / float(Bg * L_min)
        ys_in, ys_out = build_targets(ys, m.decoder.bos, m.decoder.eos, m.decoder.pad, L_min)
        B, L = ys_out.shape
        key = (B, T, xs.shape[2], L)
        st = _Views(self._arena(), *key)
        if inv is not None:
            st.inv.fill_(inv)
        return key, st, host_lens, Tl, ys_in, ys_out

    @staticmethod
    def _copy_x(dst, xs, Tl):
        """Features into the [B, T, D] staging view; frames past the shard's own longest utterance are zero (global-exact
        padding)."""
        dst[:, :Tl].copy_(xs[:, :Tl], non_blocking=True)
        if Tl < dst.shape[1]:
            dst[:, Tl:].zero_()

    def staged(self, key):
        """The static device buffers (x, lens, ys_in, ys_out views) a step on geometry `key` reads."""
        return _Views(self._arena(), *key)

    def snapshot(self, key):
        """Device copies of the batch currently staged for `key` (the staging arenas are shared by all geometries, so a
        later stage() overwrites them); `restore` puts one back with device-to-device copies only."""
        st = self.staged(key)
        return tuple(t.clone() for t in (st.x, st.lens, st.ys_in, st.ys_out, st.inv))

    def restore(self, key, snap):
        st = self.staged(key)
        for dst, src in zip((st.x, st.lens, st.ys_in, st.ys_out, st.inv), snap):
            dst.copy_(src, non_blocking=True)
        return key

    def stage(self, xs, ilens, ys):
        """Host -> device copies of one batch into the static buffers (views of its geometry)."""
        key, st, host_lens, T, ys_in, ys_out = self._geometry(xs, ilens, ys)
        self._copy_x(st.x, xs, T)
        if self.input_noise_std > 0:
            st.x.add_(torch.randn_like(st.x), alpha=float(self.input_noise_std))
        self._staging_wait()
        st.h_lens.copy_(torch.tensor(host_lens, dtype=torch.int32))
        st.h_ys_in.copy_(torch.from_numpy(ys_in))
        st.h_ys_out.copy_(torch.from_numpy(ys_out))
        st.lens.copy_(st.h_lens, non_blocking=True)
        st.ys_in.copy_(st.h_ys_in, non_blocking=True)
        st.ys_out.copy_(st.h_ys_out, non_blocking=True)
        self._stage_ev = torch.cuda.Event()
        self._stage_ev.record(torch.cuda.current_stream(st.x.device))
        return key

    def _staging_wait(self):
        """The pinned staging of stage() is shared by all geometries: the previous batch's H2D copies must have
        left the host before it is overwritten."""
        if self._stage_ev is not None:
            self._stage_ev.synchronize()

    def run(self, key):
        """One train step on the batch currently staged for `key`. Returns (loss, grad_norm) device tensors.

        Single GPU: the whole step is one CUDA graph. Data parallel: forward+backward is one graph and the
        optimiser another, with the NCCL all-reduce of the flat gradient issued eagerly between them (a
        collective recorded inside a capture is not executed at capture time, which would desynchronise ranks
        that meet a new batch geometry at different steps)."""
        B, T, D, L = key
        st = _Views(self._arena(), B, T, D, L)
        self.model.train()
        if not self.use_graph:
            return self._body(st, L)
        stamp = (self.arena.gen, _hyper_stamp(self.opt), self.max_grad_norm, self.global_exact, self._split())
        ent = self.cache.lookup(key, stamp)
        if ent is None:
            if self.cache.sighting(key) == 0:      # first sight of a geometry: eager (also warms lazy init)
                return self._body(st, L)
            g = torch.cuda.CUDAGraph()
            Fn.warm_deferred(st.x.device)       # side stream + its workspace exist before the capture
            pool = self.cache.pool_handle()
            torch.cuda.synchronize()
            if self.cap_stream is None:
                # the critical path is captured on a high-priority stream: when both are pending, its CTAs are
                # scheduled before those of the (default-priority) weight-gradient side stream
                self.cap_stream = torch.cuda.Stream(device=st.x.device, priority=-1)
            ent = {"stamp": stamp}
            if self._split():
                with torch.cuda.graph(g, pool=pool, stream=self.cap_stream):
                    ent["loss"], y0, y0_leaf = self._part1(st, L)
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, pool=pool, stream=self.cap_stream):
                    self._part2(y0, y0_leaf)
                ent["graph2"] = g2
                ent["keep"] = (y0, y0_leaf)
            else:
                with torch.cuda.graph(g, pool=pool, stream=self.cap_stream):
                    if self.world == 1:
                        ent["loss"], ent["norm"] = self._body(st, L)
                    else:
                        ent["loss"] = self._fwd_bwd(st, L)
            ent["graph"] = g
            self.cache.store(key, ent)
        two_part = "graph2" in ent
        if (self.world > 1 or two_part) and (self.update_graph is None or self.update_stamp != stamp[1:]):
            torch.cuda.synchronize()
            gu = torch.cuda.CUDAGraph()
            step0 = self.opt.step_dev.clone()
            with torch.cuda.graph(gu):
                self.update_norm = self._update()
            self.update_graph, self.update_stamp = gu, stamp[1:]
            self.opt.step_dev.copy_(step0)     # nothing ran during capture, but keep the counter explicit
        ent["graph"].replay()
        if self.world == 1 and not two_part:
            return ent["loss"], ent["norm"]
        if two_part:
            self._exchange_overlapped(ent["graph2"].replay)
        else:
            torch.distributed.all_reduce(self.opt.flat_grad, group=self.pg)
        self.update_graph.replay()
        return ent["loss"], self.update_norm

    def step(self, xs, ilens, ys):
        return self.run(self.stage(xs, ilens, ys))

    # ---- pipelined epoch: the host->device copy of batch i+1 overlaps the train step of batch i
    def upload(self, xs, ilens, ys, slot):
        """Start the H2D copies of one batch into upload slot `slot` (0/1) on the copy stream. Returns a handle."""
        key, st, host_lens, T, ys_in, ys_out = self._geometry(xs, ilens, ys)
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(device=st.x.device)
        if self.slots is None:
            self.slots = [_Slot(), _Slot()]
        sl = self.slots[slot]
        if sl.used:
            sl.ready.synchronize()                     # the previous upload from these pinned buffers has left the host
        sl.views = v = _Views(self.arena, *key, prefix=f"slot{slot}.")
        sl.used = True
        v.h_lens.copy_(torch.tensor(host_lens, dtype=torch.int32))
        v.h_ys_in.copy_(torch.from_numpy(ys_in))
        v.h_ys_out.copy_(torch.from_numpy(ys_out))
        cs = self.copy_stream
        if sl.consumed is not None:
            cs.wait_event(sl.consumed)                 # the step that used this slot has copied it out
        with torch.cuda.stream(cs):
            self._copy_x(v.x, xs, T)
            v.lens.copy_(v.h_lens, non_blocking=True)
            v.ys_in.copy_(v.h_ys_in, non_blocking=True)
            v.ys_out.copy_(v.h_ys_out, non_blocking=True)
            sl.ready.record(cs)
        return key, slot

    def run_uploaded(self, handle):
        key, slot = handle
        st = _Views(self.arena, *key)
        sl = self.slots[slot]
        v = _Views(self.arena, *key, prefix=f"slot{slot}.")     # (the arena may have grown since upload(): re-view)
        main = torch.cuda.current_stream(st.x.device)
        main.wait_event(sl.ready)
        st.x.copy_(v.x, non_blocking=True)             # device-to-device: ~10 us for 32 MB
        if self.input_noise_std > 0:
            st.x.add_(torch.randn_like(st.x), alpha=float(self.input_noise_std))
        st.lens.copy_(v.lens, non_blocking=True)
        st.ys_in.copy_(v.ys_in, non_blocking=True)
        st.ys_out.copy_(v.ys_out, non_blocking=True)
        sl.consumed = torch.cuda.Event()
        sl.consumed.record(main)
        return self.run(key)

    def _readback(self, out, k):
        """Snapshot (loss, grad_norm) of the step just enqueued and start their D2H copy on the copy stream.
        Returns (host_pair [2] pinned f32, event)."""
        loss, norm = out
        dev = loss.device
        if self.rb_dev is None:
            self.rb_dev = torch.zeros(4, 2, device=dev, dtype=torch.float32)
            self.rb_host = torch.zeros(4, 2, dtype=torch.float32).pin_memory()
        main = torch.cuda.current_stream(dev)
        self.rb_dev[k, 0].copy_(loss.reshape(()), non_blocking=True)   # the static outputs are overwritten by the next replay
        self.rb_dev[k, 1].copy_(torch.as_tensor(norm, device=dev).reshape(()).float(), non_blocking=True)
        done = torch.cuda.Event()
        done.record(main)
        if self.rb_stream is None:     # NOT the upload stream: this one waits for the step, the uploads must not
            self.rb_stream = torch.cuda.Stream(device=dev)
        cs = self.rb_stream
        cs.wait_event(done)
        with torch.cuda.stream(cs):
            self.rb_host[k].copy_(self.rb_dev[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        return self.rb_host[k], ev

    def steps(self, batches, lag=2):
        """Train on an iterable of host batches (xs [B, T, D] ideally pinned, ilens, ys); yields (loss, grad_norm)
        per batch as HOST scalars (0-dim CPU tensors), in order. Two overlaps: the host->device copy of batch i+1
        runs during step i, and the results of step i are read back on the copy stream and handed out only after
        step i+`lag` has been enqueued, so the GPU never waits for the host between steps (lag 2 also rides out
        host stalls of up to one step time: driver-lock contention, garbage collection)."""
        import collections
        it = iter(batches)
        nxt = next(it, None)
        k, n = 0, 0
        handle = self.upload(*nxt, k) if nxt is not None else None
        pending = collections.deque()
        while handle is not None:
            out = self.run_uploaded(handle)            # asynchronous: returns as soon as the step is enqueued
            pending.append(self._readback(out, n % 4))
            n += 1
            nxt = next(it, None)
            k ^= 1
            handle = self.upload(*nxt, k) if nxt is not None else None
            if len(pending) > lag:
                host, ev = pending.popleft()
                ev.synchronize()
                yield host[0].clone(), host[1].clone()
        while pending:
            host, ev = pending.popleft()
            ev.synchronize()
            yield host[0].clone(), host[1].clone()


def _clip_and_step(opt, params, max_grad_norm):
    """(data-parallel gradient all-reduce +) clip_grad_norm_ + optimizer.step, for the fused optimiser or a stock
    torch optimiser."""
    world = 1
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        world = torch.distributed.get_world_size()
    if hasattr(opt, "clip_and_step"):
        if world > 1:
            torch.distributed.all_reduce(opt.flat_grad)
        return opt.clip_and_step(max_grad_norm, grad_scale=1.0 / world)
    if world > 1:
        for p in params:
            if p.grad is not None:
                torch.distributed.all_reduce(p.grad)
                p.grad.div_(world)
    norm = torch.nn.utils.clip_grad_norm_(params, max_norm=max_grad_norm)
    opt.step()
    Fn.advance_dropout_seed()
    return norm


class SSLTrainer:
    """`Solver.gen_train_one_iteration` (solver.py:460-495): free-running decode of the unpaired speech with the
    smooth embedding, LM ("judge") probabilities of the decoded tokens as per-token weights, paired supervised
    loss, loss = sup + unsup_weight * unsup, backward, clip, generator step.

    The judge's output enters the loss only as a weight (its input is the discrete prediction), so
    it is evaluated under no_grad here; the reference leaves it attached, which only fills the judge's unused
    `.grad` fields (dis_opt never steps in this phase)."""

    def __init__(self, model, judge, optimizer, max_grad_norm=5.0, unsup_weight=0.001, proportion=0.125,
                 smooth=True, scaling=3.0, guard_empty_mask=False, use_graph=False, max_graphs=4):
        self.guard_empty_mask = guard_empty_mask
        self.model, self.judge, self.opt = model, judge, optimizer
        self.max_grad_norm, self.unsup_weight, self.proportion = max_grad_norm, unsup_weight, proportion
        self.smooth, self.scaling = smooth, scaling
        # graph mode (fused optimiser only): the whole generator step -- both encoder/decoder passes, the judge, the
        # backward with its side-stream weight gradients, clip + AMSGrad -- is captured once per batch geometry
        self.use_graph = use_graph and hasattr(optimizer, "clip_and_step") and os.environ.get("LAS_NO_GRAPH", "0") != "1"
        self.arena = None
        self.cache = _GraphCache(max_graphs)
        self.cap_stream = None
        self.pair_stream = None
        self.overlap_passes = os.environ.get("LAS_SSL_OVERLAP", "1") != "0"
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size()

    def losses(self, lab, unlab):
        m = self.model
        (xs, ilens, ys), (uxs, uilens) = lab, unlab
        Lu = int(uxs.size(1) * self.proportion)                                    # solver.py:469
        if os.environ.get("LAS_DEBUG_SSL"):
            torch.cuda.synchronize()
            print(f"[ssl] lab x {tuple(xs.shape)} ilens {list(ilens)} ylens {[len(y) for y in ys]} | unlab x {tuple(uxs.shape)} "
                  f"uilens {list(uilens)} Lu {Lu}", flush=True)
        # the paired pass on its own stream, next to the free-running loop of the unpaired pass (see _body_dev): autograd
        # runs its backward nodes on that stream too and joins them into the caller's stream when backward() returns
        dev = next(m.parameters()).device
        ps = self._pair_stream(dev) if self.overlap_passes else None

        def paired():
            _, logp, _, _ = m(xs, ilens, ys)
            return -torch.mean(logp)                                                # solver.py:482

        if ps is not None:
            main = torch.cuda.current_stream(dev)
            ps.wait_stream(main)
            with torch.cuda.stream(ps):
                sup = paired()
        _, u_logp, u_pred, _ = m(uxs, uilens, ys=None, sample=False, label_smoothing=False, max_dec_timesteps=Lu,
                                 smooth=self.smooth, scaling=self.scaling)
        dbg = os.environ.get("LAS_DEBUG_SSL")
        if dbg:
            torch.cuda.synchronize()
            print(f"[ssl] free-run ok: pred range {int(u_pred.min())}..{int(u_pred.max())} shape {tuple(u_pred.shape)} "
                  f"logp finite {bool(torch.isfinite(u_logp).all())}", flush=True)
        with torch.no_grad():
            _, lm_probs, _ = self.judge(ys=u_pred, discrete_input=False)           # solver.py:473
        if dbg:
            torch.cuda.synchronize()
            print(f"[ssl] judge ok: finite {bool(torch.isfinite(lm_probs).all())}", flush=True)
        mask = (u_pred != m.decoder.eos).float()                                    # solver.py:477
        denom = torch.sum(mask)
        if self.guard_empty_mask:
            # solver.py:478 divides by sum(mask); when every free-run token is <EOS> that is 0/0 = NaN, which the
            # reference then back-propagates into every generator weight. Here the term is 0 in that case (the only
            # inputs on which the two differ are those where the reference's step is NaN).
            denom = denom.clamp_min(1.0)
        unsup = -torch.sum(lm_probs * u_logp * mask) / denom
        if ps is not None:
            main.wait_stream(ps)
        else:
            sup = paired()
        return sup + self.unsup_weight * unsup, sup, unsup, (u_logp, u_pred, lm_probs)

    # ---- device-resident body (no host work; CUDA-graph capturable)
    def _body_dev(self, st, L, Lu):
        """The paired pass runs on a stream of its own (`overlap_passes`, default on; LAS_SSL_OVERLAP=0 turns it off):
        its forward is forked at the beginning of the step, and autograd runs every backward node on the stream of its
        forward, so the paired pass's cluster-persistent kernels (64-80 SMs, latency-bound) execute next to the
        unpaired pass's free-running decoder (~2 500 small per-timestep kernels) instead of in front of it. The two
        passes share nothing but the weights (read-only until the optimiser) and the gradient buffers, which are only
        written on the single weight-gradient stream (functional.wgrad_scope) or by autograd's own accumulation."""
        m = self.model
        enc, dec = m.encoder.enc2, m.decoder
        dev = st.x.device
        jobs = enc.prep_jobs() + dec.prep_jobs(2 if self.smooth else 1) + dec.prep_jobs(0)
        jobs.sort(key=lambda j: "bwd" in j[0])
        Fn.prepare_ahead(jobs)
        main = torch.cuda.current_stream(dev)
        ps = self._pair_stream(dev) if self.overlap_passes else None

        def paired():
            enc_h = enc.forward_dev(st.x, st.lens)
            _, logp, _, _ = dec.forward_dev(enc_h, enc.out_lens_dev(st.lens), st.ys_in, st.ys_out, L, 0)
            return -torch.mean(logp)

        u_enc = enc.forward_dev(st.ux, st.ulens)
        if ps is not None:
            # forked AFTER the unpaired encoder: two BLSTM launches side by side need 160 of 148 SMs and each took
            # twice as long (3.7 ms for the unpaired encoder instead of 1.9); the unpaired chain is the critical path
            ps.wait_stream(main)
            with torch.cuda.stream(ps):
                sup = paired()
        _, u_logp, u_pred, _ = dec.forward_dev(u_enc, enc.out_lens_dev(st.ulens), None, None, Lu, 2 if self.smooth else 1,
                                               self.scaling, False)
        with torch.no_grad():
            _, lm_probs, _ = self.judge(ys=u_pred, discrete_input=False)
        mask = (u_pred != dec.eos).float()
        denom = torch.sum(mask)
        if self.guard_empty_mask:
            denom = denom.clamp_min(1.0)
        unsup = -torch.sum(lm_probs * u_logp * mask) / denom
        if ps is not None:
            main.wait_stream(ps)
        else:
            sup = paired()
        loss = sup + self.unsup_weight * unsup
        self.opt.zero_grad()
        with Fn.deferred_wgrad():
            loss.backward()
        if ps is not None:
            main.wait_stream(ps)
        if self.world > 1:
            return loss.detach(), sup.detach(), unsup.detach(), None
        norm = self.opt.clip_and_step(self.max_grad_norm, grad_scale=1.0)
        return loss.detach(), sup.detach(), unsup.detach(), norm

    def _pair_stream(self, dev):
        if self.pair_stream is None:
            self.pair_stream = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(self.pair_stream):
                Fn._gemm_workspace(dev)           # allocated outside any capture
        return self.pair_stream

    def stage(self, lab, unlab):
        """Host -> device copies of one (paired, unpaired) batch pair into the static buffers (arena views)."""
        (xs, ilens, ys), (uxs, uilens) = lab, unlab
        m = self.model
        dev = next(m.parameters()).device
        host_lens, uhost = _host_lens(ilens), _host_lens(uilens)
        T, Tu = max(host_lens), max(uhost)
        ys_in, ys_out = build_targets(ys, m.decoder.bos, m.decoder.eos, m.decoder.pad)
        B, L = ys_out.shape
        Lu = int(uxs.size(1) * self.proportion)                                    # solver.py:469 (padded extent as given)
        key = (B, T, xs.shape[2], L, len(uhost), Tu, Lu)
        if self.arena is None:
            self.arena = _Arena(dev)
        st = self._views(key)
        st.x.copy_(xs[:, :T], non_blocking=True)
        st.ux.copy_(uxs[:, :Tu], non_blocking=True)
        st.lens.copy_(torch.tensor(host_lens, dtype=torch.int32), non_blocking=True)
        st.ulens.copy_(torch.tensor(uhost, dtype=torch.int32), non_blocking=True)
        st.ys_in.copy_(torch.from_numpy(ys_in), non_blocking=True)
        st.ys_out.copy_(torch.from_numpy(ys_out), non_blocking=True)
        return key

    def _views(self, key):
        B, T, D, L, Bu, Tu, Lu = key
        st = _Views(self.arena, B, T, D, L)
        st.ux = self.arena.get("ux", (Bu, Tu, D), torch.float32)
        st.ulens = self.arena.get("ulens", (Bu,), torch.int32)
        return st

    def run(self, key):
        st = self._views(key)
        L, Lu = key[3], key[6]
        self.model.train()
        self.judge.train()
        stamp = (self.arena.gen, _hyper_stamp(self.opt), self.max_grad_norm, self.unsup_weight, self.scaling)
        ent = self.cache.lookup(key, stamp)
        if ent is None:
            if self.cache.sighting(key) == 0:      # first sight of a geometry: eager (also warms lazy init)
                return self._finish(self._body_dev(st, L, Lu))
            g = torch.cuda.CUDAGraph()
            Fn.warm_deferred(st.x.device)
            pool = self.cache.pool_handle()
            torch.cuda.synchronize()
            if self.cap_stream is None:
                self.cap_stream = torch.cuda.Stream(device=st.x.device, priority=-1)
            ent = {"stamp": stamp}
            with torch.cuda.graph(g, pool=pool, stream=self.cap_stream):
                ent["out"] = self._body_dev(st, L, Lu)
            ent["graph"] = g
            self.cache.store(key, ent)
        ent["graph"].replay()
        return self._finish(ent["out"])

    def _finish(self, out):
        loss, sup, unsup, norm = out
        if self.world > 1:                         # data parallel: all-reduce + update outside the captured part
            torch.distributed.all_reduce(self.opt.flat_grad)
            norm = self.opt.clip_and_step(self.max_grad_norm, grad_scale=1.0 / self.world)
        return loss, sup, unsup, norm

    def step(self, lab, unlab):
        if self.use_graph:
            return self.run(self.stage(lab, unlab))
        self.model.train()
        self.judge.train()
        loss, sup, unsup, _ = self.losses(lab, unlab)
        self.opt.zero_grad()
        with Fn.deferred_wgrad():
            loss.backward()
        if self.pair_stream is not None:
            torch.cuda.current_stream(self.pair_stream.device).wait_stream(self.pair_stream)
        norm = _clip_and_step(self.opt, list(self.model.parameters()), self.max_grad_norm)
        return loss.detach(), sup.detach(), unsup.detach(), norm


class JudgeTrainer:
    """`Solver.judge_train_one_iteration` (solver.py:288-301): masked LM loss over len+5 positions, backward,
    clip, plain Adam step. With the fused optimiser on one GPU the whole step is captured as one CUDA graph per
    (B, Lmax+5) geometry (a few hundred timestep launches otherwise issued from Python)."""

    def __init__(self, judge, optimizer, max_grad_norm=5.0, use_graph=True, max_graphs=32):
        self.judge, self.opt, self.max_grad_norm = judge, optimizer, max_grad_norm
        world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size()
        self.use_graph = (use_graph and hasattr(optimizer, "clip_and_step") and world == 1
                          and os.environ.get("LAS_NO_GRAPH", "0") != "1")
        self.arena = None
        self.cache = _GraphCache(max_graphs)

    def losses(self, ys):
        log_probs, probs, _ = self.judge(ys)
        loss = -self.judge.mask_and_cal_sum(log_probs, ys)
        avg_prob = self.judge.mask_and_cal_sum(probs, ys)
        return loss, avg_prob

    # ---- device-resident step (graph mode)
    def _views(self, B, Lm):
        g = self.arena.get
        st = type("JudgeViews", (), {})()
        st.ys_in, st.ys_out = g("ys_in", (B, Lm), torch.int64), g("ys_out", (B, Lm), torch.int64)
        st.lens, st.mask, st.inv = g("lens", (B,), torch.int32), g("mask", (B, Lm), torch.float32), g("inv", (1,), torch.float32)
        return st

    def stage(self, ys):
        j = self.judge
        dev = j.embedding.weight.device
        ys_in, ys_out, lens = j.targets(ys)
        B, Lm = ys_in.shape
        if self.arena is None:
            self.arena = _Arena(dev)
        st = self._views(B, Lm)
        mask = (np.arange(Lm)[None, :] < np.asarray(lens)[:, None]).astype(np.float32)   # utils._seq_mask (model.py:565-573)
        st.ys_in.copy_(torch.from_numpy(ys_in), non_blocking=True)
        st.ys_out.copy_(torch.from_numpy(ys_out), non_blocking=True)
        st.lens.copy_(torch.tensor(lens, dtype=torch.int32), non_blocking=True)
        st.mask.copy_(torch.from_numpy(mask), non_blocking=True)
        st.inv.copy_(torch.tensor([1.0 / float(sum(lens))]), non_blocking=True)
        return B, Lm

    def _body(self, st, Lm):
        log_probs, probs, _ = self.judge.forward_dev(st.ys_in, st.ys_out, st.lens, Lm)
        loss = -torch.sum(log_probs * st.mask) * st.inv[0]
        avg_prob = torch.sum(probs * st.mask) * st.inv[0]
        self.opt.zero_grad()
        with Fn.deferred_wgrad():
            loss.backward()
        norm = self.opt.clip_and_step(self.max_grad_norm, grad_scale=1.0)
        return loss.detach(), avg_prob.detach(), norm

    def run(self, key):
        B, Lm = key
        st = self._views(B, Lm)
        stamp = (self.arena.gen, _hyper_stamp(self.opt), self.max_grad_norm)
        ent = self.cache.lookup(key, stamp)
        if ent is None:
            if self.cache.sighting(key) == 0:
                return self._body(st, Lm)
            g = torch.cuda.CUDAGraph()
            Fn.warm_deferred(st.mask.device)
            pool = self.cache.pool_handle()
            torch.cuda.synchronize()
            ent = {"stamp": stamp}
            with torch.cuda.graph(g, pool=pool):
                ent["out"] = self._body(st, Lm)
            ent["graph"] = g
            self.cache.store(key, ent)
        ent["graph"].replay()
        return ent["out"]

    def step(self, ys):
        self.judge.train()
        if self.use_graph:
            return self.run(self.stage(ys))
        loss, avg_prob = self.losses(ys)
        self.opt.zero_grad()
        with Fn.deferred_wgrad():
            loss.backward()
        norm = _clip_and_step(self.opt, list(self.judge.parameters()), self.max_grad_norm)
        return loss.detach(), avg_prob.detach(), norm
