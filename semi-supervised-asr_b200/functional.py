"""torch.autograd.Function wrappers around the C-ABI kernels (include/las_b200.h).

Everything here runs on the CUDA stream torch considers current; tensors are allocated by torch,
raw device pointers go through ctypes. There is no eager-PyTorch or CPU fallback: without
liblas_b200.so (or without a CUDA device) the calls raise.

Precision: bf16 MMA operands, f32 accumulation, f32 recurrent cell state, f32 master weights and
gradients.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import DecArgs, call, ptr  # noqa: F401  (re-exported for the modules)

BF16 = torch.bfloat16
# cluster-persistent attention decoder (csrc/decoder_persistent.cu); LAS_DEC_PERSIST=0 selects the per-step kernels
DEC_PERSISTENT = os.environ.get("LAS_DEC_PERSIST", "1") == "1"


# --------------------------------------------------------------------------------------------
# dropout: masks are a function of (device seed, site id, element index) -- see las_dropout. The seed advances
# once per optimiser step (FusedAdam / the trainers call advance_dropout_seed()); every Function.forward call
# inside a step draws fresh site ids, so two forwards of the same module in one step use different masks.
# --------------------------------------------------------------------------------------------
_DROP = {"seed": None, "calls": 0}


def dropout_seed(device):
    if _DROP["seed"] is None or _DROP["seed"].device != device:
        _DROP["seed"] = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFF], device=device, dtype=torch.int64)
    return _DROP["seed"]


def set_dropout_seed(value, device):
    """Restore the device seed (resume): written INTO the existing tensor, whose address captured graphs hold."""
    dropout_seed(device).fill_(int(value) & 0x7FFFFFFFFFFF)
    _DROP["calls"] = 0


def current_dropout_seed():
    """Host copy of the device seed (None before the first dropout call). Synchronises; checkpoint time only."""
    return None if _DROP["seed"] is None else int(_DROP["seed"].item())


def advance_dropout_seed():
    """New masks from the next forward on. Call once per training step (after backward)."""
    if _DROP["seed"] is not None:
        _DROP["seed"] += 1
    _DROP["calls"] = 0


def new_sites(n=64):
    """Base site id for one Function.forward call (n consecutive ids reserved)."""
    _DROP["calls"] += 1
    return (_DROP["calls"] * n) & 0x7FFFFFFF


def dropout_(x, B, T, W, ld_b, ld_t, rep, p, site):
    """In-place dropout on a strided [B, T(+rep), W] view of a bf16 or f32 tensor."""
    call("las_dropout", ptr(x), int(x.dtype == BF16), B, T, W, ld_b, ld_t, int(rep), float(p),
         ptr(dropout_seed(x.device)), int(site))


def _r8(n):
    return (n + 7) // 8 * 8


def _r16(n):
    return (n + 15) // 16 * 16


# --------------------------------------------------------------------------------------------
# thin op wrappers
# --------------------------------------------------------------------------------------------
_GEMM_WS = {}


def _gemm_workspace(device):
    """Split-K workspace (32 MiB of f32, one per device and stream; stream-ordered use only)."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _GEMM_WS.get(key)
    if ws is None:
        ws = _GEMM_WS[key] = torch.empty(8 << 20, device=device, dtype=torch.float32)
    return ws


# --------------------------------------------------------------------------------------------
# deferred weight gradients: the serial kernels of the backward pass (BPTT recurrences, decoder loop) occupy
# at most 64 of the 148 SMs, and every parameter gradient is a dense contraction that nothing on that
# critical path waits for. Inside `with deferred_wgrad():` the Functions below fork those contractions onto a
# side stream, accumulate them straight into `param.grad` there and hand autograd `None`; leaving the context
# joins the side stream. Only trainers whose optimiser pre-allocates `.grad` (optim.FusedAdam) use it; a bare
# `loss.backward()` keeps the ordinary single-stream behaviour.
# --------------------------------------------------------------------------------------------
_DEFER = {"on": False, "streams": {}, "keep": [], "used": False, "late": []}


class deferred_wgrad:
    def __enter__(self):
        _DEFER["on"] = os.environ.get("LAS_NO_DEFER", "0") != "1"     # debugging switch: single-stream backward
        return self

    def __exit__(self, *exc):
        _DEFER["on"] = False
        join_deferred()
        return False


def run_late_jobs():
    """Side-stream jobs that were parked until a long serial kernel occupies the main stream (DecoderFn.backward parks
    the energy-MLP parameter sums; EncoderFn.backward releases them next to layer 0's BPTT). Call on the side stream."""
    jobs, _DEFER["late"] = _DEFER["late"], []
    for _, job in jobs:
        job()


def join_deferred():
    """The current stream waits for every weight-gradient kernel forked so far."""
    if _DEFER["late"]:                    # nobody released them (no encoder backward followed): run them now
        dev = _DEFER["late"][0][0]
        side = warm_deferred(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        _DEFER["used"] = True
        with torch.cuda.stream(side):
            run_late_jobs()
    if _DEFER["used"]:
        for dev, side in _DEFER["streams"].items():
            torch.cuda.current_stream(dev).wait_stream(side)
        _DEFER["used"] = False
    _DEFER["keep"].clear()
    _PREP.clear()
    _PREP_PENDING["jobs"], _PREP_PENDING["start"] = [], None


def warm_deferred(device):
    """Create the side stream and its GEMM workspace outside any graph capture."""
    side = _DEFER["streams"].get(device)
    if side is None:
        side = _DEFER["streams"][device] = torch.cuda.Stream(device=device)
    with torch.cuda.stream(side):
        _gemm_workspace(device)
    return side


class wgrad_scope:
    """`with wgrad_scope(weights, keep...) as sc:` runs its body on the side stream when deferral is active for
    these weights; `sc.deliver(grads)` then either accumulates into `.grad` (returning Nones) or passes through."""

    def __init__(self, weights, *keep):
        self.weights = weights
        self.deferred = _DEFER["on"] and all(
            (not w.requires_grad) or (isinstance(w, torch.nn.Parameter) and w.grad is not None) for w in weights)
        self.keep = keep
        self.cm = None

    def __enter__(self):
        if self.deferred:
            dev = self.weights[0].device
            side = warm_deferred(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            _DEFER["keep"].append(self.keep)      # operands allocated on the main stream stay alive until the join
            _DEFER["used"] = True
            self.cm = torch.cuda.stream(side)
            self.cm.__enter__()
        return self

    def __exit__(self, *exc):
        if self.cm is not None:
            self.cm.__exit__(*exc)
        return False

    def deliver(self, grads):
        if not self.deferred:
            # autograd may adopt a returned tensor as `.grad` without copying: two parameters that share one
            # gradient (b_ih / b_hh) must not end up sharing its memory (clip_grad_norm_ would scale it twice,
            # a second backward would add into it twice)
            out, seen = [], set()
            for g in grads:
                if g is not None:
                    key = (g.data_ptr(), tuple(g.shape), tuple(g.stride()))
                    if key in seen:
                        g = g.clone()
                    seen.add(key)
                out.append(g)
            return out
        for w, g in zip(self.weights, grads):
            if g is not None and w.requires_grad:
                w.grad.add_(g.reshape(w.grad.shape) if g.shape != w.grad.shape else g)
        _DEFER["keep"].append(list(grads))
        return [None] * len(grads)


def zeros_many(dev, specs):
    """One zeroed allocation carved into tensors: specs = [(shape, dtype), ...] -> list of views (256-byte aligned).
    One fill instead of one tiny kernel per buffer on the critical path."""
    offs, total = [], 0
    for shape, dtype in specs:
        nbytes = int(torch.Size(shape).numel()) * torch.empty((), dtype=dtype).element_size()
        offs.append((total, nbytes))
        total += (nbytes + 255) // 256 * 256
    buf = torch.zeros(total, device=dev, dtype=torch.uint8)
    return [buf[o:o + nb].view(dtype).view(shape) for (o, nb), (shape, dtype) in zip(offs, specs)]


def gemm(A, lda, a_mn, Bm, ldb, b_mn, M, N, K, out=None, out_bf16=False, bias=None, relu=False,
         accumulate=False, ldc=None):
    """D[m,n] = sum_k A[m,k] B[n,k] (+bias[n]) (relu) (+=D). A/Bm: bf16 tensors (base pointers)."""
    if out is None:
        out = torch.empty(M, N, device=A.device, dtype=BF16 if out_bf16 else torch.float32)
    if ldc is None:
        ldc = N
    ws = _gemm_workspace(A.device)
    call("las_gemm_bf16_ws", ptr(A), lda, int(a_mn), ptr(Bm), ldb, int(b_mn), ptr(out), ldc, int(out.dtype == BF16),
         ptr(bias), M, N, K, int(relu), int(accumulate), ptr(ws), ws.numel() * 4)
    return out


def cvt_bf16(src, cols=None, ld_dst=None, extra_rows=0):
    """f32 [rows, cols] (row stride src.stride(0)) -> bf16 [rows(+extra zero rows), ld_dst], zero padded."""
    src2 = src.reshape(-1, src.shape[-1]) if src.dim() != 2 else src
    rows = src2.shape[0]
    cols = src2.shape[1] if cols is None else cols
    ld_dst = _r8(cols) if ld_dst is None else ld_dst
    if extra_rows:
        dst = torch.zeros(rows + extra_rows, ld_dst, device=src.device, dtype=BF16)
    else:
        dst = torch.empty(rows, ld_dst, device=src.device, dtype=BF16)
    call("las_cvt_pad_bf16", ptr(src2), src2.stride(0), rows, cols, ptr(dst), ld_dst)
    return dst


def pack_afrag(W, mode, H=0, transposed=False, col_offset=0, cols=None):
    """f32 weight matrix -> mma A fragments. Logical A is [rows, cols] (see las_pack_afrag)."""
    W = W.contiguous()
    if transposed:
        rows = W.shape[1] if cols is None else cols
        kcols = W.shape[0]
        rows_arg, cols_arg = rows, kcols
    else:
        rows_arg = W.shape[0]
        cols_arg = (W.shape[1] - col_offset) if cols is None else cols
    nbytes = _lib.lib().las_afrag_bytes(rows_arg, cols_arg, mode, H)
    out = torch.empty(nbytes // 4, device=W.device, dtype=torch.int32)
    call("las_pack_afrag", ptr(W), W.stride(0), rows_arg, cols_arg, col_offset, mode, H, int(transposed), ptr(out))
    return out


def pack_whhT(w_hh):
    """W_hh^T fragments for the BPTT kernels. Returns (packed, layout): layout 1 = owner-ordered for
    the cluster-persistent kernel when it serves this hidden size, 0 = plain transposed."""
    H = w_hh.shape[1]
    if _lib.lib().las_lstm_persistent_geometry(H, None, None):
        out = torch.empty(_lib.lib().las_whhT_owner_bytes(H) // 4, device=w_hh.device, dtype=torch.int32)
        call("las_pack_whhT_owner", ptr(w_hh.contiguous()), H, ptr(out))
        return out, 1
    return pack_afrag(w_hh, 0, transposed=True), 0


def colsum(x, cols, out=None, ld=None, rows=None):
    x_bf = x.dtype == BF16
    ld = x.stride(0) if ld is None else ld
    rows = x.shape[0] if rows is None else rows
    if out is None:
        out = torch.zeros(cols, device=x.device, dtype=torch.float32)
    call("las_colsum", ptr(x), int(x_bf), ld, rows, cols, ptr(out))
    return out


def lens_tensor(lens, device):
    if torch.is_tensor(lens):
        return lens.to(device=device, dtype=torch.int32)
    return torch.tensor([int(l) for l in lens], dtype=torch.int32, device=device)


# --------------------------------------------------------------------------------------------
# weight preparation ahead of use: the bf16 copies / MMA fragment packs / permutations of the parameters are pure
# functions of the weights (~110 small kernels per step). The Functions obtain them through `_prep(kind, tensors,
# builder)`; a trainer may run the same builders at the very beginning of the step on the side stream
# (`prepare_ahead(jobs)`), so that only layer 0's preparation stays on the critical path and the rest overlaps
# the first recurrence. Without `prepare_ahead` every builder simply runs inline.
# --------------------------------------------------------------------------------------------
_PREP = {}
_PREP_PENDING = {"jobs": [], "start": None}


def _prep_key(kind, tensors):
    return (kind,) + tuple(id(t) for t in tensors)


def _prep(kind, tensors, builder):
    key = _prep_key(kind, tensors)
    ent = _PREP.get(key)
    if ent is None:
        # not prepared ahead (or not yet: the side-stream fork is still parked): build inline, and take it off the list
        _PREP_PENDING["jobs"] = [j for j in _PREP_PENDING["jobs"] if _prep_key(j[0], j[1]) != key]
        return builder()
    val, ev = ent
    torch.cuda.current_stream(tensors[0].device).wait_event(ev)
    return val


def prepare_ahead(jobs):
    """jobs: iterable of (kind, tensors, builder) in the order the step needs them. The builders run on the side
    stream; the results are handed to the matching `_prep` calls of this step. `join_deferred()` (end of the
    backward pass) joins the stream and drops the cache.

    The fork is PARKED until the first recurrence of the step has been enqueued (`release_prepared()`, called by
    lstm_layer_fwd): the first layer's operands are built inline on the critical path, and everything else is captured
    AFTER the step's first long kernel. Enqueued first, the ~60 small preparation kernels delayed the start of the
    critical path by 0.3 ms in the captured graph (the executor issues nodes in capture order)."""
    jobs = list(jobs)
    if not jobs or os.environ.get("LAS_NO_PREP", "0") == "1":        # debugging switch: build everything inline
        return
    dev = jobs[0][1][0].device
    warm_deferred(dev)
    start = torch.cuda.Event()
    start.record(torch.cuda.current_stream(dev))                       # the side stream depends on the step's beginning only
    _PREP_PENDING["jobs"], _PREP_PENDING["start"] = jobs, start
    if os.environ.get("LAS_PREP_LATE", "1") != "1":                   # debugging switch: fork at the very beginning
        release_prepared()


def release_prepared():
    """Run the parked preparation jobs on the side stream now (no-op when there are none)."""
    jobs, start = _PREP_PENDING["jobs"], _PREP_PENDING["start"]
    _PREP_PENDING["jobs"], _PREP_PENDING["start"] = [], None
    if not jobs:
        return
    dev = jobs[0][1][0].device
    side = warm_deferred(dev)
    side.wait_event(start)
    _DEFER["used"] = True
    with torch.cuda.stream(side):
        for kind, tensors, builder in jobs:
            val = builder()
            ev = torch.cuda.Event()
            ev.record(side)
            _PREP[_prep_key(kind, tensors)] = (val, ev)


def _build_lstm_fwd(w_ih, w_hh, b_ih, b_hh, Dp):
    """-> (persist, wcat_bf [ndir*4H, Dp] bf16, bcat [ndir*4H] f32, whh_pk) for one LSTM layer."""
    dev = w_hh[0].device
    ndir = len(w_hh)
    H = w_hh[0].shape[1]
    persist = bool(_lib.lib().las_lstm_persistent_geometry(H, None, None))
    wcat = torch.cat(w_ih, dim=0) if ndir > 1 else w_ih[0]
    bias = [bi + bh for bi, bh in zip(b_ih, b_hh)]
    bcat = torch.cat(bias) if ndir > 1 else bias[0]
    if persist:
        perm, _ = gate_perm(H, ndir, dev)
        wcat, bcat = wcat.index_select(0, perm), bcat.index_select(0, perm)
    wcat_bf = cvt_bf16(wcat, ld_dst=Dp)
    whh_pk = torch.cat([pack_afrag(w, 3 if persist else 1, H) for w in w_hh])
    return persist, wcat_bf, bcat, whh_pk


def _build_lstm_bwd(w_hh):
    H = w_hh[0].shape[1]
    if _lib.lib().las_lstm_persistent_geometry(H, None, None):
        return torch.cat([pack_whhT(w)[0] for w in w_hh])
    return torch.cat([pack_afrag(w, 0, transposed=True) for w in w_hh])


def lstm_prep_jobs(w_ih, w_hh, b_ih, b_hh, Dp):
    """The prepare_ahead jobs of one LSTM layer (forward and backward packs)."""
    w_ih, w_hh, b_ih, b_hh = list(w_ih), list(w_hh), list(b_ih), list(b_hh)
    return [("lstm_fwd", w_ih + w_hh + b_ih + b_hh, lambda: _build_lstm_fwd(w_ih, w_hh, b_ih, b_hh, Dp)),
            ("lstm_bwd", w_hh, lambda: _build_lstm_bwd(w_hh))]


# --------------------------------------------------------------------------------------------
# one LSTM layer over padded sequences (shared by the listener and the LM)
# --------------------------------------------------------------------------------------------
_GATE_PERM = {}


def gate_perm(H, ndir, device):
    """Row permutation between torch's gate-major order [dir][gate][unit] and the gate-minor order
    [dir][unit][gate] of the cluster-persistent kernels. Returns (perm, inv): X_minor = X_major[perm],
    X_major = X_minor[inv]."""
    key = (H, ndir, device)
    pi = _GATE_PERM.get(key)
    if pi is None:
        perm = torch.arange(ndir * 4 * H, device=device).view(ndir, 4, H).permute(0, 2, 1).reshape(-1).contiguous()
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(ndir * 4 * H, device=device)
        pi = _GATE_PERM[key] = (perm, inv)
    return pi


def lstm_layer_fwd(xin, Dp, w_ih, w_hh, b_ih, b_hh, lens, B, T, Tp, rep):
    """xin bf16 [B*T, Dp]; w_ih / w_hh / b_ih / b_hh: per-direction lists. Returns (y bf16 [B, Tp, ndir*H], zero
    past each length, `rep` replicated row written; saved state for lstm_layer_bwd)."""
    dev = xin.device
    ndir = len(w_hh)
    H = w_hh[0].shape[1]
    w_ih, w_hh, b_ih, b_hh = list(w_ih), list(w_hh), list(b_ih), list(b_hh)
    persist, wcat_bf, bcat, whh_pk = _prep("lstm_fwd", w_ih + w_hh + b_ih + b_hh,
                                           lambda: _build_lstm_fwd(w_ih, w_hh, b_ih, b_hh, Dp))
    assert wcat_bf.shape[1] == Dp
    xproj = gemm(xin, Dp, 0, wcat_bf, Dp, 0, B * T, ndir * 4 * H, Dp, bias=bcat)   # f32 [B*T, ndir*4H]
    hprev = torch.empty(B, T, ndir * H, device=dev, dtype=BF16)
    # the persistent kernel writes every row of y (zeros past each length); the per-timestep kernels write active rows only
    y = (torch.empty if persist else torch.zeros)(B, Tp, ndir * H, device=dev, dtype=BF16)
    if persist:
        rec = torch.empty(ndir * B * T * H, 4, device=dev, dtype=torch.int32)
        call("las_lstm_persist_fwd", ptr(xproj), ptr(whh_pk), ptr(lens), B, T, H, ndir, ptr(y), Tp * ndir * H, ndir * H,
             rep, ptr(hprev), T * ndir * H, ndir * H, ptr(rec))
        release_prepared()        # parked weight preparation: forked now, underneath this recurrence
        act = (rec,)
    else:
        gates = torch.empty(ndir, B, T, H, 4, device=dev, dtype=torch.float16)
        csave = torch.empty(ndir, B, T, H, device=dev, dtype=torch.float32)
        ws = torch.empty(_lib.lib().las_lstm_ws_bytes(B, H, ndir), device=dev, dtype=torch.uint8)
        call("las_lstm_seq_fwd", ptr(xproj), ptr(whh_pk), ptr(lens), B, T, H, ndir, ptr(y), Tp * ndir * H, ndir * H,
             rep, ptr(hprev), T * ndir * H, ndir * H, ptr(gates), ptr(csave), ptr(ws))
        release_prepared()
        act = (gates, csave)
    return y, (xin, wcat_bf, hprev, act, persist, lens, B, T, Tp, rep, Dp, H, ndir)


def lstm_layer_bwd(saved, w_hh, dy, need_dx=True):
    """BPTT of one layer (the critical path). dy f32 [B, Tp, ndir*H]. Returns (dx f32 [B*T, Dp] or None, dG bf16
    [B*T, ndir*4H] for lstm_layer_wgrad)."""
    xin, wcat_bf, hprev, act, persist, lens, B, T, Tp, rep, Dp, H, ndir = saved
    dev = dy.device
    G = ndir * 4 * H
    dG = torch.empty(B * T, G, device=dev, dtype=BF16)
    w_hh = list(w_hh)
    whhT = _prep("lstm_bwd", w_hh, lambda: _build_lstm_bwd(w_hh))
    if persist:
        call("las_lstm_persist_bwd", ptr(dy), Tp * ndir * H, ndir * H, rep, ptr(whhT), ptr(lens), B, T, H, ndir,
             ptr(act[0]), ptr(dG), T * G, G)
    else:
        ws = torch.empty(ndir * B * H, device=dev, dtype=torch.float32)
        call("las_lstm_seq_bwd", ptr(dy), Tp * ndir * H, ndir * H, rep, ptr(whhT), 0, ptr(lens), B, T, H, ndir,
             ptr(act[0]), ptr(act[1]), ptr(dG), T * G, G, ptr(ws))
    dx = gemm(dG, G, 0, wcat_bf, Dp, 1, B * T, Dp, G) if need_dx else None           # f32 [B*T, Dp]
    return dx, dG


def lstm_layer_wgrad(saved, dG):
    """Weight gradients of one LSTM layer as dense contractions over all (b, t) (off the critical path: call it
    inside a wgrad_scope). Returns (d_w_ih [ndir*4H, Dp], d_w_hh list, d_bias [ndir*4H]) in torch's gate-major order."""
    xin, wcat_bf, hprev, act, persist, lens, B, T, Tp, rep, Dp, H, ndir = saved
    dev = dG.device
    G = ndir * 4 * H
    d_wcat = gemm(dG, G, 1, xin, Dp, 1, G, Dp, B * T)                                # [ndir*4H, Dp]
    d_bcat = colsum(dG, G)
    hp2 = hprev.view(B * T, ndir * H)
    d_whh = [gemm(dG[:, 4 * H * d:], G, 1, hp2[:, H * d:], ndir * H, 1, 4 * H, H, B * T) for d in range(ndir)]
    if persist:
        _, inv = gate_perm(H, ndir, dev)
        _, inv1 = gate_perm(H, 1, dev)
        d_wcat, d_bcat = d_wcat.index_select(0, inv), d_bcat.index_select(0, inv)
        d_whh = [g.index_select(0, inv1) for g in d_whh]
    return d_wcat, d_whh, d_bcat


# --------------------------------------------------------------------------------------------
# pyramidal BLSTM encoder (model.py:58-98)
# --------------------------------------------------------------------------------------------
class EncoderFn(torch.autograd.Function):
    """x f32 [B, T, D] (zeros past each length), lens int32 [B] (device) ->
    enc_h f32 [B, Te, H]. `weights` per layer: w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r,
    b_hh_r, proj_w, proj_b."""

    @staticmethod
    def forward(ctx, x, lens, subsample, p_drop, *weights):
        B, T, D = x.shape
        site0 = new_sites() if p_drop > 0 else 0
        dev = x.device
        n_layers = len(subsample)
        assert len(weights) == 10 * n_layers
        xin = cvt_bf16(x.reshape(B * T, D))          # [B*T, Dp] bf16, zero padded columns
        Dp = xin.shape[1]
        cur_lens = lens
        saved = []
        for i, sub in enumerate(subsample):
            w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r, proj_w, proj_b = weights[10 * i:10 * i + 10]
            H = w_hh.shape[1]
            Tp = T + (T % 2) if sub > 1 else T
            rep = int(sub > 1 and T % 2 == 1)
            y, lsaved = lstm_layer_fwd(xin, Dp, [w_ih, w_ih_r], [w_hh, w_hh_r], [b_ih, b_ih_r], [b_hh, b_hh_r], cur_lens,
                                       B, T, Tp, rep)
            if p_drop > 0:      # model.py:82 (before the replicate pad of an odd extent: the extra row shares its mask)
                dropout_(y, B, T, 2 * H, Tp * 2 * H, 2 * H, rep, p_drop, site0 + 2 * i)
            if sub > 1:
                T2, Kp = Tp // 2, 4 * H
            else:
                T2, Kp = T, 2 * H
            wp = _prep("cvt", [proj_w], lambda: cvt_bf16(proj_w))                   # [H, Kp]
            out = gemm(y, Kp, 0, wp, Kp, 0, B * T2, proj_w.shape[0], Kp, out_bf16=True, bias=proj_b, relu=True)
            if p_drop > 0:      # model.py:95 (all rows, incl. the relu(bias) rows past each length)
                dropout_(out, B * T2, 1, proj_w.shape[0], proj_w.shape[0], proj_w.shape[0], 0, p_drop, site0 + 2 * i + 1)
            saved.append((lsaved, y, out, wp, T, Tp, T2, Dp, H, rep))
            if sub > 1:
                nl = torch.empty_like(cur_lens)
                call("las_pyramid_lens", ptr(cur_lens), B, sub, ptr(nl))
                cur_lens = nl
            xin, T, Dp = out, T2, proj_w.shape[0]
        ctx.saved = saved
        ctx.subsample = list(subsample)
        ctx.drop = (float(p_drop), site0)
        ctx.B = B
        ctx.weights = weights
        ctx.D = D
        return out.view(B, T, Dp).float()

    @staticmethod
    def backward(ctx, denc):
        B = ctx.B
        n_layers = len(ctx.subsample)
        grads = [None] * (10 * n_layers)
        p_drop, site0 = ctx.drop
        dout = denc.contiguous().float()
        if p_drop > 0:
            dout = dout.clone()                       # masked in place below
        for i in reversed(range(n_layers)):
            lsaved, y, out, wp, T, Tp, T2, Dp, H, rep = ctx.saved[i]
            lw = ctx.weights[10 * i:10 * i + 10]
            w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r, proj_w, proj_b = lw
            dev = dout.device
            Ho, Kp = proj_w.shape
            n = B * T2
            if p_drop > 0:
                dropout_(dout, n, 1, Ho, Ho, Ho, 0, p_drop, site0 + 2 * i + 1)
            dz = torch.empty(n, Ho, device=dev, dtype=BF16)
            call("las_relu_bwd", ptr(dout), ptr(out), 1, ptr(dz), n * Ho)
            # critical path: dy = dz W_proj -> BPTT -> dx
            dy = gemm(dz, Ho, 0, wp, Kp, 1, n, Kp, Ho)                               # f32 [B, Tp, 2H] view
            if p_drop > 0:
                dropout_(dy, B, T, 2 * H, Tp * 2 * H, 2 * H, rep, p_drop, site0 + 2 * i)
            # projection gradients dW = dz^T yview, db = colsum(dz) need only dz: forked BEFORE this layer's BPTT
            with wgrad_scope(lw[8:10], dz, y) as sc:
                grads[10 * i + 8:10 * i + 10] = sc.deliver([gemm(dz, Ho, 1, y, Kp, 1, Ho, Kp, n), colsum(dz, Ho)])
                if i == 0 and sc.deferred:
                    run_late_jobs()           # parked side-stream work runs next to the longest BPTT kernel
            need_x = i > 0 or ctx.needs_input_grad[0]      # the input's gradient only when somebody asked (split encoder)
            dx, dG = lstm_layer_bwd(lsaved, [w_hh, w_hh_r], dy, need_dx=need_x)
            del dy
            # LSTM weight gradients from dG (overlap the next layer's BPTT; layer 0's are the exposed tail)
            with wgrad_scope(lw[:8], dG, lsaved) as sc:
                d_wcat, d_whh, d_bcat = lstm_layer_wgrad(lsaved, dG)
                Din = w_ih.shape[1]
                grads[10 * i:10 * i + 8] = sc.deliver([d_wcat[:4 * H, :Din], d_whh[0], d_bcat[:4 * H], d_bcat[:4 * H],
                                                       d_wcat[4 * H:, :Din], d_whh[1], d_bcat[4 * H:], d_bcat[4 * H:]])
            if i > 0:
                dout = dx
        dinput = None
        if ctx.needs_input_grad[0]:
            Dp0 = dx.shape[1]
            dinput = dx.view(B, -1, Dp0)[:, :, :ctx.D]
            if Dp0 != ctx.D:
                dinput = dinput.contiguous()
        ctx.saved = None
        return (dinput, None, None, None, *grads)


# --------------------------------------------------------------------------------------------
# attention decoder (model.py:139-173, 283-367)
# --------------------------------------------------------------------------------------------
# Greedy decoding with an early stop (see DecoderFn.forward); switched on by Solver around its scoring loops.
GREEDY_EARLY_STOP = {"on": False, "eos": 2, "chunk": 16, "last_steps": None}

DEC_WEIGHTS = ("emb_w", "w_ih", "w_hh", "b_ih", "b_hh", "out_w", "out_b", "mlp_enc_w", "mlp_enc_b", "mlp_dec_w",
               "mlp_att_w", "conv_w", "gvec_w", "mlp_o_w", "mlp_o_b")


def _dec_common(enc_h, W, L, K):
    B, Te, H = enc_h.shape
    Hd = W["w_hh"].shape[1]
    O = W["mlp_o_w"].shape[0]
    A = W["mlp_enc_w"].shape[0]
    V, E = W["emb_w"].shape
    C = W["conv_w"].shape[0]
    a = DecArgs()
    a.B, a.L, a.Te, a.Hd, a.O, a.A, a.V, a.E, a.H, a.C, a.K = B, L, Te, Hd, O, A, V, E, H, C, K
    return a, (B, Te, H, Hd, O, A, V, E, C)


def _build_dec_fwd(W, mode):
    """Weight-only operands of the decoder forward (see prepare_ahead)."""
    V, E = W["emb_w"].shape
    Hd = W["w_hh"].shape[1]
    C = W["conv_w"].shape[0]
    Ep = _r16(E)
    d = {}
    d["mlp_enc_bf"] = cvt_bf16(W["mlp_enc_w"])
    d["wr_cat"] = torch.cat([W["w_hh"], W["w_ih"][:, E:]], dim=1).contiguous()      # [4Hd, Hd+O]
    d["wr_pk"] = pack_afrag(d["wr_cat"], 1, Hd)
    d["mlp_dec_pk"] = pack_afrag(W["mlp_dec_w"], 0)
    d["mlp_o_pk"] = pack_afrag(W["mlp_o_w"], 0)
    d["cell_bias"] = (W["b_ih"] + W["b_hh"]).contiguous()
    d["conv_w"] = W["conv_w"].reshape(C, -1).contiguous()
    d["mlp_att"] = W["mlp_att_w"].contiguous()
    d["gvec"] = W["gvec_w"].reshape(-1).contiguous()
    we = W["w_ih"][:, :E]
    d["we_bf"] = cvt_bf16(we, ld_dst=Ep)
    if mode != 0:
        d["we_pk"] = pack_afrag(we.contiguous(), 1, Hd)
        d["out_pk"] = pack_afrag(W["out_w"], 0)
    if DEC_PERSISTENT and mode in (0, 1):
        d["mlp_o_bf"] = cvt_bf16(W["mlp_o_w"])
        d["wr2_pk"] = pack_afrag(d["wr_cat"], 3, Hd)               # quad tiles, quad-permuted K (persistent kernel)
        d["mlp_dec_pk_p"] = pack_afrag(W["mlp_dec_w"], 4)
    d["out_bf"] = cvt_bf16(W["out_w"])
    return d


def _build_dec_bwd(W, mode, wr_cat, persistent):
    """Weight-only operands of the decoder backward."""
    V, E = W["emb_w"].shape
    Hd = W["w_hh"].shape[1]
    O, A = W["mlp_o_w"].shape[0], W["mlp_enc_w"].shape[0]
    dev = wr_cat.device
    d = {}
    d["wrT_pk"] = pack_afrag(wr_cat, 0, transposed=True)
    d["mlp_oT_pk"] = pack_afrag(W["mlp_o_w"], 0, transposed=True)
    d["mlp_decT_pk"] = pack_afrag(W["mlp_dec_w"], 0, transposed=True)
    if mode == 2:
        d["weT_pk"] = pack_afrag(W["w_ih"][:, :E].contiguous(), 0, transposed=True)
        d["outT_pk"] = pack_afrag(W["out_w"], 0, transposed=True)
    if persistent:
        L_ = _lib.lib()
        ZC = Hd + O
        wrT2 = torch.empty(L_.las_dec_persistent_pack_bytes(0, Hd, O, A) // 4, device=dev, dtype=torch.int32)
        call("las_dec_persistent_pack", 0, ptr(wr_cat), ZC, Hd, O, A, ptr(wrT2))
        mlp_dec_w = W["mlp_dec_w"].contiguous()
        decT2 = torch.empty(L_.las_dec_persistent_pack_bytes(1, Hd, O, A) // 4, device=dev, dtype=torch.int32)
        call("las_dec_persistent_pack", 1, ptr(mlp_dec_w), Hd, Hd, O, A, ptr(decT2))
        d["wrT2"], d["decT2"] = wrT2, decT2
    return d


def decoder_prep_jobs(wts, mode):
    """prepare_ahead jobs of the decoder: forward operands; backward packs for the teacher-forced persistent path."""
    wts = list(wts)
    W = dict(zip(DEC_WEIGHTS, wts))
    jobs = [("dec_fwd%d" % mode, wts, lambda: _build_dec_fwd(W, mode))]
    Hd, O, A = W["w_hh"].shape[1], W["mlp_o_w"].shape[0], W["mlp_enc_w"].shape[0]
    L_ = _lib.lib()
    packable = L_.las_dec_persistent_pack_bytes(0, Hd, O, A) > 0 and L_.las_dec_persistent_pack_bytes(1, Hd, O, A) > 0
    if mode == 0 and DEC_PERSISTENT and packable:
        def bwd():
            wr_cat = torch.cat([W["w_hh"], W["w_ih"][:, W["emb_w"].shape[1]:]], dim=1).contiguous()
            return _build_dec_bwd(W, mode, wr_cat, True), wr_cat
        jobs.append(("dec_bwd%d_p" % mode, wts, bwd))
    return jobs


class DecoderFn(torch.autograd.Function):
    """Runs all L decoder steps. Returns (logits_alloc f32 [B, L+1, V] (row r = step r-1; row 0
    unused), ws_alloc f32 [B, L+1, Te] (row 0 = initial alignment), pred int64 [B, L] or None)."""

    @staticmethod
    def forward(ctx, enc_h, enc_lens, ys_in, L, mode, smooth_scaling, att_scaling, K, bos, p_drop, tf_mask, sample,
                need_grad, *wts):
        """need_grad: a backward pass may follow (the caller's grad mode: inside Function.forward autograd is always
        switched off, so `torch.is_grad_enabled()` cannot tell). tf_mask: None, or a uint8 tensor [L+1] for scheduled sampling (mode 1 with teacher tokens `ys_in`: step r
        consumes the teacher token where tf_mask[r] != 0, else the previous prediction; model.py:327-329). sample:
        the prediction is drawn from softmax(logits) instead of the argmax (mode 1, model.py:349-351)."""
        W = dict(zip(DEC_WEIGHTS, wts))
        dev = enc_h.device
        site0 = new_sites() if p_drop > 0 else 0
        a, (B, Te, H, Hd, O, A, V, E, C) = _dec_common(enc_h, W, L, K)
        R, ZC, Ep = L + 1, Hd + O, _r16(E)
        a.mode, a.att_scaling, a.smooth_scaling = mode, att_scaling, smooth_scaling
        f32 = dict(device=dev, dtype=torch.float32)
        keep = []  # keep tensors referenced by raw pointers alive until the launches are enqueued
        if p_drop > 0:                                 # cell-input dropout (model.py:285)
            a.drop_p, a.drop_site, a.seed_dev = float(p_drop), site0, ptr(dropout_seed(dev))

        wts = list(wts)
        Pk = _prep("dec_fwd%d" % mode, wts, lambda: _build_dec_fwd(W, mode))
        enc_bf = cvt_bf16(enc_h.reshape(B * Te, H))
        mlp_enc_bf = Pk["mlp_enc_bf"]
        Pm = gemm(enc_bf, H, 0, mlp_enc_bf, H, 0, B * Te, A, H, bias=W["mlp_enc_b"])
        wr_cat, wr_pk, mlp_dec_pk, mlp_o_pk = Pk["wr_cat"], Pk["wr_pk"], Pk["mlp_dec_pk"], Pk["mlp_o_pk"]
        cell_bias = Pk["cell_bias"]
        ws, zc, cx, c_state = zeros_many(dev, [((B, R, Te), torch.float32), ((B * R * ZC + 64,), BF16),
                                               ((B * R * H + 64,), BF16), ((B, Hd), torch.float32)])
        call("las_att_init", ptr(enc_lens), B, Te, ptr(ws), R * Te)
        e_buf = torch.empty(B, Te, **f32)
        dzf = torch.empty(B, L, A, **f32)
        gates = torch.empty(B, L, Hd, 4, device=dev, dtype=torch.float16)
        csave = torch.empty(B, L, Hd, **f32)
        conv_w, mlp_att, gvec = Pk["conv_w"], Pk["mlp_att"], Pk["gvec"]
        a.enc_h, a.P = ptr(enc_bf), ptr(Pm)
        a.wr_pk, a.mlp_dec_pk, a.mlp_o_pk, a.mlp_o_b = ptr(wr_pk), ptr(mlp_dec_pk), ptr(mlp_o_pk), ptr(W["mlp_o_b"])
        a.conv_w, a.mlp_att, a.gvec = ptr(conv_w), ptr(mlp_att), ptr(gvec)
        a.ws, a.zc, a.ctx, a.c_state, a.e_buf, a.dzf = ptr(ws), ptr(zc), ptr(cx), ptr(c_state), ptr(e_buf), ptr(dzf)
        a.gates_save, a.c_save = ptr(gates), ptr(csave)
        emb_in = None
        pred = None
        logits = None
        if mode == 0:
            # teacher forcing: the embedding half of the LSTMCell input projection for all steps at once
            emb_in = torch.empty(B * R, Ep, device=dev, dtype=BF16)
            call("las_gather_rows_bf16", ptr(W["emb_w"]), E, ptr(ys_in), B * R, ptr(emb_in), Ep)
            if p_drop > 0:
                dropout_(emb_in, B, R, E, R * Ep, Ep, 0, p_drop, site0 + 1)
            we_bf = Pk["we_bf"]
            embx = gemm(emb_in, Ep, 0, we_bf, Ep, 0, B * R, 4 * Hd, Ep, bias=cell_bias)
            a.embx = ptr(embx)
            keep += [embx, we_bf]
        else:
            we_pk, out_pk = Pk["we_pk"], Pk["out_pk"]
            emb_op = torch.zeros(B * R * Ep + 64, device=dev, dtype=BF16)
            bos_row = torch.zeros(Ep, **f32)
            bos_row[:E] = W["emb_w"][bos]
            emb_op[:B * R * Ep].view(B, R, Ep)[:, 0, :] = bos_row.to(BF16)
            if p_drop > 0:                             # row 0 now; later rows are masked as the kernels write them
                dropout_(emb_op, B, R, E, R * Ep, Ep, 0, p_drop, site0 + 1)
            logits = torch.zeros(B, R, V, **f32)
            pred = torch.zeros(B, L, device=dev, dtype=torch.int64)
            a.cell_bias, a.we_pk, a.out_pk, a.out_b = ptr(cell_bias), ptr(we_pk), ptr(out_pk), ptr(W["out_b"])
            a.emb_w, a.emb_op, a.logits, a.pred = ptr(W["emb_w"]), ptr(emb_op), ptr(logits), ptr(pred)
            keep += [we_pk, out_pk, emb_op]
        persist, pers = False, None
        es = GREEDY_EARLY_STOP
        # greedy decoding without autograd (Solver.validation / test, solver.py:212-286) may run as ONE cluster-persistent
        # launch too: the embedding half of the cell input becomes a per-token table, the argmax feeds back in-kernel
        greedy_p = (mode == 1 and DEC_PERSISTENT and p_drop == 0 and not need_grad and ZC % 8 == 0
                    and tf_mask is None and not sample)
        if mode == 1 and (tf_mask is not None or sample):
            if tf_mask is not None:
                a.tok_teacher, a.tf_mask = ptr(ys_in), ptr(tf_mask)
            a.sample = int(bool(sample))
            a.seed_dev, a.drop_site = ptr(dropout_seed(dev)), (site0 if p_drop > 0 else new_sites())
        if greedy_p:
            emb_all = torch.empty(V, Ep, device=dev, dtype=BF16)
            call("las_gather_rows_bf16", ptr(W["emb_w"]), E, ptr(torch.arange(V, device=dev)), V, ptr(emb_all), Ep)
            emb_tab = gemm(emb_all, Ep, 0, Pk["we_bf"], Ep, 0, V, 4 * Hd, Ep, bias=cell_bias)      # f32 [V, 4Hd]
            a.embx, a.out_bf, a.bos_token = ptr(emb_tab), ptr(Pk["out_bf"]), int(bos)
            a.stop_token = int(es["eos"]) if es["on"] else -1
            keep += [emb_all, emb_tab]
        if (mode == 0 or greedy_p) and DEC_PERSISTENT:
            # cluster-persistent decoder: c_t = w_t @ Q + bias with Q = enc_h @ mlp_o.weight^T
            # Q is stored centred over the frames of each utterance (the mean goes into a per-utterance
            # bias): bf16 rounding then applies to the deviations, which is what the softmax backward
            # (dw - <w, dw>, a cancellation when the alignment is flat) actually consumes
            mlp_o_bf = Pk["mlp_o_bf"]
            Qf = gemm(enc_bf, H, 0, mlp_o_bf, H, 0, B * Te, O, H).view(B, Te, O)
            Qbar = Qf.mean(dim=1, keepdim=True)
            Qm = (Qf - Qbar).to(BF16).view(B * Te, O)
            cbias = (Qbar.view(B, O) + W["mlp_o_b"]).contiguous()
            # likewise P = mlp_enc(enc_h): the kernels hold P - mean_te(P) in bf16 and add the mean to dz in f32
            Pbar = Pm.view(B, Te, A).mean(dim=1)
            Pc = (Pm.view(B, Te, A) - Pbar.unsqueeze(1)).contiguous()
            wr2_pk = Pk["wr2_pk"]
            a.Q, a.wr2_pk, a.cbias, a.pbar = ptr(Qm), ptr(wr2_pk), ptr(cbias), ptr(Pbar)
            a.mlp_dec_pk_p = ptr(Pk["mlp_dec_pk_p"])
            persist = bool(_lib.lib().las_dec_persistent_supported(ctypes.byref(a)))
            if persist and mode == 0:
                cpre = torch.empty(B, L, O, **f32)
                conv_save = torch.empty(B, L, Te, 16, **f32)
                a.cpre, a.conv_save = ptr(cpre), ptr(conv_save)
                a.P = ptr(Pc)
                pers = dict(Qm=Qm, wr2_pk=wr2_pk, cpre=cpre, conv_save=conv_save, mlp_o_bf=mlp_o_bf, cbias=cbias, Pc=Pc, Pbar=Pbar)
            elif persist:
                a.P = ptr(Pc)                         # inference: nothing is saved for a backward pass
                keep += [Qm, cbias, Pc, Pbar]
            else:
                a.Q, a.wr2_pk, a.cbias, a.pbar = None, None, None, None
                if greedy_p:
                    a.embx, a.out_bf = None, None
        if p_drop > 0 and not persist:
            zcd = torch.zeros(B * R * ZC + 64, device=dev, dtype=BF16)
            a.zcd = ptr(zcd)
            keep.append(zcd)
        conv_save_steps = None
        if not persist and need_grad and _lib.lib().las_att_bwd_lean_supported(A, C):
            # per-timestep path with a backward to come: the energy kernels also save the location-conv features, and
            # the time loop of the backward runs its lean tensor-core energy kernel (dP and the energy-MLP parameter
            # sums follow after the loop from these features and the energy gradients, as after the persistent kernel)
            conv_save_steps = torch.empty(B, L, Te, 16, **f32)
            a.conv_save = ptr(conv_save_steps)
        if mode == 1 and persist:
            call("las_dec_fwd", ctypes.byref(a))      # all steps (or up to every utterance's <EOS>) in one launch
            es["last_steps"] = L
        elif mode == 1 and es["on"] and not need_grad:
            # greedy decoding for scoring (Solver.test / validation): issue the steps in chunks and stop as soon as
            # every utterance has emitted <EOS>. Rows after the stop keep their initial zeros; the hypotheses are
            # identical after remove_pad_eos (utils.py:192-201), which cuts at the first <EOS>.
            t0 = 0
            while t0 < L:
                a.t_begin, a.t_end = t0, min(L, t0 + es["chunk"])
                call("las_dec_fwd", ctypes.byref(a))
                t0 = a.t_end
                if t0 < L and bool((pred[:, :t0] == es["eos"]).any(dim=1).all()):     # one host sync per chunk
                    break
            es["last_steps"] = t0
        else:
            call("las_dec_fwd", ctypes.byref(a))
        out_bf = Pk["out_bf"]                                                         # [V, ZC]
        tok_in = None
        if mode == 1:
            # the token each step consumed: <BOS>, then the previous prediction -- or the teacher's token where the
            # scheduled-sampling mask says so (the backward scatters the embedding gradient to these rows)
            tok_in = torch.cat([torch.full((B, 1), int(bos), device=dev, dtype=torch.int64), pred], dim=1)     # [B, R]
            if tf_mask is not None:
                tok_in = torch.where(tf_mask.to(torch.bool).unsqueeze(0), ys_in, tok_in)
        if mode == 0:
            logits = gemm(zc, ZC, 0, out_bf, ZC, 0, B * R, V, ZC, bias=W["out_b"]).view(B, R, V)
        ctx.geom = (B, Te, H, Hd, O, A, V, E, C, K, L, mode, att_scaling, smooth_scaling)
        ctx.drop = (float(p_drop), site0)
        ctx.saved = dict(enc_bf=enc_bf, Pm=Pm, mlp_enc_bf=mlp_enc_bf, wr_cat=wr_cat, ws=ws, zc=zc, cx=cx, dzf=dzf,
                         gates=gates, csave=csave, conv_w=conv_w, mlp_att=mlp_att, gvec=gvec, emb_in=emb_in,
                         out_bf=out_bf, ys_in=ys_in, keep=keep, pers=pers, bos=bos, we_bf=Pk["we_bf"], Pk=Pk, tok_in=tok_in,
                         logits=logits if mode != 0 else None, conv_save_steps=conv_save_steps,
                         emb_op=(emb_op[:B * R * Ep].view(B * R, Ep) if mode != 0 else None))
        ctx.W = W
        ctx.wts = wts
        ctx.mark_non_differentiable(ws)
        if pred is not None:
            ctx.mark_non_differentiable(pred)
            return logits, ws, pred
        return logits, ws, None

    @staticmethod
    def backward(ctx, dlogits, _dws, _dpred):
        B, Te, H, Hd, O, A, V, E, C, K, L, mode, att_scaling, smooth_scaling = ctx.geom
        S, W = ctx.saved, ctx.W
        dev = dlogits.device
        R, ZC, Ep = L + 1, Hd + O, _r16(E)
        f32 = dict(device=dev, dtype=torch.float32)
        n = B * R
        a, _ = _dec_common(S["enc_bf"].view(B, Te, H), W, L, K)
        a.mode, a.att_scaling, a.smooth_scaling = mode, att_scaling, smooth_scaling
        p_drop, site0 = ctx.drop
        if p_drop > 0:
            a.drop_p, a.drop_site, a.seed_dev = p_drop, site0, ptr(dropout_seed(dev))
        # output layer (model.py:293): dW = dlogits^T [z;c], d[z;c] = dlogits W
        dl = dlogits.contiguous().view(n, V)
        zc = S["zc"]
        dl_tot = None
        if mode == 2:
            # smooth free-run: d logit_t also arrives through emb_{t+1} = softmax(s logit_t) @ E, so the output
            # layer's backward runs inside the time loop (las_dec_bwd fills dl_tot and dzc_all)
            Vq = (V + 3) // 4 * 4
            dl_tot = torch.zeros(n * Vq + 64, **f32)
            dzc_all = torch.zeros(n, ZC, **f32)
            demb_buf = torch.empty(B, Ep, **f32)
            a.dlogits, a.dl_tot, a.demb_buf = ptr(dl), ptr(dl_tot), ptr(demb_buf)
            a.logits, a.emb_w = ptr(S["logits"]), ptr(W["emb_w"])
        else:
            dl_bf = cvt_bf16(dl)                                                      # [n, Vp]
            Vp = dl_bf.shape[1]
            dzc_all = gemm(dl_bf, Vp, 0, S["out_bf"], ZC, 1, n, ZC, V)                # f32 [n, ZC]
        pers = S["pers"]
        wts = list(ctx.wts)
        if mode == 0 and pers is not None:
            Bk, _ = _prep("dec_bwd%d_p" % mode, wts, lambda: (_build_dec_bwd(W, mode, S["wr_cat"], True), S["wr_cat"]))
        else:
            Bk = _build_dec_bwd(W, mode, S["wr_cat"], pers is not None)
        wrT_pk, mlp_oT_pk, mlp_decT_pk = Bk["wrT_pk"], Bk["mlp_oT_pk"], Bk["mlp_decT_pk"]
        if mode == 2:
            a.weT_pk, a.outT_pk = ptr(Bk["weT_pk"]), ptr(Bk["outT_pk"])
        dcz_tot = torch.empty(B, ZC, **f32)
        dctx_all = torch.empty(B, L, H, **f32)
        dw_buf = torch.empty(B, Te, **f32)
        dattc_all = torch.empty(L, B, Te, C, **f32)
        att_part = torch.empty(_lib.lib().las_att_scratch_floats(B, L, Te, A, C, K), **f32)
        dc_state = torch.empty(B, Hd, **f32)
        F32 = torch.float32
        zspecs = [((n, ZC), BF16), ((n * A + 64,), F32), ((B * Te, A), F32), ((n, 4 * Hd), BF16), ((A, C), F32), ((A,), F32),
                  ((C, 2 * K + 1), F32)]
        lean = pers is None and S.get("conv_save_steps") is not None
        if pers is not None or lean:
            zspecs += [((B, L, Te), F32), ((B, L, O), F32)]
        zb = zeros_many(dev, zspecs)
        dcz_all, ddz_all, dP, dgates, d_mlp_att, d_gvec, d_conv = zb[:7]
        denc = torch.empty(B, Te, H, **f32)
        a.enc_h, a.P = ptr(S["enc_bf"]), ptr(S["Pm"])
        a.conv_w, a.mlp_att, a.gvec = ptr(S["conv_w"]), ptr(S["mlp_att"]), ptr(S["gvec"])
        a.ws, a.zc, a.ctx, a.dzf = ptr(S["ws"]), ptr(zc), ptr(S["cx"]), ptr(S["dzf"])
        a.gates_save, a.c_save = ptr(S["gates"]), ptr(S["csave"])
        a.wrT_pk, a.mlp_oT_pk, a.mlp_decT_pk = ptr(wrT_pk), ptr(mlp_oT_pk), ptr(mlp_decT_pk)
        a.dzc_all, a.dcz_tot, a.dcz_all, a.dctx_all, a.dw_buf = (ptr(dzc_all), ptr(dcz_tot), ptr(dcz_all),
                                                                 ptr(dctx_all), ptr(dw_buf))
        a.dattc_all, a.ddz_all, a.dP, a.att_part, a.dc_state = (ptr(dattc_all), ptr(ddz_all), ptr(dP),
                                                                ptr(att_part), ptr(dc_state))
        a.dgates, a.dmlp_att, a.dgvec, a.dconv_w, a.denc = (ptr(dgates), ptr(d_mlp_att), ptr(d_gvec), ptr(d_conv),
                                                            ptr(denc))
        a.denc_accumulate = 0
        if pers is not None:
            # cluster-persistent backward: the serial loop only produces per-step gradients; every
            # parameter gradient is reduced afterwards by the GEMMs / parallel kernels below
            L_ = _lib.lib()
            wrT2, decT2 = Bk["wrT2"], Bk["decT2"]
            de_all, dc_all = zb[7], zb[8]
            a.Q, a.wr2_pk, a.cpre, a.conv_save = ptr(pers["Qm"]), ptr(pers["wr2_pk"]), ptr(pers["cpre"]), ptr(pers["conv_save"])
            a.mlp_dec_pk_p = ptr(S["Pk"]["mlp_dec_pk_p"])
            a.wrT2_pk, a.mlp_decT2_pk, a.de_all, a.dc_all = ptr(wrT2), ptr(decT2), ptr(de_all), ptr(dc_all)
            a.cbias, a.pbar, a.P = ptr(pers["cbias"]), ptr(pers["Pbar"]), ptr(pers["Pc"])
            if not L_.las_dec_persistent_supported(ctypes.byref(a)):
                raise _lib.LasError("decoder backward: the forward ran the persistent kernel but the backward would not")
        if lean:
            de_all = zb[7]
            a.conv_save, a.de_all = ptr(S["conv_save_steps"]), ptr(de_all)
        call("las_dec_bwd", ctypes.byref(a))
        dq_fork = torch.cuda.Event()
        dq_fork.record(torch.cuda.current_stream(dev))      # what the dQ chain below depends on (not the dP kernel)
        # ---- critical path: gradient w.r.t. the encoder states
        dP_bf = None
        post = pers is not None or lean           # dP / energy-MLP parameter sums are produced after the time loop
        scope = wgrad_scope(ctx.wts, S, pers, (de_all if post else None), dl, (dl_bf if mode != 2 else None),
                            dlogits, dl_tot, dgates, dcz_all, ddz_all, dP, dattc_all, att_part, d_mlp_att, d_gvec)
        if post:
            # dP is all the encoder's backward waits for; when the weight gradients are deferred, the energy-MLP
            # parameter sums (the same walk over (b, t, te, a) again, tanh recomputed) run on the side stream
            apg_P = pers["Pc"] if pers is not None else S["Pm"]
            apg_conv = pers["conv_save"] if pers is not None else S["conv_save_steps"]
            apg = (ptr(apg_P), ptr(S["dzf"]), ptr(apg_conv), ptr(de_all), ptr(S["mlp_att"]), ptr(S["gvec"]),
                   B, L, Te, A, C)
            apg_split = scope.deferred and os.environ.get("LAS_APG_SPLIT", "1") == "1"
            call("las_att_param_grads_part", *apg, 1 if apg_split else 3, ptr(dP), ptr(att_part), ptr(d_mlp_att),
                 ptr(d_gvec))
        if pers is not None:
            # c_t = w_t @ Q + b, Q = enc_h @ mlp_o.weight^T: d enc_h = dQ mlp_o.weight (+ dP mlp_enc.weight below).
            # dQ does not depend on dP: with a weight-gradient stream at hand its three kernels run there, next to the
            # dP kernel (which leaves ~20 SMs free), and the accumulating GEMM below waits for them
            def dq_chain():
                dQ = torch.empty(B * Te, O, **f32)
                call("las_att_dq", ptr(S["ws"]), ptr(dc_all), L, B, Te, O, ptr(dQ))
                dQ_bf = cvt_bf16(dQ)
                gemm(dQ_bf, dQ_bf.shape[1], 0, pers["mlp_o_bf"], H, 1, B * Te, H, O, out=denc.view(B * Te, H))
                return dQ, dQ_bf

            if scope.deferred and os.environ.get("LAS_DQ_OVERLAP", "1") == "1":
                main = torch.cuda.current_stream(dev)
                side = warm_deferred(dev)
                side.wait_event(dq_fork)
                _DEFER["used"] = True
                with torch.cuda.stream(side):
                    dQ, dQ_bf = dq_chain()
                    dq_done = torch.cuda.Event()
                    dq_done.record(side)
                _DEFER["keep"].append((dQ, dQ_bf, denc, dc_all))
                main.wait_event(dq_done)
            else:
                dQ, dQ_bf = dq_chain()
        dP_bf = cvt_bf16(dP)
        Ap8 = dP_bf.shape[1]
        gemm(dP_bf, Ap8, 0, S["mlp_enc_bf"], H, 1, B * Te, H, A, out=denc.view(B * Te, H), accumulate=True)
        # ---- weight gradients deferred out of the time loop, as dense contractions over all (b, t); nothing on
        # the critical path (the encoder's backward) waits for them
        scope.keep = scope.keep + (dP_bf, (dQ_bf if pers is not None else None))
        with scope as sc:
            if pers is not None:
                d_conv_t = torch.zeros(C, 2 * K + 1, **f32)
                call("las_att_dconv", ptr(dattc_all), ptr(S["ws"]), L, B, Te, C, K, ptr(d_conv_t), ptr(att_part))
                d_conv = d_conv_t
            if mode == 2:
                dlt = dl_tot[:n * Vq].view(n, Vq)[:, :V]
                dlt_bf = cvt_bf16(dlt)
                d_out_w = gemm(dlt_bf, dlt_bf.shape[1], 1, zc, ZC, 1, V, ZC, n)
                d_out_b = colsum(dlt, V)
            else:
                d_out_w = gemm(dl_bf, Vp, 1, zc, ZC, 1, V, ZC, n)
                d_out_b = colsum(dl, V)
            zc_in = zc
            if p_drop > 0:                                 # the cell saw [z_{t-1} | dropout(c_{t-1})]
                zc_in = zc.clone()
                dropout_(zc_in[Hd:], B, R, O, R * ZC, ZC, 0, p_drop, site0)
            d_wr = gemm(dgates, 4 * Hd, 1, zc_in, ZC, 1, 4 * Hd, ZC, n)                   # [4Hd, Hd+O]
            emb_in = S["emb_in"] if mode == 0 else S["emb_op"]                            # bf16 [n, Ep] step inputs
            d_we = gemm(dgates, 4 * Hd, 1, emb_in, Ep, 1, 4 * Hd, Ep, n)                  # [4Hd, Ep]
            d_w_ih = torch.cat([d_we[:, :E], d_wr[:, Hd:]], dim=1)
            d_w_hh = d_wr[:, :Hd].contiguous()
            d_b = colsum(dgates, 4 * Hd)
            we_bf = S["we_bf"]
            demb_rows = gemm(dgates, 4 * Hd, 0, we_bf, Ep, 1, n, Ep, 4 * Hd)              # f32 [n, Ep]
            if p_drop > 0:                                 # gradient w.r.t. the un-dropped step embeddings
                dropout_(demb_rows, B, R, E, R * Ep, Ep, 0, p_drop, site0 + 1)
            if mode == 0:
                d_emb = torch.zeros(V, E, **f32)
                call("las_scatter_add_rows", ptr(demb_rows), Ep, E, ptr(S["ys_in"]), n, 0, ptr(d_emb))
            else:
                # emb_0 = E[BOS]; emb_{t+1} = p_t @ E with p_t = softmax(s logit_t) (smooth, model.py:341: a matmul,
                # so every row of E receives gradient) or one-hot(argmax) (greedy): d E = sum_rows p^T demb
                lg = S["logits"]                                                          # [B, R, V], row r = step r-1
                if mode == 2:
                    p_all = torch.softmax(lg * smooth_scaling, dim=-1)
                    p_all = torch.cat([torch.nn.functional.one_hot(torch.full((B, 1), S["bos"], device=dev), V).float(),
                                       p_all[:, 1:]], dim=1)                              # row r feeds step r
                else:
                    p_all = torch.nn.functional.one_hot(S["tok_in"], V).float()          # the tokens the steps consumed
                p_bf = cvt_bf16(p_all.reshape(n, V))
                de_bf = cvt_bf16(demb_rows)
                d_emb = gemm(p_bf, p_bf.shape[1], 1, de_bf, de_bf.shape[1], 1, V, Ep, n)[:, :E].contiguous()
            dc_att = dcz_all[:, Hd:]                                                      # bf16 view, ld ZC
            if pers is not None:
                d_mlp_o_w = gemm(dQ_bf, dQ_bf.shape[1], 1, S["enc_bf"], H, 1, O, H, B * Te)   # d mlp_o.weight = dQ^T enc_h
            else:
                d_mlp_o_w = gemm(dc_att, ZC, 1, S["cx"], H, 1, O, H, n)
            d_mlp_o_b = colsum(dc_att, O, ld=ZC, rows=n)
            ddz_bf = cvt_bf16(ddz_all[:n * A].view(n, A))
            d_mlp_dec = gemm(ddz_bf, ddz_bf.shape[1], 1, zc, ZC, 1, A, Hd, n)
            d_mlp_enc_w = gemm(dP_bf, Ap8, 1, S["enc_bf"], H, 1, A, H, B * Te)
            d_mlp_enc_b = colsum(dP, A)
            grads = dict(emb_w=d_emb, w_ih=d_w_ih, w_hh=d_w_hh, b_ih=d_b, b_hh=d_b, out_w=d_out_w, out_b=d_out_b,
                         mlp_enc_w=d_mlp_enc_w, mlp_enc_b=d_mlp_enc_b, mlp_dec_w=d_mlp_dec, mlp_att_w=d_mlp_att,
                         conv_w=d_conv.view_as(W["conv_w"]), gvec_w=d_gvec.view_as(W["gvec_w"]), mlp_o_w=d_mlp_o_w,
                         mlp_o_b=d_mlp_o_b)
            if post and apg_split:
                # This kernel fills every free SM with 150 us CTAs; run right away it delays the small kernels between
                # the BPTT launches (each finds no free CTA slot). It is parked until the encoder's layer-0 BPTT, the
                # longest serial kernel of the step, has the main stream to itself.
                w_matt, w_gvec = ctx.wts[DEC_WEIGHTS.index("mlp_att_w")], ctx.wts[DEC_WEIGHTS.index("gvec_w")]
                hold = (apg_P, S["dzf"], apg_conv, de_all, S["mlp_att"], S["gvec"], dP, att_part)

                def late(apg=apg, hold=hold, d_mlp_att=d_mlp_att, d_gvec=d_gvec, w_matt=w_matt, w_gvec=w_gvec):
                    call("las_att_param_grads_part", *apg, 2, ptr(hold[6]), ptr(hold[7]), ptr(d_mlp_att), ptr(d_gvec))
                    if w_matt.requires_grad:
                        w_matt.grad.add_(d_mlp_att.view_as(w_matt.grad))
                    if w_gvec.requires_grad:
                        w_gvec.grad.add_(d_gvec.view_as(w_gvec.grad))
                    _DEFER["keep"].append((hold, d_mlp_att, d_gvec))

                _DEFER["late"].append((dev, late))
                grads["mlp_att_w"] = None
                grads["gvec_w"] = None
            glist = sc.deliver([grads[k] for k in DEC_WEIGHTS])
        ctx.saved = None
        return (denc, None, None, None, None, None, None, None, None, None, None, None, None, *glist)


# --------------------------------------------------------------------------------------------
# stand-alone single steps of the reference's module surface (AttLoc.forward model.py:139-173, Decoder.forward_step
# model.py:283-294, LM.forward_step model.py:535-542). Inference-only thin callers of the per-step kernels: training
# goes through DecoderFn / LMFn, which run whole sequences. Results carry bf16 rounding of z, c and h (the kernels'
# state operands are bf16).
# --------------------------------------------------------------------------------------------
def attention_precompute(enc_pad, W):
    """The per-utterance cache of AttLoc (model.py:141-144): (enc_bf [B*Te, H] bf16, P = mlp_enc(enc_h) f32 [B*Te, A])."""
    B, Te, H = enc_pad.shape
    enc_bf = cvt_bf16(enc_pad.float().contiguous().reshape(B * Te, H))
    mlp_enc_bf = cvt_bf16(W["mlp_enc_w"])
    A = W["mlp_enc_w"].shape[0]
    return enc_bf, gemm(enc_bf, H, 0, mlp_enc_bf, H, 0, B * Te, A, H, bias=W["mlp_enc_b"])


def _step_args(W, enc_bf, Pm, B, Te, K, att_scaling):
    """las_dec_args for ONE step (L = 1, two rows per buffer) with the attention operands filled in."""
    dev = enc_bf.device
    H = enc_bf.shape[1]
    a, (_, _, _, Hd, O, A, V, E, C) = _dec_common(enc_bf.view(B, Te, H), W, 1, K)
    a.mode, a.att_scaling, a.smooth_scaling = 1, float(att_scaling), 1.0
    ZC = Hd + O
    ws, zc, cx = zeros_many(dev, [((B, 2, Te), torch.float32), ((B * 2 * ZC + 64,), BF16), ((B * 2 * H + 64,), BF16)])
    e_buf = torch.empty(B, Te, device=dev, dtype=torch.float32)
    dzf = torch.empty(B, 1, A, device=dev, dtype=torch.float32)
    ops = dict(mlp_dec_pk=pack_afrag(W["mlp_dec_w"], 0), mlp_o_pk=pack_afrag(W["mlp_o_w"], 0),
               conv_w=W["conv_w"].reshape(C, -1).contiguous(), mlp_att=W["mlp_att_w"].contiguous(),
               gvec=W["gvec_w"].reshape(-1).contiguous(), mlp_o_b=W["mlp_o_b"].contiguous())
    a.enc_h, a.P = ptr(enc_bf), ptr(Pm)
    a.mlp_dec_pk, a.mlp_o_pk, a.mlp_o_b = ptr(ops["mlp_dec_pk"]), ptr(ops["mlp_o_pk"]), ptr(ops["mlp_o_b"])
    a.conv_w, a.mlp_att, a.gvec = ptr(ops["conv_w"]), ptr(ops["mlp_att"]), ptr(ops["gvec"])
    a.ws, a.zc, a.ctx, a.e_buf, a.dzf = ptr(ws), ptr(zc), ptr(cx), ptr(e_buf), ptr(dzf)
    return a, dict(ws=ws, zc=zc.narrow(0, 0, B * 2 * ZC).view(B, 2, ZC), cx=cx, e_buf=e_buf, dzf=dzf, ops=ops), (Hd, O, A, V, E, C)


def _initial_alignment(bufs, enc_lens, att_prev, B, Te):
    if att_prev is None:
        call("las_att_init", ptr(lens_tensor(enc_lens, bufs["ws"].device)), B, Te, ptr(bufs["ws"]), 2 * Te)
    else:
        bufs["ws"][:, 0].copy_(att_prev.reshape(B, Te))


@torch.no_grad()
def attention_step(W, enc_bf, Pm, B, Te, enc_lens, dec_z, att_prev, K, scaling):
    """AttLoc.forward for one decoder state: -> (c f32 [B, O], w f32 [B, Te])."""
    a, bufs, (Hd, O, A, V, E, C) = _step_args(W, enc_bf, Pm, B, Te, K, scaling)
    _initial_alignment(bufs, enc_lens, att_prev, B, Te)
    if dec_z is not None:
        bufs["zc"][:, 1, :Hd].copy_(dec_z.reshape(B, Hd))
    call("las_att_step", ctypes.byref(a), 0)
    return bufs["zc"][:, 1, Hd:].float(), bufs["ws"][:, 1].clone()


@torch.no_grad()
def decoder_step(W, enc_bf, Pm, B, Te, enc_lens, emb, dec_z, dec_c, c, w, K, p_drop, att_scaling=2.0):
    """Decoder.forward_step: dropout([emb; c]) -> LSTMCell -> attention -> output layer.
    -> (logit [B, V], dec_z [B, Hd], dec_c [B, Hd], c [B, O], w [B, Te]), all f32."""
    dev = enc_bf.device
    a, bufs, (Hd, O, A, V, E, C) = _step_args(W, enc_bf, Pm, B, Te, K, att_scaling)
    ZC, Ep = Hd + O, _r16(E)
    _initial_alignment(bufs, enc_lens, w, B, Te)
    zc = bufs["zc"]
    zc[:, 0, :Hd].copy_(dec_z.reshape(B, Hd))
    zc[:, 0, Hd:].copy_(c.reshape(B, O))
    c_state = dec_c.reshape(B, Hd).float().clone()
    emb_op = torch.zeros(B * 2 * Ep + 64, device=dev, dtype=BF16)
    emb_rows = emb_op[:B * 2 * Ep].view(B, 2, Ep)
    emb_rows[:, 0, :E].copy_(emb.reshape(B, E))
    wr_cat = torch.cat([W["w_hh"], W["w_ih"][:, E:]], dim=1).contiguous()
    ops = dict(wr_pk=pack_afrag(wr_cat, 1, Hd), we_pk=pack_afrag(W["w_ih"][:, :E].contiguous(), 1, Hd),
               out_pk=pack_afrag(W["out_w"], 0), cell_bias=(W["b_ih"] + W["b_hh"]).contiguous())
    logits = torch.zeros(B, 2, V, device=dev, dtype=torch.float32)
    pred = torch.zeros(B, 1, device=dev, dtype=torch.int64)
    gates = torch.empty(B, 1, Hd, 4, device=dev, dtype=torch.float16)
    csave = torch.empty(B, 1, Hd, device=dev, dtype=torch.float32)
    a.wr_pk, a.we_pk, a.out_pk, a.cell_bias, a.out_b = (ptr(ops["wr_pk"]), ptr(ops["we_pk"]), ptr(ops["out_pk"]),
                                                       ptr(ops["cell_bias"]), ptr(W["out_b"]))
    a.emb_w, a.emb_op, a.logits, a.pred, a.c_state = ptr(W["emb_w"]), ptr(emb_op), ptr(logits), ptr(pred), ptr(c_state)
    a.gates_save, a.c_save = ptr(gates), ptr(csave)
    zcd = None
    if p_drop > 0:                                     # model.py:284-285: dropout over the concatenated [emb; c]
        site0 = new_sites()
        a.drop_p, a.drop_site, a.seed_dev = float(p_drop), site0, ptr(dropout_seed(dev))
        dropout_(emb_op, B, 2, E, 2 * Ep, Ep, 0, p_drop, site0 + 1)
        zcd = torch.zeros(B * 2 * ZC + 64, device=dev, dtype=BF16)
        zcd[:B * 2 * ZC].view(B, 2, ZC).copy_(zc)
        dropout_(zcd[Hd:], B, 2, O, 2 * ZC, ZC, 0, p_drop, site0)
        a.zcd = ptr(zcd)
    call("las_dec_fwd", ctypes.byref(a))
    return logits[:, 1].clone(), zc[:, 1, :Hd].float(), c_state, zc[:, 1, Hd:].float(), bufs["ws"][:, 1].clone()


def lm_step_operands(lstm_weights, out_w):
    """Weight packs of LM.forward_step (build once per decode): per layer (whh_pk, wih_pk, bias, H, Kin), output pack."""
    layers = []
    for w_ih, w_hh, b_ih, b_hh in lstm_weights:
        H = w_hh.shape[1]
        layers.append((pack_afrag(w_hh, 1, H), pack_afrag(w_ih, 1, H), (b_ih + b_hh).contiguous(), H, w_ih.shape[1]))
    return layers, pack_afrag(out_w, 0)


@torch.no_grad()
def lm_step(ops, out_b, V, x, h, c, p_drop=0.0):
    """One timestep of the n-layer LM LSTM with explicit state + the output layer (model.py:535-542).
    x f32 [B, E]; h, c f32 [n_layers, B, H] or None (zeros). -> (logit [B, V], h, c) f32."""
    layers, out_pk = ops
    dev = x.device
    B = x.shape[0]
    n = len(layers)
    H = layers[0][3]
    h_new = torch.empty(n, B, H, device=dev, dtype=torch.float32)
    c_new = torch.zeros(n, B, H, device=dev, dtype=torch.float32) if c is None else c.float().clone().contiguous()
    cur = x.reshape(B, -1).float()
    site0 = new_sites() if p_drop > 0 else 0
    for l, (whh_pk, wih_pk, bias, Hl, Kin) in enumerate(layers):
        Kp, Hp = _r16(Kin), _r16(Hl)
        x_bf = torch.zeros(B, Kp, device=dev, dtype=BF16)
        x_bf[:, :Kin].copy_(cur)
        h_in = torch.zeros(B, Hp, device=dev, dtype=BF16)
        if h is not None:
            h_in[:, :Hl].copy_(h[l])
        h_out = torch.zeros(B, Hp, device=dev, dtype=BF16)
        call("las_lstm_cell_step", ptr(whh_pk), ptr(wih_pk), ptr(bias), ptr(h_in), Hp, ptr(x_bf), Kp, Kin, ptr(c_new[l]),
             ptr(h_out), Hp, B, Hl)
        h_new[l].copy_(h_out[:, :Hl])
        cur = h_out[:, :Hl].float()
        if p_drop > 0 and l + 1 < n:                   # nn.LSTM(dropout=p) between layers, training mode only
            cur = cur.contiguous()
            dropout_(cur, B, 1, Hl, Hl, Hl, 0, p_drop, site0 + l)
    top = torch.zeros(B, _r16(H), device=dev, dtype=BF16)
    top[:, :H].copy_(cur)
    logit = torch.empty(B, V, device=dev, dtype=torch.float32)
    call("las_smallmm", ptr(out_pk), V, H, ptr(top), 0, top.shape[1], B, ptr(out_b), None, 0, ptr(logit), V, None, 0)
    return logit, h_new, c_new


# --------------------------------------------------------------------------------------------
# log-softmax + gather + unigram label smoothing (model.py:353-366, 523-530)
# --------------------------------------------------------------------------------------------
class CELabelSmoothFn(torch.autograd.Function):
    """logits_alloc f32 [B, R, V]; rows r0..r0+L-1 of every utterance are scored against
    targets [B, L] (None: each row's own argmax). Returns (logp [B, L], prob [B, L], pred [B, L])."""

    @staticmethod
    def forward(ctx, logits_alloc, targets, dist, ls, r0, L):
        B, R, V = logits_alloc.shape
        dev = logits_alloc.device
        logp = torch.empty(B, L, device=dev, dtype=torch.float32)
        prob = torch.empty(B, L, device=dev, dtype=torch.float32)
        pred = torch.empty(B, L, device=dev, dtype=torch.int64)
        base = logits_alloc.data_ptr() + r0 * V * 4
        call("las_ce_ls_fwd", base, V, R * V, L, B * L, V, ptr(targets), ptr(dist), float(ls), ptr(logp), ptr(prob),
             ptr(pred))
        ctx.args = (logits_alloc, targets, dist, float(ls), r0, L)
        ctx.mark_non_differentiable(pred)
        return logp, prob, pred

    @staticmethod
    def backward(ctx, g_logp, g_prob, _g_pred):
        logits_alloc, targets, dist, ls, r0, L = ctx.args
        B, R, V = logits_alloc.shape
        dl = torch.zeros_like(logits_alloc)
        gl = g_logp.contiguous() if g_logp is not None else None
        gp = g_prob.contiguous() if g_prob is not None else None
        off = r0 * V * 4
        call("las_ce_ls_bwd", logits_alloc.data_ptr() + off, V, R * V, L, B * L, V, ptr(targets), ptr(dist), ls,
             ptr(gl), ptr(gp), dl.data_ptr() + off)
        return dl, None, None, None, None, None


# --------------------------------------------------------------------------------------------
# LM "judge": embedding -> n-layer unidirectional LSTM -> Linear (model.py:492-532)
# --------------------------------------------------------------------------------------------
class LMFn(torch.autograd.Function):
    """ys_in int64 [B, Lm] (device), lens int32 [B] (device) or None -> logits f32 [B, Lm, V].
    weights: emb_w, then (w_ih, w_hh, b_ih, b_hh) per layer, then out_w, out_b."""

    @staticmethod
    def forward(ctx, ys_in, lens, n_layers, pad, p_drop, *wts):
        site0 = new_sites() if p_drop > 0 else 0
        emb_w = wts[0]
        out_w, out_b = wts[-2], wts[-1]
        B, Lm = ys_in.shape
        dev = emb_w.device
        V, E = emb_w.shape
        Ep = _r8(E)
        n = B * Lm
        xin = torch.empty(n, Ep, device=dev, dtype=BF16)
        call("las_gather_rows_bf16", ptr(emb_w), E, ptr(ys_in), n, ptr(xin), Ep)
        if p_drop > 0:                                # model.py:512
            dropout_(xin, n, 1, E, Ep, Ep, 0, p_drop, site0)
        Dp = Ep
        saved = []
        for l in range(n_layers):
            w_ih, w_hh, b_ih, b_hh = wts[1 + 4 * l:5 + 4 * l]
            H = w_hh.shape[1]
            y, lsaved = lstm_layer_fwd(xin, Dp, [w_ih], [w_hh], [b_ih], [b_hh], lens, B, Lm, Lm, 0)
            y = y.view(n, H)
            saved.append(lsaved)
            if p_drop > 0:                            # nn.LSTM(dropout=p) between layers (model.py:466) and model.py:520 on top
                dropout_(y, n, 1, H, H, H, 0, p_drop, site0 + 1 + l)
            xin, Dp = y, H
        out_bf = cvt_bf16(out_w)
        logits = gemm(xin, Dp, 0, out_bf, Dp, 0, n, V, Dp, bias=out_b).view(B, Lm, V)
        ctx.saved = (saved, xin, out_bf, ys_in, lens)
        ctx.geom = (B, Lm, V, E, Ep, n_layers, pad)
        ctx.drop = (float(p_drop), site0)
        ctx.wts = wts
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        saved, ylast, out_bf, ys_in, lens = ctx.saved
        B, Lm, V, E, Ep, n_layers, pad = ctx.geom
        wts = ctx.wts
        dev = dlogits.device
        n = B * Lm
        dl = dlogits.contiguous().view(n, V)
        dl_bf = cvt_bf16(dl)
        Vp = dl_bf.shape[1]
        Hl = ylast.shape[1]
        d_out_w = gemm(dl_bf, Vp, 1, ylast, Hl, 1, V, Hl, n)
        d_out_b = colsum(dl, V)
        dy = gemm(dl_bf, Vp, 0, out_bf, Hl, 1, n, Hl, V)
        p_drop, site0 = ctx.drop
        grads = [None] * len(wts)
        grads[-2], grads[-1] = d_out_w, d_out_b
        for l in reversed(range(n_layers)):
            w_ih, w_hh, b_ih, b_hh = wts[1 + 4 * l:5 + 4 * l]
            H = w_hh.shape[1]
            if p_drop > 0:                            # gradient of this layer's (dropped-out) output
                dropout_(dy, n, 1, H, dy.stride(0), dy.stride(0), 0, p_drop, site0 + 1 + l)
            dy, dG = lstm_layer_bwd(saved[l], [w_hh], dy)
            with wgrad_scope(wts[1 + 4 * l:5 + 4 * l], dG, saved[l]) as sc:
                d_w_ih, d_w_hh, d_b = lstm_layer_wgrad(saved[l], dG)
                grads[1 + 4 * l:5 + 4 * l] = sc.deliver([d_w_ih[:, :w_ih.shape[1]], d_w_hh[0], d_b, d_b])
        if p_drop > 0:
            dropout_(dy, n, 1, E, Ep, Ep, 0, p_drop, site0)
        d_emb = torch.zeros(V, E, device=dev, dtype=torch.float32)
        call("las_scatter_add_rows", ptr(dy), Ep, E, ptr(ys_in), n, pad, ptr(d_emb))
        grads[0] = d_emb
        ctx.saved = None
        return (None, None, None, None, None, *grads)
