"""CPU ORACLE for the LAS training-step hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module, and only as the checker or the timed CPU baseline. The product path
(semi-supervised-asr_b200/) never imports it and has no CPU fallback.

What it is: a from-scratch fp32 restatement of the arithmetic of jjery2243542/semi-supervised-ASR's
model.py / solver.py hot path, written as pure functions over a flat {state_dict name: tensor}
parameter dict, with integer/mask/padding logic in numpy and floating point in torch fp32 on CPU
(gradients through torch.autograd). Every function cites the reference file:line it follows.

Pinning: the reference ships no tests or golden vectors. The oracle is pinned
against outputs of the reference itself, generated in the build container by
tests/golden/make_golden.py (which imports /root/reference) and committed under tests/golden/;
tests/test_oracle_golden.py checks every function below against them.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

PAD, BOS, EOS = 0, 1, 2


# --------------------------------------------------------------------------------------------
# integer / mask / padding logic (bit-exact parity class)
# --------------------------------------------------------------------------------------------
def pyramid_lengths(ilens, subsample):
    """Per-layer valid lengths and padded extents of the pyramid (model.py:76-98).

    Returns (lens_per_layer_input, T_per_layer_input, enc_lens, Te): layer i consumes
    sequences of lens[i] (max T[i]); `(length + 1) // sub` at model.py:92, frame pairs at
    model.py:88-91 (an odd padded extent grows by one replicated frame first)."""
    lens = [int(v) for v in ilens]
    all_lens, all_T = [], []
    for sub in subsample:
        T = max(lens)
        all_lens.append(list(lens))
        all_T.append(T)
        if sub > 1:
            lens = [(l + 1) // sub for l in lens]
    # after the last layer the padded extent is the concat view's extent
    T_last = all_T[-1]
    if subsample[-1] > 1:
        T_last = (T_last + (T_last % 2)) // 2
    return all_lens, all_T, lens, T_last


def seq_mask(seq_len, max_len):
    """utils.py:181-190 — mask[b, t] = 1.0 if t < seq_len[b] else 0.0 (float32)."""
    seq_len = np.asarray(seq_len, dtype=np.int64)
    return (np.arange(max_len, dtype=np.int64)[None, :] < seq_len[:, None]).astype(np.float32)


def pad_list(seqs, pad_value):
    """utils.py:173-179 — right-pad 1-D integer sequences to the longest with pad_value."""
    L = max(len(s) for s in seqs)
    out = np.full((len(seqs), L), pad_value, dtype=np.int64)
    for i, s in enumerate(seqs):
        out[i, : len(s)] = np.asarray(s, dtype=np.int64)
    return out


def decoder_targets(ys, bos=BOS, eos=EOS):
    """model.py:301-306 — ys_in = [BOS, y], ys_out = [y, EOS], both padded with EOS."""
    ys = [np.asarray(y, dtype=np.int64) for y in ys]
    ys_in = pad_list([np.concatenate([[bos], y]) for y in ys], eos)
    ys_out = pad_list([np.concatenate([y, [eos]]) for y in ys], eos)
    return ys_in, ys_out


def lm_targets(ys, bos=BOS, eos=EOS):
    """model.py:496-499 — ys_in = [BOS, y, EOS x4], ys_out = [y, EOS x5], EOS padded;
    packed lengths len(y)+5 (model.py:514)."""
    ys = [np.asarray(y, dtype=np.int64) for y in ys]
    ys_in = pad_list([np.concatenate([[bos], y, [eos] * 4]) for y in ys], eos)
    ys_out = pad_list([np.concatenate([y, [eos] * 5]) for y in ys], eos)
    lens = [len(y) + 5 for y in ys]
    return ys_in, ys_out, lens


def initial_attention(enc_lens, Te):
    """model.py:151-153 — uniform 1/len over valid frames, exact zeros beyond."""
    w = np.zeros((len(enc_lens), Te), dtype=np.float32)
    for b, l in enumerate(enc_lens):
        w[b, :l] = np.float32(1.0) / np.float32(l)
    return w


def label_distribution(token_lists, vocab_size, n_items=None, pad=PAD, bos=BOS, eos=EOS):
    """solver.py:69-78 — unigram distribution with one EOS per utterance, PAD/BOS zeroed."""
    cnt = np.zeros(vocab_size)
    for y in token_lists:
        for ind in y:
            cnt[ind] += 1.0
    cnt[eos] += len(token_lists) if n_items is None else n_items
    cnt[pad] = 0
    cnt[bos] = 0
    return cnt / np.sum(cnt)


# --------------------------------------------------------------------------------------------
# LSTM pieces
# --------------------------------------------------------------------------------------------
def _lstm_cell(gates, c_prev):
    """PyTorch gate order i, f, g, o (nn.LSTM / nn.LSTMCell; model.py:67, 262)."""
    H = c_prev.shape[-1]
    i = torch.sigmoid(gates[..., 0:H])
    f = torch.sigmoid(gates[..., H:2 * H])
    g = torch.tanh(gates[..., 2 * H:3 * H])
    o = torch.sigmoid(gates[..., 3 * H:4 * H])
    c = f * c_prev + i * g
    h = o * torch.tanh(c)
    return h, c


def lstm_direction(x, lens, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of a packed LSTM (model.py:79-81 semantics): each sequence b runs over its
    own lens[b] frames (the reverse direction starts at frame lens[b]-1), outputs beyond lens[b]
    are exact zeros (pad_packed_sequence)."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    xp = x @ w_ih.t() + (b_ih + b_hh)
    lens_t = torch.as_tensor(np.asarray(lens, dtype=np.int64))
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    outs = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        m = (lens_t > t).to(x.dtype).unsqueeze(1)
        hn, cn = _lstm_cell(xp[:, t] + h @ w_hh.t(), c)
        h = m * hn + (1 - m) * h
        c = m * cn + (1 - m) * c
        outs[t] = m * hn
    return torch.stack(outs, dim=1)


def blstm_layer(x, lens, P, prefix, fast=False):
    """Bidirectional single-layer LSTM over the first max(lens) frames (model.py:79-81).
    fast=True uses torch's packed CPU LSTM kernel (the op the reference itself calls) — used by the
    timed CPU baseline; tests check it equals the explicit recurrence."""
    T = max(int(l) for l in lens)
    x = x[:, :T]
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    fw = [P[f"{prefix}.{n}"] for n in names]
    bw = [P[f"{prefix}.{n}_reverse"] for n in names]
    if fast:
        packed = torch.nn.utils.rnn.pack_padded_sequence(x, [int(l) for l in lens], batch_first=True)
        H = fw[1].shape[1]
        out, _, _ = torch._VF.lstm(packed.data, packed.batch_sizes,
                                   (x.new_zeros(2, x.shape[0], H), x.new_zeros(2, x.shape[0], H)),
                                   fw + bw, True, 1, 0.0, False, True)
        out = torch.nn.utils.rnn.PackedSequence(out, packed.batch_sizes)
        y, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True)
        return y
    yf = lstm_direction(x, lens, *fw, reverse=False)
    yb = lstm_direction(x, lens, *bw, reverse=True)
    return torch.cat([yf, yb], dim=2)


def pyramid_concat(y):
    """model.py:88-91 — if the padded extent is odd, append a copy of the padded tensor's last
    frame (replicate pad), then view frame pairs (2j, 2j+1) side by side."""
    B, T, Fd = y.shape
    if T % 2 == 1:
        y = torch.cat([y, y[:, -1:, :]], dim=1)
        T += 1
    return y.reshape(B, T // 2, 2 * Fd)


def encoder_forward(x, ilens, P, subsample, fast=False, prefix="encoder.enc2"):
    """pBLSTM.forward (model.py:76-98), dropout disabled. Returns (enc_h [B,Te,H], enc_lens)."""
    lens = [int(l) for l in ilens]
    for i, sub in enumerate(subsample):
        y = blstm_layer(x, lens, P, f"{prefix}.layers.{i}", fast=fast)
        if sub > 1:
            y = pyramid_concat(y)
            lens = [(l + 1) // sub for l in lens]
        x = F.relu(y @ P[f"{prefix}.project_layers.{i}.weight"].t() + P[f"{prefix}.project_layers.{i}.bias"])
    return x, lens


# --------------------------------------------------------------------------------------------
# attention + decoder
# --------------------------------------------------------------------------------------------
def attloc_step(enc_h, pre_enc, dec_z, att_prev, P, prefix="attention", scaling=2.0):
    """AttLoc.forward (model.py:139-173): location conv over the previous alignment, additive
    tanh energy, UNMASKED softmax over all Te padded frames, context, mlp_o."""
    B, Te, _ = enc_h.shape
    conv_w = P[f"{prefix}.loc_conv.weight"]           # [C,1,1,2k+1]
    C, ksz = conv_w.shape[0], conv_w.shape[3]
    k = (ksz - 1) // 2
    padded = F.pad(att_prev, (k, k))                  # zeros each side (model.py:121)
    win = padded.unfold(1, ksz, 1)                    # [B, Te, ksz]
    att_conv = win @ conv_w.view(C, ksz).t()          # [B, Te, C]
    att_conv = att_conv @ P[f"{prefix}.mlp_att.weight"].t()
    dz = dec_z @ P[f"{prefix}.mlp_dec.weight"].t()
    e = torch.tanh(pre_enc + dz.unsqueeze(1) + att_conv) @ P[f"{prefix}.gvec.weight"].view(-1)
    w = torch.softmax(scaling * e, dim=1)
    ctx = torch.einsum("bt,btd->bd", w, enc_h)
    c = ctx @ P[f"{prefix}.mlp_o.weight"].t() + P[f"{prefix}.mlp_o.bias"]
    return c, w


def decoder_forward(enc_h, enc_lens, P, ys=None, max_dec_timesteps=500, smooth=False, scaling=1.0,
                    label_smoothing=True, ls_weight=0.0, labeldist=None, training=True,
                    bos=BOS, eos=EOS, tf_draws=None):
    """Decoder.forward (model.py:296-367) with dropout disabled. ys given -> teacher forcing; ys None -> free run
    with argmax feedback, or the smooth embedding softmax(scaling*logit) @ E (model.py:341).
    tf_draws (scheduled sampling, model.py:327-329): one boolean per step -- the outcome of the reference's
    `np.random.random_sample() <= tf_rate` -- where False feeds the previous step's argmax back instead of the
    teacher's token (step 0 always takes the teacher's <BOS>). None = tf_rate 1.0 (config.yaml pins it)."""
    B, Te, _ = enc_h.shape
    emb_w = P["decoder.embedding.weight"]
    w_ih, w_hh = P["decoder.LSTMCell.weight_ih"], P["decoder.LSTMCell.weight_hh"]
    bias = P["decoder.LSTMCell.bias_ih"] + P["decoder.LSTMCell.bias_hh"]
    Hd = w_hh.shape[1]
    out_w, out_b = P["decoder.output_layer.weight"], P["decoder.output_layer.bias"]
    att_odim = P["attention.mlp_o.weight"].shape[0]
    if ys is not None:
        ys_in, ys_out = decoder_targets(ys, bos, eos)
        ys_in_t, ys_out_t = torch.from_numpy(ys_in), torch.from_numpy(ys_out)
        olength = ys_out.shape[1]
        eys = F.embedding(ys_in_t, emb_w, padding_idx=PAD)                 # model.py:261, 310
    else:
        olength = max_dec_timesteps
    pre_enc = enc_h @ P["attention.mlp_enc.weight"].t() + P["attention.mlp_enc.bias"]  # model.py:144
    dec_z = enc_h.new_zeros(B, Hd)
    dec_c = enc_h.new_zeros(B, Hd)
    c = enc_h.new_zeros(B, att_odim)
    w = torch.from_numpy(initial_attention(enc_lens, Te))
    logits, preds, ws = [], [], []
    logit = None
    for t in range(olength):
        if ys is not None:
            if tf_draws is None or t == 0 or bool(tf_draws[t]):
                emb = eys[:, t]
            else:
                emb = F.embedding(preds[-1], emb_w, padding_idx=PAD)
        elif t == 0:
            emb = F.embedding(torch.full((B,), bos, dtype=torch.long), emb_w, padding_idx=PAD)
        elif smooth:
            emb = torch.softmax(logit * scaling, dim=-1) @ emb_w
        else:
            emb = F.embedding(preds[-1], emb_w, padding_idx=PAD)
        gates = torch.cat([emb, c], dim=-1) @ w_ih.t() + dec_z @ w_hh.t() + bias      # model.py:284-286
        dec_z, dec_c = _lstm_cell(gates, dec_c)
        c, w = attloc_step(enc_h, pre_enc, dec_z, w, P)                               # scaling 2.0: SURVEY D4
        logit = torch.cat([dec_z, c], dim=-1) @ out_w.t() + out_b                     # model.py:290-293
        logits.append(logit)
        ws.append(w)
        preds.append(torch.argmax(logit, dim=-1))
    logits = torch.stack(logits, dim=1)
    log_probs = F.log_softmax(logits, dim=2)
    prediction = torch.stack(preds, dim=1)
    ws = torch.stack(ws, dim=1)
    idx = ys_out_t if ys is not None else prediction
    ys_log_probs = torch.gather(log_probs, 2, idx.unsqueeze(2)).squeeze(2)
    if label_smoothing and ls_weight > 0 and training:                                  # model.py:364-366
        ld = torch.as_tensor(np.asarray(labeldist, dtype=np.float32))
        ys_log_probs = (1 - ls_weight) * ys_log_probs + ls_weight * torch.sum(log_probs * ld, dim=2)
    return logits, ys_log_probs, prediction, ws


def e2e_forward(x, ilens, P, subsample, ys=None, fast=False, **dec_kwargs):
    """E2E.forward (model.py:439-445)."""
    enc_h, enc_lens = encoder_forward(x, ilens, P, subsample, fast=fast)
    return decoder_forward(enc_h, enc_lens, P, ys=ys, **dec_kwargs)


def masked_loss(log_probs, ys):
    """E2E.mask_and_cal_loss with mask=None (model.py:447-456): -sum(logp*mask)/sum(len+1)."""
    seq_len = [len(y) + 1 for y in ys]
    mask = torch.from_numpy(seq_mask(seq_len, log_probs.shape[1]))
    return -torch.sum(log_probs * mask) / sum(seq_len)


# --------------------------------------------------------------------------------------------
# LM ("judge")
# --------------------------------------------------------------------------------------------
def _uni_lstm_layer(x, lens, w_ih, w_hh, b_ih, b_hh):
    return lstm_direction(x, lens, w_ih, w_hh, b_ih, b_hh, reverse=False)


def lm_forward(ys, P, n_layers=2, discrete_input=True, ls_weight=0.0, labeldist=None, training=True,
               bos=BOS, eos=EOS):
    """LM.forward (model.py:492-532), dropout disabled.
    discrete_input=True: ys is a list of token lists (sorted by length, descending);
    False: ys is an int64 [B, L] array of hypotheses (BOS prepended, last dropped, no packing)."""
    if discrete_input:
        ys_in, ys_out, lens = lm_targets(ys, bos, eos)
    else:
        ys = np.asarray(ys, dtype=np.int64)
        ys_in = np.concatenate([np.full((ys.shape[0], 1), bos, dtype=np.int64), ys[:, :-1]], axis=1)
        ys_out = ys
        lens = [ys.shape[1]] * ys.shape[0]
    h = F.embedding(torch.from_numpy(ys_in), P["embedding.weight"], padding_idx=PAD)   # model.py:465, 510
    for l in range(n_layers):
        h = _uni_lstm_layer(h, lens, P[f"LSTM.weight_ih_l{l}"], P[f"LSTM.weight_hh_l{l}"],
                            P[f"LSTM.bias_ih_l{l}"], P[f"LSTM.bias_hh_l{l}"])
    logits = h @ P["output_layer.weight"].t() + P["output_layer.bias"]
    log_probs = F.log_softmax(logits, dim=2)
    probs = F.softmax(logits, dim=2)
    idx = torch.from_numpy(ys_out).unsqueeze(2)
    ys_log_probs = torch.gather(log_probs, 2, idx).squeeze(2)
    ys_probs = torch.gather(probs, 2, idx).squeeze(2)
    if ls_weight > 0 and training:                                                      # model.py:528-530
        ld = torch.as_tensor(np.asarray(labeldist, dtype=np.float32))
        ys_log_probs = (1 - ls_weight) * ys_log_probs + ls_weight * torch.sum(log_probs * ld, dim=2)
    return ys_log_probs, ys_probs, torch.argmax(logits, dim=-1)


def lm_masked_sum(vals, ys):
    """LM.mask_and_cal_sum with mask=None (model.py:565-573): sum(vals*mask)/sum(len+5)."""
    seq_len = [len(y) + 5 for y in ys]
    mask = torch.from_numpy(seq_mask(seq_len, vals.shape[1]))
    return torch.sum(vals * mask) / sum(seq_len)


# --------------------------------------------------------------------------------------------
# losses and optimiser steps (solver.py)
# --------------------------------------------------------------------------------------------
def supervised_loss(x, ilens, ys, P, subsample, ls_weight, labeldist, fast=False):
    """solver.py:375-377 — -mean over ALL B x (Lmax+1) positions, train mode."""
    _, logp, _, _ = e2e_forward(x, ilens, P, subsample, ys=ys, ls_weight=ls_weight, labeldist=labeldist,
                                training=True, fast=fast)
    return -torch.mean(logp)


def ssl_losses(lab_x, lab_ilens, lab_ys, unlab_x, unlab_ilens, P, PJ, subsample, proportion, ls_weight,
               labeldist, judge_labeldist, softmax_scaling=3.0, smooth=True, unsup_weight=0.001,
               judge_layers=2, fast=False):
    """gen_train_one_iteration (solver.py:460-483). Returns (loss, sup_loss, unsup_loss)."""
    Lu = int(unlab_x.shape[1] * proportion)                                           # solver.py:469
    _, u_logp, u_pred, _ = e2e_forward(unlab_x, unlab_ilens, P, subsample, ys=None, max_dec_timesteps=Lu,
                                       smooth=smooth, scaling=softmax_scaling, label_smoothing=False,
                                       ls_weight=ls_weight, labeldist=labeldist, training=True, fast=fast)
    _, lm_probs, _ = lm_forward(u_pred.numpy(), PJ, n_layers=judge_layers, discrete_input=False,
                                ls_weight=ls_weight, labeldist=judge_labeldist, training=True)
    mask = (u_pred != EOS).float()                                                     # solver.py:477
    unsup = -torch.sum(lm_probs * u_logp * mask) / torch.sum(mask)
    sup = supervised_loss(lab_x, lab_ilens, lab_ys, P, subsample, ls_weight, labeldist, fast=fast)
    return sup + unsup_weight * unsup, sup, unsup


def judge_loss(ys, PJ, ls_weight, labeldist, n_layers=2):
    """judge_train_one_iteration (solver.py:288-291). Returns (loss, avg_prob)."""
    logp, probs, _ = lm_forward(ys, PJ, n_layers=n_layers, discrete_input=True, ls_weight=ls_weight,
                                labeldist=labeldist, training=True)
    return -lm_masked_sum(logp, ys), lm_masked_sum(probs, ys)


def unique_params(P):
    """Parameters as nn.Module.parameters() yields them: `decoder.attention.*` aliases of
    `attention.*` (model.py:421, 427-430) are the same tensors and counted once."""
    return {k: v for k, v in P.items() if not k.startswith("decoder.attention.")}


def clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (solver.py:296, 384, 488): global L2 norm over all grads,
    coefficient max_norm/(norm+1e-6) clamped to <= 1."""
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values()))
    coef = min(max_norm / (total + 1e-6), 1.0)
    return {k: g * coef for k, g in grads.items()}, total


def adam_step(params, grads, state, lr, weight_decay=0.0, amsgrad=False, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam single step (solver.py:152-153 amsgrad=True, wd=1e-6; solver.py:171-173 plain):
    L2 weight decay folded into the gradient, bias-corrected moments, AMSGrad keeps the running max
    of the second moment."""
    b1, b2 = betas
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
    new = {}
    for k, p in params.items():
        g = grads[k]
        if weight_decay != 0:
            g = g + weight_decay * p
        m = state.setdefault(("m", k), torch.zeros_like(p))
        v = state.setdefault(("v", k), torch.zeros_like(p))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        if amsgrad:
            vmax = state.setdefault(("vmax", k), torch.zeros_like(p))
            torch.maximum(vmax, v, out=vmax)
            denom = vmax.sqrt() / math.sqrt(bc2) + eps
        else:
            denom = v.sqrt() / math.sqrt(bc2) + eps
        new[k] = p - (lr / bc1) * m / denom
    return new


def _with_grad(P):
    Pu = unique_params(P)
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in Pu.items()}
    full = dict(leaves)
    for k in P:
        if k.startswith("decoder.attention."):
            full[k] = leaves[k[len("decoder."):]]
    return leaves, full


def supervised_step(x, ilens, ys, P, opt_state, subsample, ls_weight, labeldist, lr=5e-4,
                    weight_decay=1e-6, max_grad_norm=5.0, fast=False):
    """One supervised train step (solver.py:375-385). Returns (loss, grads, grad_norm, new_params)."""
    leaves, full = _with_grad(P)
    loss = supervised_loss(x, ilens, ys, full, subsample, ls_weight, labeldist, fast=fast)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    clipped, norm = clip_grad_norm(grads, max_grad_norm)
    new = adam_step({k: v.detach() for k, v in leaves.items()}, clipped, opt_state, lr, weight_decay, amsgrad=True)
    return float(loss), grads, norm, new


def ssl_step(lab, unlab, P, PJ, opt_state, subsample, proportion, ls_weight, labeldist, judge_labeldist,
             lr=1e-4, weight_decay=1e-6, max_grad_norm=5.0, **kw):
    """One semi-supervised generator step (solver.py:460-495)."""
    leaves, full = _with_grad(P)
    lab_x, lab_ilens, lab_ys = lab
    unlab_x, unlab_ilens = unlab
    loss, sup, unsup = ssl_losses(lab_x, lab_ilens, lab_ys, unlab_x, unlab_ilens, full, PJ, subsample,
                                  proportion, ls_weight, labeldist, judge_labeldist, **kw)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    clipped, norm = clip_grad_norm(grads, max_grad_norm)
    new = adam_step({k: v.detach() for k, v in leaves.items()}, clipped, opt_state, lr, weight_decay, amsgrad=True)
    return (float(loss), float(sup), float(unsup)), grads, norm, new


def judge_step(ys, PJ, opt_state, ls_weight, labeldist, lr=2e-4, max_grad_norm=5.0, n_layers=2):
    """One judge (LM) pre-train step (solver.py:288-301)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in PJ.items()}
    loss, avg_prob = judge_loss(ys, leaves, ls_weight, labeldist, n_layers=n_layers)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    clipped, norm = clip_grad_norm(grads, max_grad_norm)
    new = adam_step({k: v.detach() for k, v in leaves.items()}, clipped, opt_state, lr, 0.0, amsgrad=False)
    return (float(loss), float(avg_prob)), grads, norm, new
