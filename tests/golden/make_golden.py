"""Generate golden fixtures by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz. The fixtures pin oracle/las_oracle.py (tests/test_oracle_golden.py).
All cases run in train mode with dropout_rate = 0 (label smoothing active, no RNG).
"""
import os
import sys
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    tbx = types.ModuleType("tensorboardX")
    tbx.SummaryWriter = type("SummaryWriter", (), {"__init__": lambda s, *a, **k: None,
                                                   "add_scalar": lambda s, *a, **k: None,
                                                   "add_text": lambda s, *a, **k: None})
    sys.modules["tensorboardX"] = tbx
    ed = types.ModuleType("editdistance")
    ed.eval = lambda a, b: 0
    sys.modules["editdistance"] = ed
    sys.path.insert(0, "/root/reference")
    import model as ref_model  # noqa
    return ref_model


def synth_batch(rng, B, Tmax, D, V, ratio=0.25, min_frac=0.4):
    lens = sorted([Tmax] + [int(rng.randint(int(min_frac * Tmax), Tmax + 1)) for _ in range(B - 1)], reverse=True)
    x = np.zeros((B, Tmax, D), dtype=np.float32)
    ys = []
    for b, l in enumerate(lens):
        x[b, :l] = rng.randn(l, D).astype(np.float32)
        L = max(2, int(round(ratio * l)))
        ys.append(rng.randint(3, V, size=L).astype(np.int64))
    return x, lens, ys


def sd_numpy(module):
    return {k: v.detach().numpy().copy() for k, v in module.state_dict().items()}


def case_supervised(ref, name, seed, B, Tmax, D, H, n_layers, subsample, V, E, A, C, ksz, ls):
    torch.manual_seed(seed)
    rng = np.random.RandomState(seed)
    x, lens, ys = synth_batch(rng, B, Tmax, D, V)
    cnt = np.zeros(V)
    for y in ys:
        for t in y:
            cnt[t] += 1
    cnt[2] += len(ys)
    labeldist = cnt / cnt.sum()
    m = ref.E2E(input_dim=D, enc_hidden_dim=H, enc_n_layers=n_layers, subsample=subsample, dropout_rate=0.0,
                dec_hidden_dim=H, att_dim=A, conv_channels=C, conv_kernel_size=ksz, att_odim=H,
                embedding_dim=E, output_dim=V, ls_weight=ls, labeldist=labeldist)
    m.train()
    params0 = sd_numpy(m)
    opt = torch.optim.Adam(m.parameters(), lr=5e-4, weight_decay=1e-6, amsgrad=True)
    xs = torch.from_numpy(x)
    yst = [torch.from_numpy(y) for y in ys]
    enc_h, enc_lens = m.encoder(xs, lens)
    logits, log_probs, prediction, ws = m(xs, lens, yst, tf_rate=1.0, sample=False)
    loss = -torch.mean(log_probs)
    val_loss = m.mask_and_cal_loss(log_probs, yst)
    opt.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().numpy().copy() for k, p in m.named_parameters()}
    norm = torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=5)
    opt.step()
    params1 = sd_numpy(m)
    # greedy free-run decode in eval mode from the ORIGINAL parameters
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params0.items()})
    m.eval()
    with torch.no_grad():
        g_logits, g_logp, g_pred, g_ws = m(xs, lens, ys=None, max_dec_timesteps=12)
    out = {"x": x, "ilens": np.array(lens), "enc_h": enc_h.detach().numpy(), "enc_lens": np.array(enc_lens),
           "logits": logits.detach().numpy(), "log_probs": log_probs.detach().numpy(),
           "prediction": prediction.numpy(), "ws": ws.detach().numpy(), "loss": np.float32(loss.item()),
           "val_loss": np.float32(val_loss.item()), "grad_norm": np.float32(float(norm)),
           "labeldist": labeldist, "subsample": np.array(subsample), "ls_weight": np.float32(ls),
           "greedy_logits": g_logits.numpy(), "greedy_logp": g_logp.numpy(), "greedy_pred": g_pred.numpy(),
           "n_ys": np.int64(len(ys))}
    for i, y in enumerate(ys):
        out[f"ys_{i}"] = y
    for k, v in params0.items():
        out["p0/" + k] = v
    for k, v in grads.items():
        out["g/" + k] = v
    for k, v in params1.items():
        out["p1/" + k] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", loss.item(), "norm", float(norm), "enc_lens", enc_lens)


def case_scheduled(ref, name, seed, B, Tmax, D, H, n_layers, subsample, V, E, A, C, ksz, ls, tf_rate, np_seed):
    """Scheduled sampling (model.py:327-329, tf_rate < 1): the reference draws np.random.random_sample() once per decoder
    step and feeds the previous ARGMAX back where the draw exceeds tf_rate. The model is first trained (tf_rate = 1) on
    its own batch until every step of the scheduled forward is decisive (top-2 margin well above the bf16 tolerance), so
    that the fed-back tokens -- and with them every later step -- are pinned by the reference."""
    torch.manual_seed(seed)
    rng = np.random.RandomState(seed)
    x, lens, ys = synth_batch(rng, B, Tmax, D, V)
    cnt = np.zeros(V)
    for y in ys:
        for t in y:
            cnt[t] += 1
    cnt[2] += len(ys)
    labeldist = cnt / cnt.sum()
    m = ref.E2E(input_dim=D, enc_hidden_dim=H, enc_n_layers=n_layers, subsample=subsample, dropout_rate=0.0,
                dec_hidden_dim=H, att_dim=A, conv_channels=C, conv_kernel_size=ksz, att_odim=H,
                embedding_dim=E, output_dim=V, ls_weight=ls, labeldist=labeldist)
    m.train()
    xs = torch.from_numpy(x)
    yst = [torch.from_numpy(y) for y in ys]
    # ... trained towards OTHER label sequences of the same lengths, so that the decisive predictions it feeds back
    # differ from the teacher's tokens (otherwise the scheduled forward would coincide with teacher forcing)
    rng2 = np.random.RandomState(seed + 1000)
    ysa = [torch.from_numpy(rng2.randint(3, V, size=len(y)).astype(np.int64)) for y in ys]
    pre_opt = torch.optim.Adam(m.parameters(), lr=5e-3)
    steps = -1
    for it in range(3000):
        np.random.seed(np_seed)
        with torch.no_grad():
            pl, _, pp, _ = m(xs, lens, yst, tf_rate=tf_rate, sample=False)
        top2 = pl.topk(2, dim=-1).values
        if bool(((top2[..., 0] - top2[..., 1]) > 4 * 3e-2 * pl.abs().max()).all()):
            steps = it
            break
        _, lp, _, _ = m(xs, lens, ysa, tf_rate=1.0, sample=False)
        pre_opt.zero_grad()
        (-torch.mean(lp)).backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=5)
        pre_opt.step()
    assert steps >= 0, "the scheduled forward never became decisive"
    m.zero_grad()
    params0 = sd_numpy(m)
    np.random.seed(np_seed)
    L = max(len(y) for y in ys) + 1
    draws = np.array([np.random.random_sample() <= tf_rate for _ in range(L)])
    np.random.seed(np_seed)
    logits, log_probs, prediction, ws = m(xs, lens, yst, tf_rate=tf_rate, sample=False)
    loss = -torch.mean(log_probs)
    loss.backward()
    grads = {k: p.grad.detach().numpy().copy() for k, p in m.named_parameters()}
    assert not draws[1:].all() and draws[1:].any(), "choose a seed that mixes teacher and model tokens"
    out = {"x": x, "ilens": np.array(lens), "logits": logits.detach().numpy(), "log_probs": log_probs.detach().numpy(),
           "prediction": prediction.numpy(), "ws": ws.detach().numpy(), "loss": np.float32(loss.item()),
           "labeldist": labeldist, "subsample": np.array(subsample), "ls_weight": np.float32(ls),
           "tf_rate": np.float32(tf_rate), "np_seed": np.int64(np_seed), "draws": draws, "n_ys": np.int64(len(ys)),
           "pretrain_steps": np.int64(steps)}
    for i, y in enumerate(ys):
        out[f"ys_{i}"] = y
    for k, v in params0.items():
        out["p0/" + k] = v
    for k, v in grads.items():
        out["g/" + k] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "pre-trained", steps, "steps; loss", loss.item(), "draws", draws.astype(int).tolist())


def case_lm_and_ssl(ref, name, seed, B, Tmax, D, H, subsample, V, E, A, C, ksz, ls, JE, JH):
    torch.manual_seed(seed)
    rng = np.random.RandomState(seed)
    x, lens, ys = synth_batch(rng, B, Tmax, D, V)
    ux, ulens, _ = synth_batch(rng, B, Tmax - 3, D, V)
    cnt = np.zeros(V)
    for y in ys:
        for t in y:
            cnt[t] += 1
    cnt[2] += len(ys)
    labeldist = cnt / cnt.sum()
    proportion = sum(len(y) for y in ys) / float(sum(lens))
    m = ref.E2E(input_dim=D, enc_hidden_dim=H, enc_n_layers=len(subsample), subsample=subsample, dropout_rate=0.0,
                dec_hidden_dim=H, att_dim=A, conv_channels=C, conv_kernel_size=ksz, att_odim=H,
                embedding_dim=E, output_dim=V, ls_weight=ls, labeldist=labeldist)
    judge = ref.LM(output_dim=V, embedding_dim=JE, hidden_dim=JH, dropout_rate=0.0, n_layers=2,
                   bos=1, eos=2, pad=0, ls_weight=ls, labeldist=labeldist)
    m.train()
    judge.train()
    # A freshly initialised model gives nearly flat logits: the free-running argmax of the unpaired pass then sits on
    # near-ties (top-2 margins of 1e-5 on logits of 0.3), where an fp32 and a bf16 implementation legitimately pick
    # different tokens and nothing after the first difference can be compared. The reference is therefore trained on
    # a paired batch (its own supervised step, solver.py:375-385) until every free-running step of the unpaired pass
    # is decisive: top-2 margin > 4 * 3e-2 * max|logit| (twice the margin tests/test_gpu_ssl.py asserts).
    # (The training batch is a DIFFERENT random paired batch: on the fixture's own paired batch the model stays far from
    # converged, so the supervised gradient keeps its full size and bf16 rounding noise stays small next to it.)
    rng_pre = np.random.RandomState(seed + 1000)
    xa, lensa, ysa = synth_batch(rng_pre, B, Tmax, D, V)
    uxs_probe = torch.from_numpy(ux)
    Lu_probe = int(uxs_probe.size(1) * proportion)
    pre_opt = torch.optim.Adam(m.parameters(), lr=5e-3)
    pretrain_steps = -1
    for it in range(3000):
        with torch.no_grad():
            pl, _, pp, _ = m(uxs_probe, ulens, ys=None, sample=False, label_smoothing=False, max_dec_timesteps=Lu_probe,
                             smooth=True, scaling=3)
        top2 = pl.topk(2, dim=-1).values
        # ... and not all <EOS> (solver.py:478 would divide 0 by 0): at least a third of the tokens are real
        if bool(((top2[..., 0] - top2[..., 1]) > 4 * 3e-2 * pl.abs().max()).all()) and float((pp != 2).float().mean()) > 0.33:
            pretrain_steps = it
            break
        _, lp, _, _ = m(torch.from_numpy(xa), lensa, ys=[torch.from_numpy(y) for y in ysa], tf_rate=1.0, sample=False)
        pre_opt.zero_grad()
        (-torch.mean(lp)).backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=5)
        pre_opt.step()
    assert pretrain_steps >= 0, "the unpaired free run never became decisive"
    print(name, "pre-trained for", pretrain_steps, "steps; min margin / max|logit| =",
          float(((top2[..., 0] - top2[..., 1]).min() / pl.abs().max())))
    m.zero_grad()
    p0 = sd_numpy(m)
    j0 = sd_numpy(judge)
    out = {"x": x, "ilens": np.array(lens), "ux": ux, "uilens": np.array(ulens), "labeldist": labeldist,
           "proportion": np.float64(proportion), "subsample": np.array(subsample), "ls_weight": np.float32(ls),
           "n_ys": np.int64(len(ys)), "pretrain_steps": np.int64(pretrain_steps)}
    for i, y in enumerate(ys):
        out[f"ys_{i}"] = y
    # ---- judge pre-train step (solver.py:288-297), text sorted by length descending
    ys_sorted = sorted(ys, key=lambda t: len(t), reverse=True)
    for i, y in enumerate(ys_sorted):
        out[f"jys_{i}"] = y
    jt = [torch.from_numpy(y) for y in ys_sorted]
    dis_opt = torch.optim.Adam(judge.parameters(), lr=2e-4)
    logp, probs, preds = judge(ys=jt, discrete_input=True)
    jloss = -judge.mask_and_cal_sum(logp, ys=jt, mask=None)
    javg = judge.mask_and_cal_sum(probs, ys=jt, mask=None)
    dis_opt.zero_grad()
    jloss.backward()
    jgr = {k: p.grad.detach().numpy().copy() for k, p in judge.named_parameters()}
    jnorm = torch.nn.utils.clip_grad_norm_(judge.parameters(), max_norm=5)
    dis_opt.step()
    j1 = sd_numpy(judge)
    out.update({"j_logp": logp.detach().numpy(), "j_probs": probs.detach().numpy(), "j_preds": preds.numpy(),
                "j_loss": np.float32(jloss.item()), "j_avg_prob": np.float32(javg.item()),
                "j_grad_norm": np.float32(float(jnorm))})
    judge.load_state_dict({k: torch.from_numpy(v) for k, v in j0.items()})
    # ---- SSL generator step (solver.py:460-489)
    gen_opt = torch.optim.Adam(m.parameters(), lr=1e-4, weight_decay=1e-6, amsgrad=True)
    uxs = torch.from_numpy(ux)
    u_logits, u_logp, u_pred, _ = m(uxs, ulens, ys=None, sample=False, label_smoothing=False,
                                    max_dec_timesteps=int(uxs.size(1) * proportion), smooth=True, scaling=3)
    _, lm_probs, _ = judge(ys=u_pred, discrete_input=False)
    mask = (u_pred != 2).float()
    unsup = -torch.sum(lm_probs * u_logp * mask) / torch.sum(mask)
    yst = [torch.from_numpy(y) for y in ys]
    _, lab_logp, _, _ = m(torch.from_numpy(x), lens, ys=yst, tf_rate=1.0, sample=False)
    sup = -torch.mean(lab_logp)
    loss = sup + 0.001 * unsup
    gen_opt.zero_grad()
    loss.backward()
    gr = {k: p.grad.detach().numpy().copy() for k, p in m.named_parameters()}
    norm = torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=5)
    gen_opt.step()
    p1 = sd_numpy(m)
    out.update({"u_logits": u_logits.detach().numpy(), "u_logp": u_logp.detach().numpy(), "u_pred": u_pred.numpy(),
                "lm_probs": lm_probs.detach().numpy(), "unsup": np.float32(unsup.item()),
                "sup": np.float32(sup.item()), "loss": np.float32(loss.item()), "grad_norm": np.float32(float(norm))})
    for k, v in p0.items():
        out["p0/" + k] = v
    for k, v in gr.items():
        out["g/" + k] = v
    for k, v in p1.items():
        out["p1/" + k] = v
    for k, v in j0.items():
        out["j0/" + k] = v
    for k, v in jgr.items():
        out["jg/" + k] = v
    for k, v in j1.items():
        out["j1/" + k] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "sup", sup.item(), "unsup", unsup.item(), "jloss", jloss.item())



def case_host_plumbing(name, seed):
    """Batch layout and host statistics straight from the reference's own functions: dataloader._collate_fn /
    _speech_collate_fn / _text_collate_fn (dataloader.py:6-24), Solver.get_label_dist and
    calculate_length_proportion (solver.py:69-85), utils.remove_pad_eos / ind2character (utils.py:192-220)."""
    import dataloader as ref_dl
    import solver as ref_solver
    import utils as ref_utils
    rng = np.random.RandomState(seed)
    V = 12
    vocab = {"<PAD>": 0, "<BOS>": 1, "<EOS>": 2, "<NOISE>": 3, "<space>": 4}
    for i, ch in enumerate("ABCDEFG"):
        vocab[ch] = 5 + i
    non_lang = ["<NOISE>", "<PAD>", "<BOS>", "<EOS>"]
    items = []
    for i in range(9):
        T = int(rng.randint(3, 15))
        items.append((rng.randn(T, 5).astype(np.float32), [int(t) for t in rng.randint(3, V, size=int(rng.randint(1, 7)))]))
    items[3] = (items[3][0][:items[1][0].shape[0]].copy() if items[3][0].shape[0] >= items[1][0].shape[0]
                else np.concatenate([items[3][0], items[3][0]])[:items[1][0].shape[0]].copy(), items[3][1])   # a tie in length
    out = {"n_items": len(items), "vocab_keys": np.array(list(vocab.keys())), "vocab_vals": np.array(list(vocab.values())),
           "non_lang": np.array(non_lang)}
    for i, (f, t) in enumerate(items):
        out[f"feat_{i}"] = f
        out[f"tok_{i}"] = np.array(t, dtype=np.int64)
    padded, ilens, texts = ref_dl._collate_fn(list(items))
    out["c_padded"], out["c_ilens"] = padded.numpy(), np.array(ilens)
    for i, t in enumerate(texts):
        out[f"c_text_{i}"] = t.numpy()
    sp, sil = ref_dl._speech_collate_fn(list(items))
    out["s_padded"], out["s_ilens"] = sp.numpy(), np.array(sil)
    for i, t in enumerate(ref_dl._text_collate_fn(list(items))):
        out[f"t_text_{i}"] = t.numpy()
    fake = types.SimpleNamespace(vocab=vocab, train_lab_dataset=items)
    out["labeldist"] = ref_solver.Solver.get_label_dist(fake, items)
    out["proportion"] = np.float64(ref_solver.Solver.calculate_length_proportion(fake))
    seqs = [[5, 6, 2, 7, 2], [2, 5], [5, 4, 6, 3, 7], []]
    cut = ref_utils.remove_pad_eos(seqs, eos=2)
    out["rpe"] = np.array([len(c) for c in cut])
    chars = ref_utils.ind2character(cut, non_lang, vocab)
    out["sents"] = np.array(ref_utils.char_list_to_str(chars))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name)


if __name__ == "__main__":
    ref = import_reference()
    torch.set_num_threads(4)
    if os.environ.get("ONLY_HOST"):
        case_host_plumbing("host_plumbing", seed=31)
        sys.exit(0)
    if os.environ.get("ONLY_SCHED"):
        case_scheduled(ref, "sched_small", seed=17, B=3, Tmax=33, D=20, H=16, n_layers=2, subsample=[2, 2], V=10, E=8,
                       A=16, C=3, ksz=4, ls=0.05, tf_rate=0.5, np_seed=5)
        sys.exit(0)
    if os.environ.get("ONLY_SSL"):
        case_lm_and_ssl(ref, "ssl_small", seed=21, B=3, Tmax=30, D=16, H=16, subsample=[2, 2], V=11, E=8, A=16,
                        C=3, ksz=4, ls=0.05, JE=8, JH=24)
        sys.exit(0)
    # odd padded extents at every pyramid level, ragged lengths
    case_supervised(ref, "sup_small_odd", seed=11, B=4, Tmax=37, D=24, H=16, n_layers=3, subsample=[2, 2, 2],
                    V=12, E=8, A=16, C=3, ksz=5, ls=0.05)
    # sub==1 first layer (Linear(2H->H) branch, model.py:65), even extents, no label smoothing
    case_supervised(ref, "sup_sub1", seed=12, B=3, Tmax=32, D=20, H=16, n_layers=4, subsample=[1, 2, 2, 2],
                    V=10, E=8, A=24, C=4, ksz=3, ls=0.0)
    # kernel wider than Te, batch 1
    case_supervised(ref, "sup_b1_widekernel", seed=13, B=1, Tmax=19, D=16, H=8, n_layers=2, subsample=[2, 2],
                    V=9, E=8, A=8, C=2, ksz=10, ls=0.05)
    case_host_plumbing("host_plumbing", seed=31)
    case_scheduled(ref, "sched_small", seed=17, B=3, Tmax=33, D=20, H=16, n_layers=2, subsample=[2, 2], V=10, E=8,
                   A=16, C=3, ksz=4, ls=0.05, tf_rate=0.5, np_seed=5)
    case_lm_and_ssl(ref, "ssl_small", seed=21, B=3, Tmax=30, D=16, H=16, subsample=[2, 2], V=11, E=8, A=16,
                    C=3, ksz=4, ls=0.05, JE=8, JH=24)
