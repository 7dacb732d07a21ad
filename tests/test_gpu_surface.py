"""GPU parity of the stand-alone single-step surface the reference's modules expose next to the sequence-level
`forward`s: `AttLoc.forward` (model.py:139-173), `Decoder.forward_step` (model.py:283-294), `LM.forward_step` /
`LM.decode` (model.py:535-563) -- against the CPU oracle and against the golden fixtures the reference produced.
Tolerance: ACT_TOL of the tensor maximum (bf16 state operands); alignments (softmax outputs) 2e-3 absolute."""
import numpy as np
import pytest
import torch

from oracle import las_oracle as O
from tests.test_gpu_ssl import lm_from_golden
from tests.test_gpu_supervised import ACT_TOL
from tests.util import e2e_from_golden, load_golden, pkg, rel_err

pytestmark = pytest.mark.gpu


def test_attloc_forward_is_one_attention_read():
    G = load_golden("sup_small_odd")
    g, P = G["raw"], G["p0"]
    m = e2e_from_golden(G).eval()
    att = m.attention
    enc_h = torch.from_numpy(g["enc_h"])
    enc_lens = g["enc_lens"].tolist()
    B, Te, _ = enc_h.shape
    pre = enc_h @ P["attention.mlp_enc.weight"].t() + P["attention.mlp_enc.bias"]
    rng = np.random.RandomState(1)
    dec_z = torch.from_numpy(rng.randn(B, att.decoder_dim).astype(np.float32))
    w0 = torch.from_numpy(O.initial_attention(enc_lens, Te))
    prev = torch.softmax(torch.from_numpy(rng.randn(B, Te).astype(np.float32)), dim=1)
    cases = [(None, None, torch.zeros_like(dec_z), w0), (dec_z, None, dec_z, w0), (dec_z, prev, dec_z, prev)]
    for z_arg, w_arg, z_ref, w_ref in cases:
        att.reset()
        c, w = att(enc_h.cuda(), enc_lens, None if z_arg is None else z_arg.cuda(), None if w_arg is None else w_arg.cuda())
        c_o, w_o = O.attloc_step(enc_h, pre, z_ref, w_ref, P)
        assert tuple(c.shape) == tuple(c_o.shape) and tuple(w.shape) == (B, Te)
        assert float((w.cpu() - w_o).abs().max()) < 2e-3, float((w.cpu() - w_o).abs().max())
        assert abs(float(w.sum(dim=1).min()) - 1.0) < 1e-4          # unmasked softmax over all Te frames
        assert rel_err(c, c_o) < ACT_TOL
    # model.py:141-144: enc_h / mlp_enc(enc_h) stay cached until reset() -- a different enc_pad is ignored until then
    c1, _ = att(torch.zeros_like(enc_h).cuda(), enc_lens, dec_z.cuda(), prev.cuda())
    assert rel_err(c1, c_o) < ACT_TOL
    att.reset()
    c2, _ = att(torch.zeros_like(enc_h).cuda(), enc_lens, dec_z.cuda(), prev.cuda())
    assert rel_err(c2, c_o) > 0.1


def test_decoder_forward_step_chain_reproduces_teacher_forcing():
    """Stepping `forward_step` with the teacher-forced inputs reproduces the reference's `Decoder.forward` logits and
    alignments (golden fixture) step by step."""
    G = load_golden("sup_small_odd")
    g, P = G["raw"], G["p0"]
    m = e2e_from_golden(G).eval()
    dec = m.decoder
    enc_h = torch.from_numpy(g["enc_h"]).cuda()
    enc_lens = g["enc_lens"].tolist()
    ys_in, _ = O.decoder_targets(G["ys"])
    B, L = ys_in.shape
    emb_w = P["decoder.embedding.weight"]
    m.attention.reset()
    dec_z = dec.zero_state(enc_h)
    dec_c = dec.zero_state(enc_h)
    c = dec.zero_state(enc_h, dim=dec.att_odim)
    w = None
    for t in range(L):
        emb = emb_w[torch.from_numpy(ys_in[:, t])].cuda()
        logit, dec_z, dec_c, c, w = dec.forward_step(emb, dec_z, dec_c, c, w, enc_h, enc_lens)
        assert rel_err(logit, g["logits"][:, t]) < ACT_TOL, (t, rel_err(logit, g["logits"][:, t]))
        assert float((w.cpu() - torch.from_numpy(g["ws"][:, t])).abs().max()) < 5e-3, t


def test_decoder_forward_step_dropout_is_training_only():
    G = load_golden("sup_small_odd")
    g = G["raw"]
    m = e2e_from_golden(G, dropout_rate=0.5)
    dec = m.decoder
    enc_h = torch.from_numpy(g["enc_h"]).cuda()
    enc_lens = g["enc_lens"].tolist()
    B = enc_h.size(0)
    rng = np.random.RandomState(2)
    emb = torch.from_numpy(rng.randn(B, dec.embedding.weight.shape[1]).astype(np.float32)).cuda()
    z = torch.from_numpy(rng.randn(B, dec.hidden_dim).astype(np.float32)).cuda() * 0.1
    c = torch.from_numpy(rng.randn(B, dec.att_odim).astype(np.float32)).cuda()
    m.eval()
    a = dec.forward_step(emb, z, z, c, None, enc_h, enc_lens)[0]
    b = dec.forward_step(emb, z, z, c, None, enc_h, enc_lens)[0]
    assert torch.equal(a, b)
    m.train()
    d = dec.forward_step(emb, z, z, c, None, enc_h, enc_lens)[0]
    e = dec.forward_step(emb, z, z, c, None, enc_h, enc_lens)[0]
    assert bool(torch.isfinite(d).all()) and not torch.equal(d, a) and not torch.equal(d, e)


def test_lm_forward_step_chain_and_decode():
    G = load_golden("ssl_small")
    g, J = G["raw"], G["j0"]
    lm = lm_from_golden(G).eval()
    ys = G["jys"]
    ys_in, ys_out, lens = O.lm_targets(ys)
    B, Lm = ys_in.shape
    # oracle logits of the full-length (unpacked) run: feed hypotheses as a dense [B, L] tensor
    emb_w = J["embedding.weight"]
    h = torch.nn.functional.embedding(torch.from_numpy(ys_in), emb_w)
    for l in range(lm.n_layers):
        h = O._uni_lstm_layer(h, [Lm] * B, J[f"LSTM.weight_ih_l{l}"], J[f"LSTM.weight_hh_l{l}"], J[f"LSTM.bias_ih_l{l}"],
                              J[f"LSTM.bias_hh_l{l}"])
    want = h @ J["output_layer.weight"].t() + J["output_layer.bias"]
    dec_z, dec_c = None, None
    for t in range(Lm):
        emb = emb_w[torch.from_numpy(ys_in[:, t])].unsqueeze(1).cuda()
        logit, dec_z, dec_c = lm.forward_step(emb, dec_z, dec_c)
        assert tuple(dec_z.shape) == (lm.n_layers, B, lm.hidden_dim)
        assert rel_err(logit, want[:, t]) < ACT_TOL, (t, rel_err(logit, want[:, t]))
    # greedy decode: every token is the argmax of the oracle's logits for the prefix decoded so far, wherever the
    # oracle's top-2 margin exceeds the bf16 tolerance
    steps = 12
    pred = lm.decode(n_samples=3, sample=False, max_dec_timesteps=steps).cpu()
    assert tuple(pred.shape) == (3, steps) and pred.dtype == torch.int64
    assert bool((pred[0] == pred[1]).all())                           # identical <BOS> starts decode identically
    tin = torch.cat([torch.full((3, 1), 1, dtype=torch.int64), pred[:, :-1]], dim=1)
    h = torch.nn.functional.embedding(tin, emb_w)
    for l in range(lm.n_layers):
        h = O._uni_lstm_layer(h, [steps] * 3, J[f"LSTM.weight_ih_l{l}"], J[f"LSTM.weight_hh_l{l}"],
                              J[f"LSTM.bias_ih_l{l}"], J[f"LSTM.bias_hh_l{l}"])
    lg = h @ J["output_layer.weight"].t() + J["output_layer.bias"]
    top2 = lg.topk(2, dim=-1).values
    safe = (top2[..., 0] - top2[..., 1]) > 0.05 * float(lg.abs().max())
    first_unsafe = [int((~s).nonzero()[0]) if (~s).any() else steps for s in safe]
    for b in range(3):
        n = first_unsafe[b]
        assert torch.equal(pred[b, :n], lg[b, :n].argmax(-1)), b
    sampled = lm.decode(n_samples=4, sample=True, max_dec_timesteps=6)
    assert tuple(sampled.shape) == (4, 6) and int(sampled.min()) >= 0 and int(sampled.max()) < lm.output_dim
