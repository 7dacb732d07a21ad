"""Pin oracle/las_oracle.py against fixtures produced by the reference itself
(tests/golden/make_golden.py ran /root/reference's model.py + the solver.py step bodies)."""
import numpy as np
import pytest
import torch

from oracle import las_oracle as O
from tests.util import cosine, load_golden, rel_err

SUP_CASES = ["sup_small_odd", "sup_sub1", "sup_b1_widekernel"]


@pytest.mark.parametrize("name", SUP_CASES)
@pytest.mark.parametrize("fast", [False, True])
def test_encoder_matches_reference(name, fast):
    G = load_golden(name)
    g = G["raw"]
    x = torch.from_numpy(g["x"])
    enc_h, enc_lens = O.encoder_forward(x, g["ilens"].tolist(), G["p0"], g["subsample"].tolist(), fast=fast)
    assert enc_lens == g["enc_lens"].tolist()                      # integer: bit exact
    assert enc_h.shape == g["enc_h"].shape
    assert rel_err(enc_h, g["enc_h"]) < 2e-5


@pytest.mark.parametrize("name", SUP_CASES)
def test_pyramid_lengths_bit_exact(name):
    G = load_golden(name)
    g = G["raw"]
    _, _, enc_lens, Te = O.pyramid_lengths(g["ilens"].tolist(), g["subsample"].tolist())
    assert enc_lens == g["enc_lens"].tolist()
    assert Te == g["enc_h"].shape[1]


@pytest.mark.parametrize("name", SUP_CASES)
def test_padded_encoder_rows_are_relu_bias(name):
    """rows beyond enc_lens equal relu(project_layers[-1].bias), not zero."""
    G = load_golden(name)
    g = G["raw"]
    n = len(g["subsample"])
    rb = torch.relu(G["p0"][f"encoder.enc2.project_layers.{n - 1}.bias"])
    enc_h, enc_lens = O.encoder_forward(torch.from_numpy(g["x"]), g["ilens"].tolist(), G["p0"], g["subsample"].tolist())
    for b, l in enumerate(enc_lens):
        if l < enc_h.shape[1]:
            assert torch.equal(enc_h[b, l:], rb.expand(enc_h.shape[1] - l, -1))


@pytest.mark.parametrize("name", SUP_CASES)
def test_teacher_forced_decoder_and_loss(name):
    G = load_golden(name)
    g = G["raw"]
    logits, logp, pred, ws = O.e2e_forward(torch.from_numpy(g["x"]), g["ilens"].tolist(), G["p0"],
                                           g["subsample"].tolist(), ys=G["ys"], ls_weight=float(g["ls_weight"]),
                                           labeldist=g["labeldist"], training=True)
    assert rel_err(logits, g["logits"]) < 5e-5
    assert rel_err(logp, g["log_probs"]) < 5e-5
    assert rel_err(ws, g["ws"]) < 5e-5
    assert np.array_equal(pred.numpy(), g["prediction"])
    assert abs(float(-logp.mean()) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    assert abs(float(O.masked_loss(logp, G["ys"])) - float(g["val_loss"])) < 1e-5 * abs(float(g["val_loss"]))
    # attention rows are normalised over ALL Te padded frames
    assert torch.allclose(ws.sum(-1), torch.ones_like(ws.sum(-1)), atol=1e-5)


@pytest.mark.parametrize("name", SUP_CASES)
def test_greedy_decode(name):
    G = load_golden(name)
    g = G["raw"]
    logits, logp, pred, _ = O.e2e_forward(torch.from_numpy(g["x"]), g["ilens"].tolist(), G["p0"],
                                          g["subsample"].tolist(), ys=None, max_dec_timesteps=12,
                                          ls_weight=float(g["ls_weight"]), labeldist=g["labeldist"], training=False)
    assert np.array_equal(pred.numpy(), g["greedy_pred"])
    assert rel_err(logits, g["greedy_logits"]) < 1e-4
    assert rel_err(logp, g["greedy_logp"]) < 1e-4


@pytest.mark.parametrize("name", SUP_CASES)
def test_supervised_step_grads_and_adam(name):
    G = load_golden(name)
    g = G["raw"]
    state = {}
    loss, grads, norm, new = O.supervised_step(torch.from_numpy(g["x"]), g["ilens"].tolist(), G["ys"], G["p0"], state,
                                               g["subsample"].tolist(), float(g["ls_weight"]), g["labeldist"])
    assert abs(loss - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    assert abs(norm - float(g["grad_norm"])) < 1e-4 * float(g["grad_norm"])
    assert len(grads) == len(G["g"])                               # unique tensors only
    for k, gv in G["g"].items():
        assert cosine(grads[k], gv) > 0.99999, k
        assert rel_err(grads[k], gv) < 1e-3, k
    for k, pv in G["p1"].items():
        kk = k[len("decoder."):] if k.startswith("decoder.attention.") else k
        assert torch.allclose(new[kk], pv, rtol=0, atol=2e-6), k


def test_state_dict_has_attention_aliases():
    G = load_golden("sup_small_odd")
    keys = set(G["p0"].keys())
    assert any(k.startswith("attention.") for k in keys) and any(k.startswith("decoder.attention.") for k in keys)
    assert len(O.unique_params(G["p0"])) == len(G["g"])


def test_lm_forward_and_judge_step():
    G = load_golden("ssl_small")
    g = G["raw"]
    logp, probs, preds = O.lm_forward(G["jys"], G["j0"], discrete_input=True, ls_weight=float(g["ls_weight"]),
                                      labeldist=g["labeldist"], training=True)
    assert logp.shape[1] == max(len(y) for y in G["jys"]) + 5       # SURVEY §4 "LM lengths"
    assert rel_err(logp, g["j_logp"]) < 5e-5
    assert rel_err(probs, g["j_probs"]) < 5e-5
    assert np.array_equal(preds.numpy(), g["j_preds"])
    state = {}
    (loss, avg), grads, norm, new = O.judge_step(G["jys"], G["j0"], state, float(g["ls_weight"]), g["labeldist"])
    assert abs(loss - float(g["j_loss"])) < 1e-5 * abs(float(g["j_loss"]))
    assert abs(avg - float(g["j_avg_prob"])) < 1e-5
    assert abs(norm - float(g["j_grad_norm"])) < 1e-4 * float(g["j_grad_norm"])
    for k, gv in G["jg"].items():
        assert cosine(grads[k], gv) > 0.99999, k
    for k, pv in G["j1"].items():
        assert torch.allclose(new[k], pv, rtol=0, atol=2e-6), k


def test_ssl_step():
    G = load_golden("ssl_small")
    g = G["raw"]
    state = {}
    lab = (torch.from_numpy(g["x"]), g["ilens"].tolist(), G["ys"])
    unlab = (torch.from_numpy(g["ux"]), g["uilens"].tolist())
    (loss, sup, unsup), grads, norm, new = O.ssl_step(lab, unlab, G["p0"], G["j0"], state, g["subsample"].tolist(),
                                                      float(g["proportion"]), float(g["ls_weight"]), g["labeldist"],
                                                      g["labeldist"])
    assert abs(sup - float(g["sup"])) < 1e-5 * abs(float(g["sup"]))
    assert abs(unsup - float(g["unsup"])) < 1e-4 * abs(float(g["unsup"]))
    assert abs(loss - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    assert abs(norm - float(g["grad_norm"])) < 1e-4 * float(g["grad_norm"])
    for k, gv in G["g"].items():
        assert cosine(grads[k], gv) > 0.99999, k
    for k, pv in G["p1"].items():
        kk = k[len("decoder."):] if k.startswith("decoder.attention.") else k
        assert torch.allclose(new[kk], pv, rtol=0, atol=2e-6), k


def test_masks_and_targets_bit_exact():
    m = O.seq_mask([3, 1, 0], 4)
    assert m.dtype == np.float32 and m.tolist() == [[1, 1, 1, 0], [1, 0, 0, 0], [0, 0, 0, 0]]
    yi, yo = O.decoder_targets([[5, 6, 7], [8]])
    assert yi.tolist() == [[1, 5, 6, 7], [1, 8, 2, 2]] and yo.tolist() == [[5, 6, 7, 2], [8, 2, 2, 2]]
    yi, yo, lens = O.lm_targets([[5, 6], [7]])
    assert yi.tolist() == [[1, 5, 6, 2, 2, 2, 2], [1, 7, 2, 2, 2, 2, 2]] and lens == [7, 6]
    assert yo.tolist() == [[5, 6, 2, 2, 2, 2, 2], [7, 2, 2, 2, 2, 2, 2]]
    w = O.initial_attention([4, 2], 5)
    assert w[0].tolist() == [0.25, 0.25, 0.25, 0.25, 0.0] and w[1].tolist() == [0.5, 0.5, 0, 0, 0]


def test_scheduled_sampling_matches_reference():
    """tf_rate < 1 (model.py:327-329): with the reference's own per-step draws the oracle reproduces its logits, the
    fed-back predictions, the loss and every gradient (fixture: tests/golden/make_golden.py::case_scheduled)."""
    G = load_golden("sched_small")
    g = G["raw"]
    draws = g["draws"].astype(bool)
    assert not draws[1:].all() and draws[1:].any()
    leaves, full = O._with_grad(G["p0"])
    x, lens, ys = torch.from_numpy(g["x"]), g["ilens"].tolist(), G["ys"]
    logits, logp, pred, ws = O.e2e_forward(x, lens, full, g["subsample"].tolist(), ys=ys, ls_weight=float(g["ls_weight"]),
                                           labeldist=g["labeldist"], training=True, tf_draws=draws)
    assert torch.equal(pred, torch.from_numpy(g["prediction"]))
    assert torch.allclose(logits, torch.from_numpy(g["logits"]), rtol=1e-4, atol=1e-5)
    assert torch.allclose(logp, torch.from_numpy(g["log_probs"]), rtol=1e-4, atol=1e-5)
    loss = -torch.mean(logp)
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    loss.backward()
    for k, gv in G["g"].items():
        if k in leaves:
            assert cosine(leaves[k].grad, gv) > 0.99999, k
    # the teacher-forced forward of the same model differs (the fixture really mixes in model tokens)
    tf_logits = O.e2e_forward(x, lens, full, g["subsample"].tolist(), ys=ys, ls_weight=float(g["ls_weight"]),
                              labeldist=g["labeldist"], training=True)[0]
    assert float((tf_logits - logits).abs().max()) > 1e-3
