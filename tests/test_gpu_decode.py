"""GPU checks of greedy decoding (Solver.validation / test, solver.py:212-286; model.py:331-348 with ys=None) as ONE
cluster-persistent launch: the in-kernel output layer + argmax + token feedback must give what the per-timestep
kernels give, and what the CPU oracle gives, at the reference's layer sizes -- for a batch, for one utterance
(Solver.test decodes with batch size 1), with and without the in-kernel stop at <EOS>."""
import numpy as np
import pytest
import torch

from oracle import las_oracle as O
from tests.test_gpu_supervised import ACT_TOL, _random_case
from tests.util import pkg, rel_err

pytestmark = pytest.mark.gpu
CFG = dict(seed=41, B=11, T=320, D=249, H=320, sub=[2, 2, 2], V=34, E=128, A=320, C=10, K=100, ls=0.05)


def _sharp_case(**over):
    cfg = dict(CFG, **over)
    m, P, x, lens, ys, _ = _random_case(**cfg)
    with torch.no_grad():          # a freshly initialised output layer gives near-ties at every step: sharpen it
        for k in ("decoder.output_layer.weight", "decoder.output_layer.bias"):
            P[k] = P[k] * 40.0
        m.decoder.output_layer.weight.mul_(40.0)
        m.decoder.output_layer.bias.mul_(40.0)
    return cfg, m.eval(), P, x, lens


def _decode(m, x, lens, steps, persistent):
    Fn, L = pkg("functional"), pkg("_lib")
    old = Fn.DEC_PERSISTENT
    Fn.DEC_PERSISTENT = persistent
    L.path_counters(reset=True)
    try:
        with torch.no_grad():
            logits, logp, pred, ws = m(torch.from_numpy(x).cuda(), lens, ys=None, max_dec_timesteps=steps)
        torch.cuda.synchronize()
    finally:
        Fn.DEC_PERSISTENT = old
    return logits, logp, pred, ws, L.path_counters()


def _agree_until_near_tie(pred, ref_pred, ref_logits, tol=2 * ACT_TOL):
    """Token sequences must agree up to the first step whose reference top-2 margin is within the tolerance."""
    top2 = ref_logits.topk(2, dim=-1).values
    safe = (top2[..., 0] - top2[..., 1]) > tol * float(ref_logits.abs().max())
    n_checked = 0
    for b in range(pred.shape[0]):
        bad = (~safe[b]).nonzero()
        n = int(bad[0]) if len(bad) else pred.shape[1]
        assert torch.equal(pred[b, :n], ref_pred[b, :n]), (b, n, pred[b, :n].tolist(), ref_pred[b, :n].tolist())
        n_checked += n
    return n_checked


@pytest.mark.parametrize("B", [11, 1, 30])     # run-time geometry (2 utterances per cluster) / the two compile-time greedy
def test_persistent_greedy_decode_matches_per_step_kernels_and_oracle(B):      # geometries: 4 (batch of one) and 8 per cluster
    cfg, m, P, x, lens = _sharp_case(B=B)
    steps = 30
    lg_p, logp_p, pred_p, ws_p, cnt_p = _decode(m, x, lens, steps, True)
    lg_s, logp_s, pred_s, ws_s, cnt_s = _decode(m, x, lens, steps, False)
    assert cnt_p["dec_persist_fwd"] == 1 and cnt_p["dec_step_fwd"] == 0, cnt_p     # one launch for all steps
    assert cnt_s["dec_persist_fwd"] == 0 and cnt_s["dec_step_fwd"] == 1, cnt_s
    assert tuple(pred_p.shape) == (B, steps) and pred_p.dtype == torch.int64
    with torch.no_grad():
        o_logits, o_logp, o_pred, o_ws = O.e2e_forward(torch.from_numpy(x), lens, P, cfg["sub"], ys=None, max_dec_timesteps=steps,
                                                       label_smoothing=False, training=False, fast=True)
    # against the per-timestep kernels (same bf16 operands, different summation order): only real near-ties may differ
    n = _agree_until_near_tie(pred_p.cpu(), pred_s.cpu(), lg_s.cpu(), tol=5e-3)
    assert n >= B * steps // 2, n                                            # the comparison is not vacuous
    # against the fp32 oracle: up to the first step whose margin is inside the bf16 activation tolerance
    _agree_until_near_tie(pred_p.cpu(), o_pred, o_logits)
    _agree_until_near_tie(pred_s.cpu(), o_pred, o_logits)
    same = bool((pred_p.cpu() == o_pred).all())
    if same:                                                                 # identical hypotheses: every output comparable
        assert rel_err(lg_p, o_logits) < ACT_TOL
        assert rel_err(logp_p, o_logp) < ACT_TOL
        assert float((ws_p.cpu() - o_ws).abs().max()) < 5e-3
    if bool((pred_p == pred_s).all()):
        assert rel_err(lg_p, lg_s) < ACT_TOL and float((ws_p - ws_s).abs().max()) < 5e-3


def test_persistent_greedy_decode_stops_at_eos_per_cluster():
    """Solver cuts every hypothesis at its first <EOS> (utils.py:192-201): with the stop token set, a cluster ends its loop
    once each of its utterances has emitted one; the cut hypotheses are those of the full-length decode."""
    Fn, U = pkg("functional"), pkg("utils")
    cfg, m, P, x, lens = _sharp_case(B=11)
    with torch.no_grad():
        m.decoder.output_layer.bias[2] += 6.0                               # <EOS> becomes likely after a few steps
    steps = 60
    _, _, full, _, _ = _decode(m, x, lens, steps, True)
    Fn.GREEDY_EARLY_STOP.update(on=True, eos=2)
    try:
        lg, _, early, _, cnt = _decode(m, x, lens, steps, True)
    finally:
        Fn.GREEDY_EARLY_STOP["on"] = False
    assert cnt["dec_persist_fwd"] == 1
    cut = lambda p: U.remove_pad_eos(p.cpu().numpy().tolist(), eos=2)
    assert cut(full) == cut(early)
    if bool((full == 2).any(dim=1).all()):
        last = int((full == 2).float().argmax(dim=1).max())                  # the slowest utterance's <EOS> position
        if last + 2 < steps:
            assert bool((lg[:, last + 2:] == 0).all())                       # nothing was computed after the stop
