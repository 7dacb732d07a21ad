"""GPU tests of the train-step engine (engine.SupervisedTrainer + optim.FusedAdam): the step with deferred
weight gradients on a side stream, CUDA-graph capture/replay and the pipelined upload path (`steps()`), against
the reference's own post-step parameters (golden fixtures, solver.py:375-385: loss, backward, clip 5, AMSGrad)."""
import numpy as np
import pytest
import torch

from tests.util import cosine, e2e_from_golden, load_golden, pkg

pytestmark = pytest.mark.gpu
CASES = ["sup_small_odd", "sup_sub1", "sup_b1_widekernel"]


def _trainer(G, use_graph):
    E, OPT = pkg("engine"), pkg("optim")
    m = e2e_from_golden(G)
    opt = OPT.FusedAdam(m.parameters(), lr=5e-4, weight_decay=1e-6, amsgrad=True)
    return m, opt, E.SupervisedTrainer(m, opt, max_grad_norm=5.0, use_graph=use_graph)


def _batch(G):
    g = G["raw"]
    return torch.from_numpy(g["x"]).pin_memory(), g["ilens"].tolist(), [torch.from_numpy(y) for y in G["ys"]]


@pytest.mark.parametrize("name", CASES)
def test_one_step_matches_reference_update(name):
    """loss, pre-clip gradient norm, every gradient and every post-step parameter of ONE optimiser step."""
    G = load_golden(name)
    g = G["raw"]
    m, opt, tr = _trainer(G, use_graph=False)
    loss, norm = tr.step(*_batch(G))
    torch.cuda.synchronize()
    assert abs(float(loss) - float(g["loss"])) < 1e-3 * abs(float(g["loss"]))
    assert abs(float(norm) - float(g["grad_norm"])) < 1e-2 * float(g["grad_norm"])
    named = dict(m.named_parameters())
    ga = torch.cat([named[k].grad.detach().cpu().flatten() for k in G["g"]])
    gb = torch.cat([v.flatten() for v in G["g"].values()])
    assert cosine(ga, gb) >= 0.999                                   # deferred side-stream accumulation landed in .grad
    # Adam's first step moves every weight by ~lr * sign(g): compare the UPDATE, not the parameter
    sd = m.state_dict()
    da = torch.cat([(sd[k].detach().cpu() - G["p0"][k]).flatten() for k in G["p1"]])
    db = torch.cat([(G["p1"][k] - G["p0"][k]).flatten() for k in G["p1"]])
    assert cosine(da, db) >= 0.99, cosine(da, db)
    assert abs(float(da.norm()) - float(db.norm())) < 2e-2 * float(db.norm())


def test_graph_replay_equals_eager():
    """Steps 2.. of a geometry replay a captured graph (fwd + bwd + side-stream wgrads + optimiser): same losses
    as an eager trainer started from the same weights."""
    G = load_golden("sup_small_odd")
    batch = _batch(G)
    losses = {}
    for use_graph in (False, True):
        _, _, tr = _trainer(G, use_graph)
        losses[use_graph] = [float(tr.step(*batch)[0]) for _ in range(5)]
    a, b = np.array(losses[False]), np.array(losses[True])
    assert np.all(np.isfinite(a)) and a[-1] < a[0]                   # it trains
    assert np.allclose(a, b, rtol=2e-3), (a, b)


def test_pipelined_steps_equal_step():
    """`steps()` (H2D of batch i+1 overlapped with step i, double-buffered upload slots) == repeated `step()`,
    over batches of two different geometries."""
    G1, G2 = load_golden("sup_small_odd"), load_golden("sup_small_odd")
    b1 = _batch(G1)
    x2, l2, y2 = _batch(G2)
    rng = np.random.RandomState(3)
    x2 = (x2 * 0.5).pin_memory()
    y2 = [torch.from_numpy(rng.randint(3, 6, size=max(1, len(y) - 1)).astype(np.int64)) for y in y2]   # another Lmax
    seq = [b1, (x2, l2, y2), b1, b1, (x2, l2, y2), (x2, l2, y2), b1]
    _, _, tr_a = _trainer(G1, True)
    ref = [float(tr_a.step(*b)[0]) for b in seq]
    _, _, tr_b = _trainer(G1, True)
    got = [float(loss) for loss, _ in tr_b.steps(seq)]
    assert np.allclose(ref, got, rtol=2e-3), (ref, got)
