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


def test_graph_replay_equals_eager_at_config2_layer_sizes():
    """Same check at BASELINE config-2 layer sizes (H=320, att 320, conv 10x201, V=34; B=8, T=400), where the
    captured step's allocation pattern (buffers freed on the main stream while the weight-gradient side stream
    still reads them) differs from the toy fixtures."""
    M, E, OPT = pkg("model"), pkg("engine"), pkg("optim")
    rng = np.random.RandomState(7)
    B, T, D, V = 8, 400, 249, 34
    lens = sorted([T] + [int(rng.randint(T // 2, T + 1)) for _ in range(B - 1)], reverse=True)
    x = np.zeros((B, T, D), dtype=np.float32)
    ys = []
    for b, l in enumerate(lens):
        x[b, :l] = rng.standard_normal((l, D)).astype(np.float32)
        ys.append(torch.from_numpy(rng.randint(3, V, size=max(2, l // 8)).astype(np.int64)))
    cnt = np.bincount(np.concatenate([y.numpy() for y in ys] + [np.full(B, 2)]), minlength=V).astype(np.float64)
    cnt[:2] = 0
    ld = cnt / cnt.sum()
    batch = (torch.from_numpy(x).pin_memory(), lens, ys)
    traj = {}
    for use_graph in (False, True):
        torch.manual_seed(11)
        m = M.E2E(input_dim=D, enc_hidden_dim=320, enc_n_layers=3, subsample=[2, 2, 2], dropout_rate=0.0, dec_hidden_dim=320,
                  att_dim=320, conv_channels=10, conv_kernel_size=100, att_odim=320, embedding_dim=128, output_dim=V,
                  ls_weight=0.05, labeldist=ld).cuda()
        opt = OPT.FusedAdam(m.parameters(), lr=5e-4, weight_decay=1e-6, amsgrad=True)
        tr = E.SupervisedTrainer(m, opt, max_grad_norm=5.0, use_graph=use_graph)
        out = []
        for _ in range(12):
            loss, norm = tr.step(*batch)
            out.append((float(loss), float(norm)))                   # graph mode returns the SAME static tensors each step
        traj[use_graph] = (np.array([l for l, _ in out]), np.array([n for _, n in out]))
    (la, na), (lb, nb) = traj[False], traj[True]
    assert la[-1] < la[0] - 0.2, la                                  # 12 steps on one batch: it must be learning
    assert np.allclose(la, lb, rtol=5e-3), (la, lb)
    assert np.allclose(na, nb, rtol=0.1), (na, nb)


def test_on_device_input_noise():
    """solver.py:370-373 (add_gaussian): N(0, std) added to the features, here on the device after the upload."""
    G = load_golden("sup_small_odd")
    batch = _batch(G)
    _, _, tr = _trainer(G, use_graph=False)
    key = tr.stage(*batch)
    clean = tr.staged(key).x.clone()
    tr.input_noise_std = 0.5
    key = tr.stage(*batch)
    noise = (tr.staged(key).x - clean).flatten()
    assert abs(float(noise.mean())) < 0.05 and abs(float(noise.std()) - 0.5) < 0.05
    loss, _ = tr.run(key)
    assert bool(torch.isfinite(loss))


def test_geometry_stream_keeps_device_memory_flat():
    """Reference batches almost never repeat a (B, Tmax, Lmax) geometry (shuffle: True, T 30..2300, L 2..250). The
    staging buffers are views of max-extent arenas and captured graphs live in a bounded LRU sharing one pool: 300
    distinct geometries must not grow device memory (the first version allocated ~100 MB per new geometry)."""
    G = load_golden("sup_small_odd")
    g = G["raw"]
    m, opt, tr = _trainer(G, use_graph=True)
    tr.cache.max_graphs = 4
    D = g["x"].shape[2]
    rng = np.random.RandomState(0)
    V = m.decoder.output_layer.bias.numel()
    tr.reserve(4, 64, D, 14)

    def batch(T, L):
        lens = sorted([T] + [int(rng.randint(T // 2 + 1, T + 1)) for _ in range(3)], reverse=True)
        x = np.zeros((4, T, D), dtype=np.float32)
        for b, l in enumerate(lens):
            x[b, :l] = rng.randn(l, D)
        ys = [torch.from_numpy(rng.randint(3, V, size=L if b == 0 else int(rng.randint(1, L + 1))).astype(np.int64)) for b in range(4)]
        return torch.from_numpy(x), lens, ys

    geoms = [(T, L) for T in range(20, 60) for L in range(3, 13)]
    rng.shuffle(geoms)
    geoms = geoms[:300]
    mem = []
    for i, (T, L) in enumerate(geoms):
        loss, _ = tr.step(*batch(T, L))
        if i in (4, 5, 6):                       # a repeated geometry is captured on its second sighting, replayed after
            loss, _ = tr.step(*batch(*geoms[4]))
        if i % 50 == 49:
            torch.cuda.synchronize()
            assert bool(torch.isfinite(loss))
            mem.append(torch.cuda.memory_allocated())
    assert tr.cache.captures >= 1 and len(tr.cache.graphs) <= 4
    assert max(mem[1:]) - mem[1] < 8 << 20, mem    # flat after the first 100 geometries (allocator noise only)
    assert tr.arena.gen == 0                       # reserve() sized the arenas: nothing ever moved
    # a batch larger than anything reserved grows the arenas once and drops the captured graphs
    tr.step(*batch(80, 13))
    assert tr.arena.gen >= 1
    loss, _ = tr.step(*batch(*geoms[4]))
    torch.cuda.synchronize()
    assert bool(torch.isfinite(loss))


def test_learning_rate_change_reaches_a_captured_step():
    """Optimiser hyper-parameters are baked into captured launches by value: a changed lr must invalidate the graph
    (solver.py:519 adjust_learning_rate before ssl_train) instead of being silently ignored on replay."""
    G = load_golden("sup_small_odd")
    batch = _batch(G)
    m, opt, tr = _trainer(G, use_graph=True)
    for _ in range(3):
        tr.step(*batch)                              # eager, capture, replay
    assert tr.cache.captures == 1
    p = next(m.parameters())
    before = p.detach().clone()
    opt.param_groups[0]["lr"] = 0.0
    tr.step(*batch)
    torch.cuda.synchronize()
    assert tr.cache.captures == 2                    # re-captured with the new value
    wd_only = (p.detach() - before).abs().max()
    assert float(wd_only) == 0.0, float(wd_only)     # lr = 0: nothing moves


def test_global_exact_shards_sum_to_the_single_gpu_gradient():
    """SURVEY 8(e) "global-exact" data parallelism, emulated on one GPU: the shards of a batch (dealt as
    data.shard_items deals them), each padded to the global Tmax / Lmax and scaled by 1 / (global B (Lmax + 1)), give
    gradients whose SUM is the gradient of the whole batch, and losses whose sum is its loss. (tools/dp_check.py runs
    the same comparison over NCCL ranks.)"""
    G = load_golden("sup_small_odd")
    g = G["raw"]
    D = pkg("data")
    m, opt, tr = _trainer(G, use_graph=False)
    xs, ilens, ys = _batch(G)
    key = tr.stage(xs, ilens, ys)
    loss_full = float(tr._fwd_bwd(tr.staged(key), key[3]))
    g_full = opt.flat_grad.clone()
    assert abs(loss_full - float(g["loss"])) < 1e-3 * abs(float(g["loss"]))
    items = [(g["x"][b, :ilens[b]], G["ys"][b].tolist()) for b in range(len(ilens))]
    world = 2
    tr.global_exact = True
    tr.global_shape = (key[1], key[3], key[0])                     # global Tmax, Lmax + 1, B
    g_sum, loss_sum = torch.zeros_like(g_full), 0.0
    for r in range(world):
        sx, sl, sy = D.collate(D.shard_items(items, r, world, lambda it: it[0].shape[0]))
        assert len(sl) == 2
        k = tr.stage(sx, sl, sy)
        assert k[1:] == key[1:] and k[0] == 2                       # padded to the global extents
        loss_sum += float(tr._fwd_bwd(tr.staged(k), k[3]))
        g_sum += opt.flat_grad
    assert abs(loss_sum - loss_full) < 1e-4 * abs(loss_full), (loss_sum, loss_full)
    assert cosine(g_sum, g_full) >= 0.9999, cosine(g_sum, g_full)
    assert abs(float(g_sum.norm()) - float(g_full.norm())) < 1e-2 * float(g_full.norm())
    # standard DDP semantics (the default) differ: each shard is padded to its OWN extents and averaged
    tr.global_exact, tr.global_shape = False, None
    k = tr.stage(*D.collate(D.shard_items(items, 1, world, lambda it: it[0].shape[0])))
    assert k[1] <= key[1] and k[3] <= key[3]


def test_two_part_step_equals_the_single_graph_step():
    """The data-parallel trainer runs the step in two parts (everything except encoder layer 0, then layer 0's backward
    as its own autograd graph and CUDA graph) so that the gradient exchange overlaps the last BPTT kernel. On one GPU,
    with no collective issued, the two-part step must follow the one-graph step exactly -- eagerly and as graphs."""
    G = load_golden("sup_small_odd")
    batch = _batch(G)
    traj = {}
    for split in (False, True):
        for use_graph in (False, True):
            m, opt, tr = _trainer(G, use_graph)
            tr.force_split = split
            out = []
            for _ in range(5):
                loss, norm = tr.step(*batch)
                out.append((float(loss), float(norm)))
            traj[(split, use_graph)] = np.array(out)
            if split and use_graph:
                assert tr.cache.captures == 1 and "graph2" in next(iter(tr.cache.graphs.values()))
                early, late = tr._bucket_ranges()
                n = sum(t.numel() for t in early) + sum(t.numel() for t in late)
                assert n == opt.flat_grad.numel() and len(late) == 2          # layer-0 LSTM weights + its projection
    ref = traj[(False, False)]
    assert abs(ref[0, 0] - float(G["raw"]["loss"])) < 1e-3 * abs(float(G["raw"]["loss"]))
    for k, v in traj.items():
        assert np.allclose(v[:, 0], ref[:, 0], rtol=2e-3), (k, v, ref)
        assert np.allclose(v[:, 1], ref[:, 1], rtol=5e-2), (k, v, ref)
