"""CPU tests of the host side: batch layout bit-exact against the reference's own collate functions (golden
fixture made by tests/golden/make_golden.py from /root/reference), label statistics, text post-processing, the
data-parallel dealing of batches (incl. a world_size-2 gloo run), and optimiser checkpoint format."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests.util import GOLDEN_DIR, pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden():
    z = np.load(os.path.join(GOLDEN_DIR, "host_plumbing.npz"))
    g = {k: z[k] for k in z.files}
    items = [(g[f"feat_{i}"], g[f"tok_{i}"].tolist()) for i in range(int(g["n_items"]))]
    vocab = {str(k): int(v) for k, v in zip(g["vocab_keys"], g["vocab_vals"])}
    return g, items, vocab


def test_collate_matches_reference_bit_exact():
    D = pkg("data")
    g, items, _ = _golden()
    padded, ilens, texts = D.collate(list(items))
    assert padded.dtype == torch.float32 and np.array_equal(padded.numpy(), g["c_padded"])     # zero padding, order
    assert ilens == g["c_ilens"].tolist() and ilens == sorted(ilens, reverse=True)
    assert all(t.dtype == torch.int64 and np.array_equal(t.numpy(), g[f"c_text_{i}"]) for i, t in enumerate(texts))
    sp, sil = D.speech_collate(list(items))
    assert np.array_equal(sp.numpy(), g["s_padded"]) and sil == g["s_ilens"].tolist()
    for i, t in enumerate(D.text_collate(list(items))):
        assert np.array_equal(t.numpy(), g[f"t_text_{i}"])


def test_label_statistics_and_text_helpers_match_reference():
    S, U = pkg("solver"), pkg("utils")
    g, items, vocab = _golden()
    fake = type("F", (), {"vocab": vocab, "train_lab_dataset": items})()
    assert np.array_equal(S.Solver.get_label_dist(fake, items), g["labeldist"])
    assert S.Solver.calculate_length_proportion(fake) == float(g["proportion"])
    seqs = [[5, 6, 2, 7, 2], [2, 5], [5, 4, 6, 3, 7], []]
    cut = U.remove_pad_eos(seqs, eos=2)
    assert [len(c) for c in cut] == g["rpe"].tolist()
    assert U.to_sents(cut, vocab, [str(s) for s in g["non_lang"]]) == [str(s) for s in g["sents"]]


def test_edit_distance_and_cer_known_answers():
    U = pkg("utils")
    assert U.edit_distance("kitten", "sitting") == 3 and U.edit_distance("", "abc") == 3 and U.edit_distance("abc", "abc") == 0
    assert U.edit_distance(list("flaw"), list("lawn")) == 2
    assert U.calculate_cer(["abcd", "xy"], ["abed", "xyz"]) == pytest.approx(2 / 7)


def test_seq_mask_and_pad_list_bit_exact():
    U = pkg("utils")
    m = U._seq_mask([3, 1, 0], 4)
    assert m.dtype == torch.float32 and m.tolist() == [[1, 1, 1, 0], [1, 0, 0, 0], [0, 0, 0, 0]]
    p = U.pad_list([torch.tensor([5, 6, 7]), torch.tensor([8])], 2)
    assert p.tolist() == [[5, 6, 7], [8, 2, 2]]


def test_pickle_dataset_filters_and_sorting(tmp_path):
    D = pkg("data")
    rng = np.random.RandomState(0)
    data = {f"u{i}": {"feature": rng.randn(T, 3).astype(np.float32), "token_ids": list(range(3, 3 + L))}
            for i, (T, L) in enumerate([(5, 2), (50, 4), (9, 1), (20, 30), (12, 3)])}
    path = tmp_path / "set.pkl"
    with open(path, "wb") as f:
        pickle.dump(data, f)
    cfg = dict(max_feature_length=40, min_feature_length=6, max_text_length=10, min_text_length=2)
    ds = D.PickleDataset(str(path), config=cfg, sort=True)
    assert [ds[i][0].shape[0] for i in range(len(ds))] == [12]       # only u4 passes all four filters
    ds = D.PickleDataset(str(path), config=None, sort=True)
    assert [ds[i][0].shape[0] for i in range(len(ds))] == [5, 9, 12, 20, 50]
    assert [k for k in D.PickleDataset(str(path), sort=False).keys] == list(data)


def test_batch_dealing_partitions_and_balances():
    D = pkg("data")
    rng = np.random.RandomState(1)
    items = [(rng.randn(int(T), 2).astype(np.float32), [3] * 2) for T in rng.randint(10, 100, size=16)]
    key = lambda it: it[0].shape[0]
    shards = [D.shard_items(items, r, 4, key) for r in range(4)]
    assert sorted(key(it) for s in shards for it in s) == sorted(key(it) for it in items)      # a partition
    for s in shards:
        lens = [key(it) for it in s]
        assert lens == sorted(lens, reverse=True) and len(s) == 4
    tot = [sum(key(it) for it in s) for s in shards]
    assert max(tot) - min(tot) <= max(key(it) for it in items)          # round-robin keeps frame totals close
    # loaders on two ranks draw the same global batches and split them
    ds = items
    la = list(D.BatchLoader(ds, 4, True, False, D.collate, rank=0, world=2, seed=7))
    lb = list(D.BatchLoader(ds, 4, True, False, D.collate, rank=1, world=2, seed=7))
    lg = list(D.BatchLoader(ds, 8, True, False, D.collate, rank=0, world=1, seed=7))
    assert len(la) == len(lb) == len(lg) == 2
    for (xa, ia, _), (xb, ib, _), (xg, ig, _) in zip(la, lb, lg):
        assert sorted(ia + ib, reverse=True) == ig and ia == ig[0::2] and ib == ig[1::2]


def _toy_items(n, seed=3):
    rng = np.random.RandomState(seed)
    return [(rng.randn(int(T), 2).astype(np.float32), list(range(3, 3 + int(T) // 7 + 1))) for T in rng.randint(5, 60, size=n)]


def test_every_rank_takes_the_same_number_of_steps_on_a_short_tail():
    """A tail global batch with fewer utterances than ranks used to leave the high ranks without a shard (they skipped
    the step while the others entered the all-reduce: deadlock). Now it is dropped on every rank."""
    D = pkg("data")
    world, bs = 4, 3
    for n in (12, 13, 14, 15, 16, 25):
        items = _toy_items(n)
        per_rank = [list(D.BatchLoader(items, bs, True, False, D.collate, rank=r, world=world, seed=11)) for r in range(world)]
        steps = [len(b) for b in per_rank]
        assert len(set(steps)) == 1 and steps[0] == len(D.BatchLoader(items, bs, True, False, D.collate, rank=0, world=world)), (n, steps)
        assert all(len(ilens) >= 1 for b in per_rank for _, ilens, _ in b)
        seen = sum(len(ilens) for b in per_rank for _, ilens, _ in b)
        tail = n % (bs * world)
        assert seen == (n - tail if 0 < tail < world else n), (n, seen)
    # world == 1 (the reference's case) never drops anything
    assert sum(len(i) for _, i, _ in D.BatchLoader(_toy_items(13), 4, True, False, D.collate)) == 13


def test_prefetch_thread_yields_the_same_batches():
    D = pkg("data")
    items = _toy_items(23)
    a = list(D.BatchLoader(items, 4, True, False, D.collate, seed=5))
    b = list(D.BatchLoader(items, 4, True, False, D.collate, seed=5, prefetch=2, pin=False))
    assert len(a) == len(b) == 6
    for (xa, ia, ya), (xb, ib, yb) in zip(a, b):
        assert torch.equal(xa, xb) and ia == ib and all(torch.equal(u, v) for u, v in zip(ya, yb))
    # collating into a caller-provided flat buffer (the pinned ring) gives the same padded batch, stale contents cleared
    buf = torch.full((4 * 60 * 2 + 7,), 9.0)
    xs, ilens, _ = D.collate(items[:4], out=buf)
    ref, ilens2, _ = D.collate(items[:4])
    assert torch.equal(xs, ref) and ilens == ilens2 and xs.data_ptr() == buf.data_ptr()
    # an abandoned iterator stops its worker
    it = iter(D.BatchLoader(items, 4, True, False, D.collate, seed=5, prefetch=1, pin=False))
    next(it)
    it.close()

    class Boom:
        def __len__(self):
            return 8

        def __getitem__(self, i):
            raise KeyError("broken item")

    with pytest.raises(KeyError):
        list(D.BatchLoader(Boom(), 4, False, False, D.collate, prefetch=2, pin=False))


def test_bucketed_batches_hold_neighbouring_lengths():
    D = pkg("data")
    items = sorted(_toy_items(40), key=lambda it: it[0].shape[0])
    batches = list(D.BatchLoader(items, 8, True, False, D.collate, seed=1, bucket=True))
    assert len(batches) == 5
    spans = sorted((min(i), max(i)) for _, i, _ in batches)
    assert all(spans[k][1] <= spans[k + 1][0] for k in range(4))          # disjoint length ranges
    assert sorted(l for _, i, _ in batches for l in i) == [it[0].shape[0] for it in items]
    order = [max(i) for _, i, _ in batches]
    assert order != sorted(order)                                         # ... visited in shuffled order


def test_loader_shuffle_stream_round_trips_through_the_sidecar(tmp_path):
    D, U = pkg("data"), pkg("utils")
    items = _toy_items(20)
    ld = D.BatchLoader(items, 4, True, False, D.collate, seed=9)
    list(ld)                                                              # epoch 0 consumed
    U.save_resume_state(str(tmp_path / "m.resume"), {"epoch": 0}, loader_rng={"train_lab_loader": ld.rng_state()})
    want = [i for _, i, _ in ld]                                          # epoch 1 of the uninterrupted run
    ld2 = D.BatchLoader(items, 4, True, False, D.collate, seed=9)
    ld2.set_rng_state(U.load_resume_state(str(tmp_path / "m.resume"), restore_rng=False)["loader_rng"]["train_lab_loader"])
    assert [i for _, i, _ in ld2] == want


GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from tests.util import pkg
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
D = pkg("data")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.RandomState(3)
items = [(rng.randn(int(T), 2).astype(np.float32), [3, 4]) for T in rng.randint(5, 60, size=12)]
tot = torch.zeros(1, dtype=torch.float64)
n = torch.zeros(1, dtype=torch.float64)
for xs, ilens, ys in D.BatchLoader(items, 3, True, False, D.collate, rank=rank, world=world, seed=5):
    assert ilens == sorted(ilens, reverse=True)
    tot += float(xs.double().sum()); n += len(ilens)
dist.all_reduce(tot); dist.all_reduce(n)
ref = sum(float(torch.from_numpy(f).double().sum()) for f, _ in items)
assert int(n) == len(items) and abs(float(tot) - ref) < 1e-9, (float(n), float(tot), ref)
# a dataset whose tail global batch (1 utterance) is smaller than the world: every rank must take the same number of
# steps, each ending in a collective -- this used to hang
items13 = items + [(rng.randn(9, 2).astype(np.float32), [3])]
steps = 0
for xs, ilens, ys in D.BatchLoader(items13, 3, True, False, D.collate, rank=rank, world=world, seed=5, prefetch=1, pin=False):
    t = torch.ones(1); dist.all_reduce(t); assert int(t) == world
    steps += 1
cnt = torch.tensor([float(steps)]); dist.all_reduce(cnt, op=dist.ReduceOp.MAX)
assert steps == int(cnt) == 2, (steps, int(cnt))
# gradient averaging as engine._clip_and_step does it for a stock optimiser
p = torch.nn.Parameter(torch.zeros(4)); p.grad = torch.full((4,), float(rank + 1))
opt = torch.optim.SGD([p], lr=1.0)
E = pkg("engine")
E._clip_and_step(opt, [p], max_grad_norm=1e9)
assert torch.allclose(p.data, torch.full((4,), -(1 + world) / 2.0)), p.data
dist.destroy_process_group()
print("ok", rank)
"""


def test_data_parallel_dealing_under_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_zeros_many_carves_aligned_zeroed_views():
    """functional.zeros_many: one zeroed allocation, disjoint 256-byte-spaced views of the requested shapes/dtypes."""
    Fn = pkg("functional")
    specs = [((3, 5), torch.bfloat16), ((7,), torch.float32), ((2, 2, 2), torch.float32), ((1,), torch.int32)]
    views = Fn.zeros_many("cpu", specs)
    assert [tuple(v.shape) for v in views] == [s for s, _ in specs]
    assert [v.dtype for v in views] == [d for _, d in specs]
    base = views[0].data_ptr()
    offs = [v.data_ptr() - base for v in views]
    assert all(o % 256 == 0 for o in offs) and offs == sorted(offs) and len(set(offs)) == len(offs)
    for i, v in enumerate(views):
        assert not v.any()
        v.fill_(i + 1)                       # writing one view must not leak into another
    for i, v in enumerate(views):
        assert bool((v == i + 1).all())


def test_resume_sidecar_round_trip(tmp_path):
    """utils.save_resume_state / load_resume_state: counters and random streams survive a restart;
    a checkpoint without a sidecar (written by the reference) loads as 'no resume information'."""
    U = pkg("utils")
    path = str(tmp_path / "model.resume")
    assert U.load_resume_state(path) is None
    torch.manual_seed(123)
    np.random.seed(456)
    torch.rand(3), np.random.rand(3)                                  # advance both streams past their seeds
    U.save_resume_state(path, {"phase": "sup_pretrain", "epoch": 4, "best_cer": 0.25}, dropout_seed=987654321)
    want_t, want_n = torch.rand(5), np.random.rand(5)                 # what the run would have drawn next
    torch.manual_seed(1)
    np.random.seed(2)
    state = U.load_resume_state(path)
    assert state["progress"] == {"phase": "sup_pretrain", "epoch": 4, "best_cer": 0.25}
    assert state["dropout_seed"] == 987654321 and state["version"] == 1
    assert torch.equal(torch.rand(5), want_t) and np.array_equal(np.random.rand(5), want_n)
    assert not os.path.exists(path + ".tmp")
    assert U.load_resume_state(path, restore_rng=False)["progress"]["epoch"] == 4
