"""Host-side helpers with the reference's names and semantics (utils.py of
jjery2243542/semi-supervised-ASR): device placement, padding, masks, LSTM re-initialisation,
learning-rate edits. Only what the hot path and its Solver entry points touch."""
import numpy as np
import torch
from torch.nn import init


def cc(net):
    """utils.py:150-152."""
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    return net.to(device)


def to_gpu(data):
    """utils.py:154-158."""
    xs, ilens, ys = data
    return cc(xs), ilens, [cc(y) for y in ys]


def pad_list(xs, pad_value=0):
    """utils.py:173-179: right-pad a list of tensors to the longest."""
    batch_size = len(xs)
    max_length = max(x.size(0) for x in xs)
    pad = xs[0].data.new(batch_size, max_length, *xs[0].size()[1:]).fill_(pad_value)
    for i in range(batch_size):
        pad[i, :xs[i].size(0)] = xs[i]
    return pad


def _seq_mask(seq_len, max_len, is_list=True):
    """utils.py:181-190: mask[b, t] = 1.0 where t < seq_len[b]."""
    if is_list:
        seq_len = torch.from_numpy(np.array(seq_len))
    seq_range = torch.arange(0, max_len).long().to(seq_len.device)
    return (seq_range.unsqueeze(0) < seq_len.unsqueeze(1)).float()


def weight_init(m):
    """utils.py:97-103 as used by LM.__init__ (model.py:470): orthogonal matrices, normal biases."""
    if isinstance(m, (torch.nn.LSTM, torch.nn.LSTMCell, torch.nn.GRU, torch.nn.GRUCell)):
        for param in m.parameters():
            if len(param.shape) >= 2:
                init.orthogonal_(param.data)
            else:
                init.normal_(param.data)


def adjust_learning_rate(optimizer, lr):
    """utils.py:134-139."""
    for param_group in optimizer.param_groups:
        param_group["lr"] = lr
    return lr


def remove_pad_eos(sequences, eos=2):
    """utils.py:192-201: cut every sequence at its first EOS."""
    out = []
    for sequence in sequences:
        try:
            eos_index = next(i for i, v in enumerate(sequence) if v == eos)
        except StopIteration:
            eos_index = len(sequence)
        out.append(sequence[:eos_index])
    return out


def infinite_iter(iterable):
    """utils.py:247-254."""
    it = iter(iterable)
    while True:
        try:
            yield next(it)
        except StopIteration:
            it = iter(iterable)


def edit_distance(a, b):
    """Levenshtein distance between two sequences (the reference delegates to the `editdistance` package,
    utils.py:4, 225)."""
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def ind2character(sequences, non_lang_syms, vocab):
    """utils.py:212-220: indices -> symbols, non-language symbols dropped."""
    inv = {v: k for k, v in vocab.items()}
    skip = {vocab[s] for s in non_lang_syms}
    return [[inv[i] for i in seq if i not in skip] for seq in sequences]


def char_list_to_str(char_lists):
    """utils.py:230-235."""
    return ["".join(" " if ch == "<space>" else ch for ch in chars) for chars in char_lists]


def to_sents(ind_seq, vocab, non_lang_syms):
    """utils.py:203-210."""
    return char_list_to_str(ind2character(ind_seq, non_lang_syms, vocab))


def calculate_cer(hyps, refs):
    """utils.py:222-228: total edit distance / total reference length."""
    dis = sum(edit_distance(h, r) for h, r in zip(hyps, refs))
    return dis / max(1, sum(len(r) for r in refs))


class Logger:
    """utils.py:237-245 writes tensorboard events; here the same calls append JSON lines to `{logdir}/events.jsonl`
    (tensorboardX is not a dependency of the hot path)."""

    def __init__(self, logdir="./log"):
        import json
        import os
        os.makedirs(logdir, exist_ok=True)
        self._f = open(os.path.join(logdir, "events.jsonl"), "a")
        self._json = json

    def scalar_summary(self, tag, value, step):
        self._f.write(self._json.dumps({"tag": tag, "value": float(value), "step": int(step)}) + "\n")
        self._f.flush()

    def text_summary(self, tag, value, step):
        self._f.write(self._json.dumps({"tag": tag, "text": str(value), "step": int(step)}) + "\n")
        self._f.flush()


# ------------------------------------------------------------------------------------------------------------------
# resume sidecar. The reference writes `<path>.ckpt` / `<path>.opt` only (solver.py:38-53), which
# is not enough to continue a run: the epoch / iteration counters, the best CER so far and the random streams (torch,
# numpy, and the device seed the dropout masks are a function of) are lost. They go into `<path>.resume`, a plain
# pickle of python / numpy objects next to the two reference files, which stay byte-compatible with the reference.
# ------------------------------------------------------------------------------------------------------------------
def save_resume_state(path, progress, dropout_seed=None, loader_rng=None):
    """`loader_rng`: {loader name: numpy RandomState state} -- the shuffle streams of the data.BatchLoader objects, so a
    resumed run draws the permutations the uninterrupted run would have drawn next."""
    import pickle
    payload = {"version": 1, "progress": dict(progress), "torch_rng": torch.get_rng_state().numpy().tobytes(),
               "numpy_rng": np.random.get_state(), "dropout_seed": None if dropout_seed is None else int(dropout_seed),
               "loader_rng": dict(loader_rng or {})}
    tmp = f"{path}.tmp"
    with open(tmp, "wb") as f:
        pickle.dump(payload, f)
    import os
    os.replace(tmp, path)          # a crash while writing never leaves a truncated sidecar behind
    return payload


def load_resume_state(path, restore_rng=True):
    """-> payload dict, or None when there is no sidecar (a checkpoint written by the reference)."""
    import os
    import pickle
    if not os.path.exists(path):
        return None
    with open(path, "rb") as f:
        payload = pickle.load(f)
    if restore_rng:
        torch.set_rng_state(torch.frombuffer(bytearray(payload["torch_rng"]), dtype=torch.uint8))
        np.random.set_state(payload["numpy_rng"])
    return payload
