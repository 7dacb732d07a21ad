"""Host-side batch assembly with the reference's on-disk format and batch layout.

On disk: `{set}.pkl` = pickle of `{utt_id: {'feature': float32[T, D], 'token_ids': list[int]}}` (dataset.py:46-80).
Batch layout (dataloader.py:6-24): utterances sorted by frame count DESCENDING (text-only batches by token count),
features zero-padded to the longest, `ilens` a host list, texts a list of 1-D int64 tensors.

New here (the reference is single-process): a batch can be dealt over the ranks of a data-parallel job. The
globally sorted batch is dealt round-robin, so every shard stays sorted descending and the shards' frame totals
are balanced (the recurrence's latency is set by the longest utterance and by sum(T)).
"""
import pickle

import numpy as np
import torch


class PickleDataset:
    """dataset.py:46-80: whole set in RAM, optional length filters (config keys max/min_feature_length,
    max/min_text_length), keys sorted by frame count ascending when `sort`."""

    def __init__(self, pickle_path, config=None, sort=True):
        with open(pickle_path, "rb") as f:
            self.data_dict = pickle.load(f)
        keys = list(self.data_dict)
        if config:
            lo_f, hi_f = config["min_feature_length"], config["max_feature_length"]
            lo_t, hi_t = config["min_text_length"], config["max_text_length"]
            keys = [k for k in keys
                    if lo_f <= self.data_dict[k]["feature"].shape[0] <= hi_f
                    and lo_t <= len(self.data_dict[k]["token_ids"]) <= hi_t]
        if sort:
            keys = sorted(keys, key=lambda k: self.data_dict[k]["feature"].shape[0])
        self.keys = keys

    def __len__(self):
        return len(self.keys)

    def __getitem__(self, index):
        item = self.data_dict[self.keys[index]]
        return item["feature"], item["token_ids"]


def _pad_features(features, out=None):
    """Zero-padded [B, Tmax, D] batch. `out` (optional): a flat host buffer (e.g. pinned) of at least B*Tmax*D
    elements to build the batch in -- the collation then writes straight into page-locked memory and the H2D copy
    needs no staging pass."""
    T = max(f.shape[0] for f in features)
    B, D = len(features), features[0].shape[1]
    dtype = torch.from_numpy(features[0]).dtype
    if out is None:
        out = torch.zeros(B, T, D, dtype=dtype)
        for b, f in enumerate(features):
            out[b, :f.shape[0]] = torch.from_numpy(f)
        return out
    assert out.dtype == dtype and out.numel() >= B * T * D
    out = out[:B * T * D].view(B, T, D)
    for b, f in enumerate(features):
        n = f.shape[0]
        out[b, :n] = torch.from_numpy(f)
        if n < T:
            out[b, n:].zero_()
    return out


def collate(items, out=None):
    """dataloader.py:6-12 -> (padded [B, Tmax, D], ilens list[int], texts list[LongTensor]); stable sort, descending."""
    items = sorted(items, key=lambda it: it[0].shape[0], reverse=True)
    feats = [f for f, _ in items]
    return _pad_features(feats, out), [f.shape[0] for f in feats], [torch.from_numpy(np.array(t)) for _, t in items]


def speech_collate(items, out=None):
    """dataloader.py:19-24."""
    items = sorted(items, key=lambda it: it[0].shape[0], reverse=True)
    feats = [f for f, _ in items]
    return _pad_features(feats, out), [f.shape[0] for f in feats]


def text_collate(items):
    """dataloader.py:14-17: sorted by token count, descending."""
    items = sorted(items, key=lambda it: len(it[1]), reverse=True)
    return [torch.from_numpy(np.array(t)) for _, t in items]


def shard_items(items, rank, world, key):
    """Deal a batch over `world` ranks: sort descending by `key`, rank r takes positions r, r+world, ..."""
    if world == 1:
        return items
    order = sorted(range(len(items)), key=lambda i: key(items[i]), reverse=True)
    return [items[i] for i in order[rank::world]]


class BatchLoader:
    """Replacement for `DataLoader(dataset, batch_size, shuffle, collate_fn, num_workers=0, drop_last)`
    (dataloader.py:26-38). With world > 1, `batch_size` is the PER-RANK batch: a global batch of
    batch_size * world utterances is drawn (same permutation on every rank: seeded) and dealt by shard_items.

    Every rank takes the SAME number of steps (each step ends in a collective): a tail global batch with fewer
    utterances than ranks -- which would leave the high ranks without a shard -- is dropped on all ranks (at most
    world - 1 utterances per epoch; world == 1 never drops anything, as the reference).

    `bucket=True` (not a reference option): batches are contiguous runs of the dataset order (PickleDataset sorts
    by frame count) and `shuffle` permutes the ORDER OF BATCHES instead of the utterances, so a batch holds
    utterances of similar length: less padding, and few distinct (Tmax, Lmax) geometries for the step graphs.

    `prefetch=n` > 0: batches are collated by a background thread into a ring of page-locked buffers, n batches
    ahead of the consumer (dataloader.py:26-38 uses num_workers=0, i.e. collates inline on the training thread)."""

    def __init__(self, dataset, batch_size, shuffle, drop_last, collate_fn=collate, rank=0, world=1, seed=0,
                 bucket=False, prefetch=0, pin=None):
        self.dataset, self.batch_size, self.shuffle, self.drop_last = dataset, batch_size, shuffle, drop_last
        self.collate_fn, self.rank, self.world = collate_fn, rank, world
        self.rng = np.random.RandomState(seed)
        self.key = (lambda it: len(it[1])) if collate_fn is text_collate else (lambda it: it[0].shape[0])
        self.bucket, self.prefetch = bool(bucket), int(prefetch)
        self.pin = (torch.cuda.is_available() if pin is None else bool(pin)) and collate_fn is not text_collate
        self._ring = []

    def _global_batches(self):
        n = len(self.dataset)
        g = self.batch_size * self.world
        starts = [s for s in range(0, n, g)
                  if not ((min(n, s + g) - s < g and self.drop_last) or min(n, s + g) - s < self.world)]
        return n, g, starts

    def __len__(self):
        return len(self._global_batches()[2])

    def rng_state(self):
        """Shuffle stream of this loader (saved in the resume sidecar)."""
        return self.rng.get_state()

    def set_rng_state(self, state):
        self.rng.set_state(state)

    def _index_batches(self):
        n, g, starts = self._global_batches()
        if self.bucket:
            order = np.arange(n)
            if self.shuffle:
                starts = [starts[i] for i in self.rng.permutation(len(starts))]
        else:
            order = self.rng.permutation(n) if self.shuffle else np.arange(n)
        for s in starts:
            yield order[s:s + g]

    def _host_buffer(self, slot, items):
        """Flat page-locked buffer of ring slot `slot`, grown to hold this batch."""
        T = max(f.shape[0] for f, _ in items)
        need = len(items) * T * items[0][0].shape[1]
        while len(self._ring) <= slot:
            self._ring.append(None)
        buf = self._ring[slot]
        dtype = torch.from_numpy(items[0][0]).dtype
        if buf is None or buf.numel() < need or buf.dtype != dtype:
            buf = self._ring[slot] = torch.empty(max(need, 0 if buf is None else int(buf.numel() * 1.25)), dtype=dtype,
                                                 pin_memory=True)
        return buf

    def _make(self, idx, slot):
        items = shard_items([self.dataset[int(i)] for i in idx], self.rank, self.world, self.key)
        if self.pin and slot is not None:
            return self.collate_fn(items, self._host_buffer(slot, items))
        return self.collate_fn(items)

    def __iter__(self):
        if self.prefetch <= 0:
            for idx in self._index_batches():
                yield self._make(idx, None)
            return
        # Ring depth: the consumer may still be copying the two batches before the one it holds (engine.
        # SupervisedTrainer.steps uploads one ahead and waits for the upload two back before re-using a slot), the
        # queue holds `prefetch` more and the producer writes one: prefetch + 4 slots can never alias.
        import queue
        import threading
        depth = self.prefetch + 4
        q = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()
        batches = list(self._index_batches())       # the shuffle stream is consumed on the caller's thread

        def put(item):
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        def work():
            try:
                for k, idx in enumerate(batches):
                    if not put(("ok", self._make(idx, k % depth))):
                        return
                put(("end", None))
            except BaseException as e:      # noqa: BLE001  (surface loader errors on the training thread)
                put(("err", e))

        th = threading.Thread(target=work, name="las-batch-loader", daemon=True)
        th.start()
        try:
            while True:
                kind, val = q.get()
                if kind == "end":
                    return
                if kind == "err":
                    raise val
                yield val
        finally:
            stop.set()
            th.join(timeout=5.0)
