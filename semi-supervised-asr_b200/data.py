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


def _pad_features(features):
    T = max(f.shape[0] for f in features)
    out = torch.zeros(len(features), T, features[0].shape[1], dtype=torch.from_numpy(features[0]).dtype)
    for b, f in enumerate(features):
        out[b, :f.shape[0]] = torch.from_numpy(f)
    return out


def collate(items):
    """dataloader.py:6-12 -> (padded [B, Tmax, D], ilens list[int], texts list[LongTensor]); stable sort, descending."""
    items = sorted(items, key=lambda it: it[0].shape[0], reverse=True)
    feats = [f for f, _ in items]
    return _pad_features(feats), [f.shape[0] for f in feats], [torch.from_numpy(np.array(t)) for _, t in items]


def speech_collate(items):
    """dataloader.py:19-24."""
    items = sorted(items, key=lambda it: it[0].shape[0], reverse=True)
    feats = [f for f, _ in items]
    return _pad_features(feats), [f.shape[0] for f in feats]


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
    batch_size * world utterances is drawn (same permutation on every rank: seeded) and dealt by shard_items."""

    def __init__(self, dataset, batch_size, shuffle, drop_last, collate_fn=collate, rank=0, world=1, seed=0):
        self.dataset, self.batch_size, self.shuffle, self.drop_last = dataset, batch_size, shuffle, drop_last
        self.collate_fn, self.rank, self.world = collate_fn, rank, world
        self.rng = np.random.RandomState(seed)
        self.key = (lambda it: len(it[1])) if collate_fn is text_collate else (lambda it: it[0].shape[0])

    def __len__(self):
        g = self.batch_size * self.world
        n = len(self.dataset)
        return n // g if self.drop_last else (n + g - 1) // g

    def __iter__(self):
        n = len(self.dataset)
        order = self.rng.permutation(n) if self.shuffle else np.arange(n)
        g = self.batch_size * self.world
        for s in range(0, n, g):
            idx = order[s:s + g]
            if len(idx) < g and self.drop_last:
                return
            items = shard_items([self.dataset[int(i)] for i in idx], self.rank, self.world, self.key)
            if items:
                yield self.collate_fn(items)
