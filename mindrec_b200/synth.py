"""Seeded synthetic Criteo-shaped batches (13 dense + 26 sparse fields, Zipf keys).

Layout follows the reference's input contract: ``(feat_ids int32 [B,39], feat_vals float32 [B,39],
label float32 [B,1])`` (models/wide_deep/src/datasets.py:212-216) with the id space of
datasets/criteo_1tb/process_data.py:61-64,117-120: dense columns own ids 0..12 (one id each, the value
is the min-max-scaled weight, missing -> 0), ids 13..38 are the per-column OOV ids, categories follow.
"""
import numpy as np

# per-field cardinalities: the reference's own list (models/wide_deep/src/datasets.py:354-379) ...
CARD_REFERENCE = [691, 540, 20855, 23639, 182, 15, 10091, 347, 4, 16366, 4494, 21293, 3103, 27, 6944,
                  22366, 11, 3267, 1610, 5, 21762, 14, 15, 15030, 61, 12220]
# ... and Criteo-Kaggle (BASELINE config 2: sum + 39 = 33 762 616 rows)
CARD_KAGGLE = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27,
               14992, 5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]
N_DENSE = 13
N_SPARSE = 26
N_FIELDS = N_DENSE + N_SPARSE


def vocab_size(cards):
    return N_FIELDS + int(sum(cards))


def field_offsets(cards):
    off = np.empty(N_SPARSE, dtype=np.int64)
    acc = N_FIELDS
    for i, c in enumerate(cards):
        off[i] = acc
        acc += c
    return off


class CriteoSynth:
    """Deterministic batch generator: seed = base_seed + rank."""

    def __init__(self, batch_size, cards=CARD_KAGGLE, alpha=1.05, seed=20260101, rank=0,
                 vocab_pad=None):
        self.batch_size = batch_size
        self.cards = np.asarray(cards, dtype=np.int64)
        self.offsets = field_offsets(cards)
        self.alpha = alpha
        self.rng = np.random.default_rng(seed + rank)
        self.vocab_size = vocab_pad if vocab_pad is not None else vocab_size(cards)

    def _keys(self, size):
        if self.alpha <= 0.0:
            return self.rng.integers(0, 1 << 62, size=size)
        if self.alpha <= 1.0:
            raise ValueError("Zipf exponent must be > 1 (or 0 for uniform)")
        return self.rng.zipf(self.alpha, size=size) - 1  # rank 0 is the hottest key

    def next(self):
        b = self.batch_size
        ids = np.empty((b, N_FIELDS), dtype=np.int32)
        wts = np.ones((b, N_FIELDS), dtype=np.float32)
        ids[:, :N_DENSE] = np.arange(N_DENSE, dtype=np.int32)[None, :]
        dense = self.rng.random((b, N_DENSE), dtype=np.float32)
        dense[self.rng.random((b, N_DENSE)) < 0.1] = 0.0  # missing values
        wts[:, :N_DENSE] = dense
        keys = self._keys((b, N_SPARSE))
        ids[:, N_DENSE:] = (self.offsets[None, :] + (keys % self.cards[None, :])).astype(np.int32)
        label = (self.rng.random((b, 1)) < 0.25).astype(np.float32)
        return ids, wts, label
