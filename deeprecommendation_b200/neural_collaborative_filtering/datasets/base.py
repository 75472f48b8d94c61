"""Dataset templates: the reference's collate / do_forward contract (neural_collaborative_filtering/datasets/base.py).

`samples` may be given as a path prefix (the reference reads `<file>.csv` / `<file>.h5`), a pandas DataFrame, or a
mapping of columns — the reference's data files do not ship, so tests pass frames directly."""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch
from torch import nn
from torch.utils.data import Dataset


def _frame(src, reader):
    if isinstance(src, pd.DataFrame):
        return src
    if isinstance(src, dict):
        return pd.DataFrame(src)
    return reader(src)


def BPR_loss(out_pos, out_neg):
    """Bayesian personalised ranking loss, summed over the batch (datasets/base.py:97-98)."""
    return -torch.nn.functional.logsigmoid(out_pos - out_neg).sum()


class _Hooks:
    """Defaults shared by both templates (datasets/base.py:34-42,86-94)."""

    def get_graph(self, device):
        return None

    def use_collate(self):
        return None

    @staticmethod
    def do_forward(*args, **kwargs):
        raise NotImplementedError


class PointwiseDataset(_Hooks, Dataset):
    """(userId, movieId, rating) triplets; sum-reduced MSE, or BCE-with-logits on rating/5 (datasets/base.py:8-32)."""

    def __init__(self, file, use_bce_loss=False):
        self.samples = _frame(file, lambda f: pd.read_csv(f + '.csv'))
        self.use_bce_loss = use_bce_loss
        self.loss_fn = nn.BCEWithLogitsLoss(reduction='sum') if use_bce_loss else nn.MSELoss(reduction='sum')
        # column views: the reference indexes the frame row by row with .iloc (its dominant epoch cost)
        self._u = self.samples['userId'].to_numpy()
        self._i = self.samples['movieId'].to_numpy()
        self._r = self.samples['rating'].to_numpy()

    def __getitem__(self, item):
        r = self._r[item]
        return self._u[item], self._i[item], (r / 5.0 if self.use_bce_loss else r)

    def __len__(self):
        return len(self._r)

    def calculate_loss(self, y_pred, y_true):
        return self.loss_fn(y_pred, y_true.view(-1, 1).float())


class RankingDataset(_Hooks, Dataset):
    """(userId, positive_movieId, negative_movieIds[], negative_ratings[]) rows; one negative is drawn per access with
    probability ∝ rating^w (`sum_dynamic`, datasets/base.py:57-78); BPR loss."""

    def __init__(self, ranking_file):
        self.samples = _frame(ranking_file, lambda f: pd.read_hdf(f + '.h5'))
        self.loss_fn = BPR_loss
        self.w = 0.0

    def _negative_sampling_probs(self, negative_ratings: np.ndarray, type='sum_dynamic'):
        # (the builtin `sum` — left to right — as in the reference, not numpy's pairwise `.sum()`: the same p to the last bit)
        if type == 'sum':
            return negative_ratings / sum(negative_ratings)
        if type == 'sum_dynamic':
            boosted = negative_ratings ** self.w
            return boosted / sum(boosted)
        if type == 'softmax':
            return torch.softmax(torch.as_tensor(negative_ratings, dtype=torch.float32), dim=0).numpy()
        return None

    def __getitem__(self, item):
        row = self.samples.iloc[item]
        probs = self._negative_sampling_probs(np.asarray(row['negative_ratings'], dtype=np.float64))
        negative = np.random.choice(row['negative_movieIds'], p=probs)
        return row['userId'], row['positive_movieId'], negative

    def __len__(self):
        return len(self.samples)

    def calculate_loss(self, out_pos, out_neg):
        return self.loss_fn(out_pos, out_neg)

    def device_sampler(self, device='cuda', seed=0):
        """The negative lists of every sample resident in HBM + one kernel launch per batch (`DeviceNegativeSampler`) instead of one
        np.random.choice per sample."""
        return DeviceNegativeSampler(self, device, seed)


class DeviceNegativeSampler:
    """Batched form of `RankingDataset.__getitem__` (datasets/base.py:57-78) on the device (K7, csrc/neg_sample.cu).  The frame's
    `negative_movieIds` / `negative_ratings` lists become one CSR in HBM (ids must be integers); `sample(rows)` returns the
    (userId, positive_movieId, negative_movieId) columns of those samples with one negative drawn per sample with probability
    ∝ rating^w (`dataset.w`, read at every call — the training loop anneals it).  Draws are a Philox stream keyed by `seed`; the position
    in the stream advances by the batch size, so an epoch never reuses a uniform."""

    def __init__(self, dataset: 'RankingDataset', device='cuda', seed=0):
        from ... import ops                       # (fails loudly without the CUDA library)
        self._ops, self.dataset, self.seed, self.offset = ops, dataset, int(seed), 0
        frame = dataset.samples
        lens = np.fromiter((len(x) for x in frame['negative_movieIds']), dtype=np.int64, count=len(frame))
        ptr = np.zeros(len(frame) + 1, dtype=np.int64)
        np.cumsum(lens, out=ptr[1:])
        cat = lambda col, dt: (np.concatenate([np.asarray(x, dtype=dt) for x in frame[col]]) if len(frame) and ptr[-1] else np.zeros(0, dt))
        self.neg_ptr = torch.from_numpy(ptr).to(device)
        self.neg_item = torch.from_numpy(cat('negative_movieIds', np.int64)).to(device)
        self.neg_rating = torch.from_numpy(cat('negative_ratings', np.float32)).to(device)
        self._users = frame['userId'].to_numpy()
        self._pos = frame['positive_movieId'].to_numpy()

    def sample(self, rows, return_uniforms=False):
        """rows: sample numbers of the batch (host array).  -> (userIds ndarray, positive ids ndarray, negative ids int64 DEVICE tensor[, list
        positions, uniforms])"""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        dev_rows = torch.from_numpy(rows).to(self.neg_ptr.device)
        res = self._ops.sample_negatives_raw(dev_rows, self.neg_ptr, self.neg_item, self.neg_rating, w=float(self.dataset.w), seed=self.seed,
                                             offset=self.offset, return_uniforms=return_uniforms)
        self.offset += len(rows)
        if return_uniforms:
            return self._users[rows], self._pos[rows], res[0], res[1], res[2]
        return self._users[rows], self._pos[rows], res[0]
