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
        if type == 'sum':
            return negative_ratings / negative_ratings.sum()
        if type == 'sum_dynamic':
            boosted = negative_ratings ** self.w
            return boosted / boosted.sum()
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
