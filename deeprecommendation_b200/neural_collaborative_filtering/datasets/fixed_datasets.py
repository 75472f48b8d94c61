"""Fixed-profile datasets for BasicNCF (reference: datasets/fixed_datasets.py)."""
from __future__ import annotations

import torch

from .base import PointwiseDataset, RankingDataset


def _tensor(p):
    """dense rows become the reference's FloatTensor; sparse row containers (content_providers.OneHotRows / MixedRows) and rows of a
    device-resident table (content_providers.ResidentRows) pass through — they
    answer `.float()` / `.to(device)` like a tensor and the models project them with the gather-sum kernel"""
    return p if (getattr(p, 'kind', None) in ('onehot', 'mixed') or hasattr(p, 'table')) else torch.FloatTensor(p)


def _profiles(cp, users, *item_lists):
    out = [_tensor(cp.get_user_profile(userID=users))]
    out += [_tensor(cp.get_item_profile(itemID=ids)) for ids in item_lists]
    return out


class FixedPointwiseDataset(PointwiseDataset):
    def __init__(self, file, content_provider):
        super().__init__(file)
        self.content_provider = content_provider

    def use_collate(self):
        cp = self.content_provider

        def collate(batch):          # (user_vecs (B,Fu), item_vecs (B,Fi), targets (B,))  fixed_datasets.py:14-20
            users, items, targets = zip(*batch)
            user_vecs, item_vecs = _profiles(cp, users, items)
            return user_vecs, item_vecs, torch.FloatTensor(targets)
        return collate

    @staticmethod
    def do_forward(model, batch, device):
        user_vec, item_vec, y_batch = batch
        return model(user_vec.float().to(device), item_vec.float().to(device)), y_batch


class FixedRankingDataset(RankingDataset):
    def __init__(self, ranking_file, content_provider):
        super().__init__(ranking_file)
        self.content_provider = content_provider

    def use_collate(self):
        cp = self.content_provider

        def collate(batch):          # (user_vecs, positive item_vecs, negative item_vecs)  fixed_datasets.py:39-46
            users, pos, neg = zip(*batch)
            return tuple(_profiles(cp, users, pos, neg))
        return collate

    @staticmethod
    def do_forward(model, batch, device):
        user_vec, item1_vec, item2_vec = batch
        u = user_vec.float().to(device)
        return model(u, item1_vec.float().to(device)), model(u, item2_vec.float().to(device))
