"""Dynamic-profile datasets for AttentionNCF (reference: datasets/dynamic_datasets.py).  The batch is the provider's
6-tuple (candidate_ids, rated_ids, candidate_items (B,F), rated_items (I,F), user_matrix (B,I), targets | items2)."""
from __future__ import annotations

from .base import PointwiseDataset, RankingDataset


def _dev(t, device):
    return t.float().to(device)


def _is_resident(x):
    return hasattr(x, 'table') and hasattr(x, 'pos')


def _forward(model, candidate_items, rated_items, user_matrix, device, **kw):
    """dense tensors (the reference's contract) or the device-resident form of `ResidentDynamicProvider`"""
    if _is_resident(candidate_items):
        return model.forward_resident(candidate_items, rated_items, user_matrix, **kw)
    return model(_dev(candidate_items, device), _dev(rated_items, device), _dev(user_matrix, device), **kw)


class DynamicPointwiseDataset(PointwiseDataset):
    def __init__(self, file, dynamic_provider):
        super().__init__(file)
        self.dynamic_provider = dynamic_provider

    def use_collate(self):
        return lambda batch: self.dynamic_provider.collate_interacted_items(batch, for_ranking=False)

    @staticmethod
    def do_forward(model, batch, device, return_attention_weights=False):
        cand_ids, rated_ids, candidate_items, rated_items, user_matrix, y_batch = batch
        res = _forward(model, candidate_items, rated_items, user_matrix, device, return_attention_weights=return_attention_weights)
        if return_attention_weights:
            out, att_weights = res
            return out, y_batch, cand_ids, rated_ids, att_weights, user_matrix
        return res, y_batch


class DynamicRankingDataset(RankingDataset):
    def __init__(self, ranking_file, dynamic_provider):
        super().__init__(ranking_file)
        self.dynamic_provider = dynamic_provider

    def use_collate(self):
        return lambda batch: self.dynamic_provider.collate_interacted_items(batch, for_ranking=True)

    @staticmethod
    def do_forward(model, batch, device):
        _, _, candidate_items1, rated_items, user_matrix, candidate_items2 = batch
        if _is_resident(candidate_items1):
            return (model.forward_resident(candidate_items1, rated_items, user_matrix),
                    model.forward_resident(candidate_items2, rated_items, user_matrix))
        rated, um = _dev(rated_items, device), _dev(user_matrix, device)
        return model(_dev(candidate_items1, device), rated, um), model(_dev(candidate_items2, device), rated, um)
