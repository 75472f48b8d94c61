"""The three abstract provider interfaces of the reference (neural_collaborative_filtering/content_providers.py:4-62).

API surface only: concrete MovieLens/IMDb providers (pandas .h5 readers) are out of scope; tests and
benchmarks implement these interfaces over synthetic tensors.
"""
from __future__ import annotations

from abc import ABC


class _Provider(ABC):
    def get_num_items(self) -> int:
        raise NotImplementedError

    def get_num_users(self) -> int:
        raise NotImplementedError


class ContentProvider(_Provider):
    """Fixed user and item profiles (BasicNCF input)."""

    def get_item_profile(self, itemID):
        raise NotImplementedError

    def get_user_profile(self, userID):
        raise NotImplementedError

    def get_item_feature_dim(self) -> int:
        raise NotImplementedError


class DynamicContentProvider(_Provider):
    """Item profiles plus the batch collate that builds AttentionNCF's ragged input.

    `collate_interacted_items(batch, for_ranking)` returns the reference's 6-tuple
    (candidate_ids, rated_ids, candidate_items (B,F), rated_items (I,F), user_matrix (B,I), targets_or_items2)."""

    def get_item_profile(self, itemID):
        raise NotImplementedError

    def get_item_feature_dim(self) -> int:
        raise NotImplementedError

    def collate_interacted_items(self, batch, for_ranking: bool):
        raise NotImplementedError


class GraphContentProvider(_Provider):
    """Node ids and the bipartite graph (GraphNCF input)."""

    def get_user_nodeID(self, userID) -> int:
        raise NotImplementedError

    def get_item_nodeID(self, itemID) -> int:
        raise NotImplementedError

    def get_graph(self):
        raise NotImplementedError
