"""Array-backed implementations of the three provider interfaces (neural_collaborative_filtering/content_providers.py).

The reference's concrete providers read MovieLens/IMDb frames from `.h5`/`.csv` files that do not ship
(src/content_providers/*.py).  These hold the same information as plain arrays, so the datasets
/ collate contract / models run end to end on synthetic data, and they are where the host-side restatement of the
dynamic collate (a-8) and the device-side graph build (a-7) plug in."""
from __future__ import annotations

import numpy as np
import torch

from . import graph as G
from .neural_collaborative_filtering.content_providers import ContentProvider, DynamicContentProvider, GraphContentProvider


class ArrayProfilesProvider(ContentProvider):
    """Fixed profiles: row k of `item_profiles` / `user_profiles` belongs to the k-th sorted id."""

    def __init__(self, item_ids, item_profiles, user_ids, user_profiles):
        self.item_ids, self.user_ids = np.asarray(item_ids), np.asarray(user_ids)
        self.item_profiles, self.user_profiles = np.asarray(item_profiles), np.asarray(user_profiles)

    def get_item_profile(self, itemID):
        return self.item_profiles[np.searchsorted(self.item_ids, np.asarray(itemID))]

    def get_user_profile(self, userID):
        return self.user_profiles[np.searchsorted(self.user_ids, np.asarray(userID))]

    def get_num_items(self):
        return len(self.item_ids)

    def get_num_users(self):
        return len(self.user_ids)

    def get_item_feature_dim(self):
        return self.item_profiles.shape[1]


class ArrayDynamicProvider(DynamicContentProvider):
    """Item profiles + every user's rating list (sorted by item id), the `user_ratings` frame of the reference as CSR.

    `collate_interacted_items` follows src/content_providers/dynamic_profiles_provider.py:30-73 with numpy index
    arithmetic instead of sklearn's MultiLabelBinarizer + pandas `.loc` (bit-exact, tests/test_providers.py)."""

    def __init__(self, item_ids, item_profiles, user_ids, row_ptr, rated_item_idx, rated_rating):
        self.item_ids, self.user_ids = np.asarray(item_ids), np.asarray(user_ids)
        self.item_profiles = np.asarray(item_profiles)
        self.row_ptr = np.asarray(row_ptr, dtype=np.int64)
        self.rated_item_idx = np.asarray(rated_item_idx, dtype=np.int64)       # index into item_ids, ascending per user
        self.rated_rating = np.asarray(rated_rating, dtype=np.float64)
        cnt = np.diff(self.row_ptr)
        sums = np.add.reduceat(self.rated_rating, self.row_ptr[:-1][cnt > 0]) if self.rated_rating.size else np.zeros(0)
        self.mean_rating = np.zeros(len(cnt))
        self.mean_rating[cnt > 0] = sums / cnt[cnt > 0]

    def get_item_profile(self, itemID):
        return self.item_profiles[np.searchsorted(self.item_ids, np.asarray(itemID))]

    def get_num_items(self):
        return len(self.item_ids)

    def get_num_users(self):
        return len(self.user_ids)

    def get_item_feature_dim(self):
        return self.item_profiles.shape[1]

    def collate_indices(self, user_idx, ignore_ratings=False):
        """(rated item indices (I,), user_matrix (B,I) float32) for a batch of user INDICES."""
        user_idx = np.asarray(user_idx, dtype=np.int64)
        starts, ends = self.row_ptr[user_idx], self.row_ptr[user_idx + 1]
        lens = ends - starts
        flat = np.concatenate([np.arange(s, e) for s, e in zip(starts, ends)]) if len(user_idx) else np.zeros(0, np.int64)
        items = self.rated_item_idx[flat]
        rated = np.unique(items)                                            # sorted unique ids of the batch (:59)
        um = np.zeros((len(user_idx), len(rated)), dtype=np.float64)
        rows = np.repeat(np.arange(len(user_idx)), lens)
        cols = np.searchsorted(rated, items)
        if ignore_ratings:
            um[rows, cols] = 1.0
        else:                                                               # rating - (mean + 2.5) / 2   (:66)
            um[rows, cols] = self.rated_rating[flat] - np.repeat((self.mean_rating[user_idx] + 2.5) / 2, lens)
        return rated, um.astype(np.float32)

    def collate_interacted_items(self, batch, for_ranking: bool, ignore_ratings=False):
        users, cands, third = zip(*batch)
        u_idx = np.searchsorted(self.user_ids, np.asarray(users))
        candidate_items = torch.FloatTensor(self.get_item_profile(cands))
        if for_ranking:
            third = torch.FloatTensor(self.get_item_profile(third))
        else:
            third = torch.FloatTensor(np.asarray(third, dtype=np.float64))
        rated_idx, um = self.collate_indices(u_idx, ignore_ratings)
        rated_items = torch.FloatTensor(self.item_profiles[rated_idx])
        return np.array(cands), self.item_ids[rated_idx], candidate_items, rated_items, torch.from_numpy(um), third


class ArrayGraphProvider(GraphContentProvider):
    """Bipartite graph of an interaction list, built ON THE DEVICE by K4 (graph.create_graph), node ids assigned like
    src/content_providers/graph_providers.py:76-80 from ALL known ids (not just the graph's interactions)."""

    def __init__(self, all_user_ids, all_item_ids, user_ids, item_ids, ratings, item_features, user_features, binary=False,
                 device='cuda'):
        dev = torch.device(device)
        t = lambda a, dt: torch.as_tensor(np.asarray(a), dtype=dt).to(dev)
        self.user_table = G.IdTable(t(all_user_ids, torch.int64))
        self.item_table = G.IdTable(t(all_item_ids, torch.int64))
        self.binary = binary
        self.graph = G.create_graph(t(user_ids, torch.int64), t(item_ids, torch.int64), t(ratings, torch.float64),
                                    torch.as_tensor(item_features, dtype=torch.float32).to(dev),
                                    torch.as_tensor(user_features, dtype=torch.float32).to(dev),
                                    self.user_table, self.item_table, binary=binary)
        self._users = self.user_table.sorted().cpu().numpy()
        self._items = self.item_table.sorted().cpu().numpy()

    def get_num_items(self):
        return self.item_table.count

    def get_num_users(self):
        return self.user_table.count

    def get_item_dim(self):
        return self.graph.item_features.shape[1]

    def get_user_dim(self):
        return self.graph.user_features.shape[1]

    def _rank(self, table, value):
        k = int(np.searchsorted(table, value))
        if k >= len(table) or table[k] != value:
            raise KeyError(value)
        return k

    def get_user_nodeID(self, userID) -> int:
        return self.get_num_items() + self._rank(self._users, userID)

    def get_item_nodeID(self, itemID) -> int:
        return self._rank(self._items, itemID)

    def get_graph(self):
        return self.graph
