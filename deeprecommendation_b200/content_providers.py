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


class ResidentRows:
    """Rows of a profile table that lives in HBM: what the dense `(n, F)` tensor of the collate contract becomes when the
    provider is device-resident.  `pos` = int64 row numbers (host, pinned when CUDA is present)."""

    def __init__(self, table: torch.Tensor, pos):
        self.table = table
        if torch.is_tensor(pos):                 # already a tensor (K6's `rated` lives on the device)
            self.pos = pos.long()
            return
        t = torch.from_numpy(np.ascontiguousarray(pos, dtype=np.int64))
        self.pos = t.pin_memory() if torch.cuda.is_available() else t

    def __len__(self):
        return int(self.pos.numel())

    @property
    def shape(self):
        return (int(self.pos.numel()), int(self.table.shape[1]))

    def float(self):                     # (the datasets' do_forward calls `.float().to(device)` on whatever the collate returned)
        return self

    def to(self, device, non_blocking=False):
        return self

    def dense(self, device=None):
        """the reference's tensor, materialised (tests / training fallback)"""
        dev = self.table.device if device is None else device
        return self.table.index_select(0, self.pos.to(self.table.device, non_blocking=True)).to(dev)


class OneHotRows:
    """The `(n, num_classes)` one-hot rows of the collate contract (src/content_providers/one_hot_provider.py:17-21,
    fixed_profiles_provider.py:49-50: `one_hot_encode(id, all_ids)`) as their column numbers: 8 bytes per row instead of 4·num_classes.
    Quacks like the tensor the datasets expect (`.float()`, `.to(device)`); models project it with K1s (ops.linear_rows)."""
    kind = 'onehot'

    def __init__(self, ids, num_classes: int):
        self.ids = ids if torch.is_tensor(ids) else torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int64))
        self.num_classes = int(num_classes)

    @property
    def shape(self):
        return (int(self.ids.numel()), self.num_classes)

    def float(self):
        return self

    def to(self, device, non_blocking=False):
        return OneHotRows(self.ids.to(device, non_blocking=non_blocking), self.num_classes)

    def dense(self) -> torch.Tensor:
        out = torch.zeros(self.shape, dtype=torch.float32, device=self.ids.device)
        ok = self.ids >= 0
        out[torch.arange(self.shape[0], device=self.ids.device)[ok], self.ids[ok]] = 1.0
        return out


class MixedRows:
    """Profile rows whose first `n_sparse` columns are multi-hot / sparse (CSR: row_ptr int32, col int32, val fp32 or None = 1) and whose
    remaining columns are dense — the item profiles of the reference: 21 genre + 945 personnel multi-hot columns, then 1,128 dense
    genome-tag relevances.  `dense` may be None (purely multi-hot rows)."""
    kind = 'mixed'

    def __init__(self, row_ptr, col, val, n_sparse: int, dense=None):
        as_t = lambda a, dt: a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a, dtype=dt))
        self.row_ptr, self.col = as_t(row_ptr, np.int32), as_t(col, np.int32)
        self.val = None if val is None else as_t(val, np.float32)
        self.n_sparse = int(n_sparse)
        self.dense = None if dense is None else (dense if torch.is_tensor(dense) else torch.from_numpy(np.ascontiguousarray(dense, dtype=np.float32)))

    @property
    def width(self):
        return self.n_sparse + (0 if self.dense is None else int(self.dense.shape[1]))

    @property
    def shape(self):
        return (int(self.row_ptr.numel() - 1), self.width)

    def float(self):
        return self

    def to(self, device, non_blocking=False):
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)
        return MixedRows(mv(self.row_ptr), mv(self.col), mv(self.val), self.n_sparse, mv(self.dense))

    def dense_rows(self) -> torch.Tensor:
        """the reference's dense (n, width) tensor, materialised (tests)"""
        n = self.shape[0]
        dev = self.row_ptr.device
        out = torch.zeros((n, self.width), dtype=torch.float32, device=dev)
        rows = torch.repeat_interleave(torch.arange(n, device=dev), (self.row_ptr[1:] - self.row_ptr[:-1]).long())
        out[rows, self.col.long()] = 1.0 if self.val is None else self.val
        if self.dense is not None:
            out[:, self.n_sparse:] = self.dense
        return out

    @classmethod
    def from_dense(cls, rows: np.ndarray, n_sparse: int):
        """host-side split of dense profile rows (what a provider does once for its whole table)"""
        rows = np.asarray(rows, dtype=np.float32)
        sp = rows[:, :n_sparse]
        r, c = np.nonzero(sp)
        row_ptr = np.zeros(rows.shape[0] + 1, dtype=np.int32)
        np.cumsum(np.bincount(r, minlength=rows.shape[0]), out=row_ptr[1:])
        v = sp[r, c]
        return cls(row_ptr, c.astype(np.int32), None if np.all(v == 1.0) else v, n_sparse, rows[:, n_sparse:] if n_sparse < rows.shape[1] else None)

    def take(self, idx: np.ndarray):
        """rows `idx` (host arrays) as a new MixedRows — the per-batch `.loc` lookup of a provider"""
        rp, col = self.row_ptr.numpy(), self.col.numpy()
        idx = np.asarray(idx, dtype=np.int64)
        lens = rp[idx + 1] - rp[idx]
        nrp = np.zeros(len(idx) + 1, dtype=np.int32)
        np.cumsum(lens, out=nrp[1:])
        src = np.repeat(rp[idx], lens) + (np.arange(int(nrp[-1])) - np.repeat(nrp[:-1], lens))
        return MixedRows(nrp, col[src], None if self.val is None else self.val.numpy()[src], self.n_sparse,
                         None if self.dense is None else self.dense[torch.from_numpy(idx)])


class ResidentProfilesProvider(ArrayProfilesProvider):
    """Fixed profiles (src/content_providers/fixed_profiles_provider.py) with BOTH tables resident in HBM: `get_user_profile` /
    `get_item_profile` hand out `ResidentRows` (8 B per row instead of 4·F), which `FixedPointwiseDataset` / `FixedRankingDataset` pass
    through to `BasicNCF.forward` unchanged.  At config 1 a 512-pair batch is 8 KB of row numbers instead of 8.6 MB."""

    def __init__(self, item_ids, item_profiles, user_ids, user_profiles, device='cuda'):
        super().__init__(item_ids, item_profiles, user_ids, user_profiles)
        self.item_table = torch.as_tensor(self.item_profiles, dtype=torch.float32).to(device)
        self.user_table = torch.as_tensor(self.user_profiles, dtype=torch.float32).to(device)

    def get_item_profile(self, itemID):
        return ResidentRows(self.item_table, np.searchsorted(self.item_ids, np.atleast_1d(np.asarray(itemID))))

    def get_user_profile(self, userID):
        return ResidentRows(self.user_table, np.searchsorted(self.user_ids, np.atleast_1d(np.asarray(userID))))


class OneHotArrayProvider(ContentProvider):
    """One-hot user AND item profiles (src/content_providers/one_hot_provider.py).  sparse=True hands out `OneHotRows` (ids), sparse=False
    the reference's dense one-hot rows."""

    def __init__(self, item_ids, user_ids, sparse=True):
        self.item_ids, self.user_ids, self.sparse = np.asarray(item_ids), np.asarray(user_ids), sparse

    def _rows(self, table, ids):
        pos = np.searchsorted(table, np.atleast_1d(np.asarray(ids)))
        if self.sparse:
            return OneHotRows(pos, len(table))
        out = np.zeros((len(pos), len(table)), dtype=np.float32)
        out[np.arange(len(pos)), pos] = 1.0
        return out

    def get_item_profile(self, itemID):
        return self._rows(self.item_ids, itemID)

    def get_user_profile(self, userID):
        return self._rows(self.user_ids, userID)

    def get_num_items(self):
        return len(self.item_ids)

    def get_num_users(self):
        return len(self.user_ids)

    def get_item_feature_dim(self):
        return len(self.item_ids)


class MixedProfilesProvider(ArrayProfilesProvider):
    """Fixed profiles whose item rows are handed out as `MixedRows` (multi-hot columns as index lists, dense columns as is)."""

    def __init__(self, item_ids, item_profiles, user_ids, user_profiles, n_sparse):
        super().__init__(item_ids, item_profiles, user_ids, user_profiles)
        self._mixed = MixedRows.from_dense(self.item_profiles, n_sparse)

    def get_item_profile(self, itemID):
        return self._mixed.take(np.searchsorted(self.item_ids, np.atleast_1d(np.asarray(itemID))))


class SparseUserMatrix:
    """CSR of the `(B, I)` user_matrix of the collate contract (exact zeros = "unrated" are simply absent)."""

    def __init__(self, row_ptr, col, val, shape):
        pin = (lambda t: t.pin_memory()) if torch.cuda.is_available() else (lambda t: t)
        self.row_ptr = pin(torch.from_numpy(np.ascontiguousarray(row_ptr, dtype=np.int32)))
        self.col = pin(torch.from_numpy(np.ascontiguousarray(col, dtype=np.int32)))
        self.val = pin(torch.from_numpy(np.ascontiguousarray(val, dtype=np.float32)))
        self.shape = tuple(shape)
        self.max_row_nnz = int(np.max(np.diff(np.asarray(row_ptr)))) if len(row_ptr) > 1 else 0     # known on the host: sizes K2's segment grid

    def to_dense(self) -> torch.Tensor:
        um = torch.zeros(self.shape, dtype=torch.float32)
        rows = torch.repeat_interleave(torch.arange(self.shape[0]), (self.row_ptr[1:] - self.row_ptr[:-1]).long())
        um[rows, self.col.long()] = self.val
        return um

    def nbytes(self) -> int:
        return int(self.row_ptr.numel() * 4 + self.col.numel() * 8)


class ResidentDynamicProvider(ArrayDynamicProvider):
    """`ArrayDynamicProvider` with the item-profile table resident in HBM (row f-4 of SURVEY.md §8: device-side data path).

    The reference's collate copies `item_profiles[rated_items_ids]` — 79 MB per batch at config 2 — to the host tensor that
    `do_forward` then ships over PCIe (dynamic_profiles_provider.py:70, dynamic_datasets.py:25-40).  Here the 6-tuple keeps
    its shape but carries row numbers (`ResidentRows`) and the user matrix as CSR (`SparseUserMatrix`): ~2.4 MB per batch.
    `DynamicPointwiseDataset.do_forward` recognises them and calls `AttentionNCF.forward_resident`; results are identical
    to the dense path (tests/test_providers.py, tests/test_models_gpu.py)."""

    def __init__(self, item_ids, item_profiles, user_ids, row_ptr, rated_item_idx, rated_rating, device='cuda'):
        super().__init__(item_ids, item_profiles, user_ids, row_ptr, rated_item_idx, rated_rating)
        self.table = torch.as_tensor(self.item_profiles, dtype=torch.float32).to(device)

    def collate_csr(self, user_idx, ignore_ratings=False):
        """(rated item indices (I,), CSR of user_matrix) — same arithmetic as `collate_indices`, without the dense matrix"""
        user_idx = np.asarray(user_idx, dtype=np.int64)
        starts, ends = self.row_ptr[user_idx], self.row_ptr[user_idx + 1]
        lens = ends - starts
        ptr = np.zeros(len(user_idx) + 1, dtype=np.int64)
        np.cumsum(lens, out=ptr[1:])
        flat = np.repeat(starts - ptr[:-1], lens) + np.arange(ptr[-1])
        items = self.rated_item_idx[flat]
        rated = np.unique(items)
        cols = np.searchsorted(rated, items)
        if ignore_ratings:
            vals = np.ones(len(flat), dtype=np.float32)
        else:
            vals = (self.rated_rating[flat] - np.repeat((self.mean_rating[user_idx] + 2.5) / 2, lens)).astype(np.float32)
        keep = vals != 0.0                      # an exactly-zero centred rating is "unrated" in the dense form too
        if not keep.all():
            rows = np.repeat(np.arange(len(user_idx)), lens)[keep]
            cols, vals = cols[keep], vals[keep]
            ptr = np.zeros(len(user_idx) + 1, dtype=np.int64)
            np.cumsum(np.bincount(rows, minlength=len(user_idx)), out=ptr[1:])
        return rated, SparseUserMatrix(ptr, cols, vals, (len(user_idx), len(rated)))

    def collate_interacted_items(self, batch, for_ranking: bool, ignore_ratings=False):
        users, cands, third = zip(*batch)
        u_idx = np.searchsorted(self.user_ids, np.asarray(users))
        candidate_items = ResidentRows(self.table, np.searchsorted(self.item_ids, np.asarray(cands)))
        if for_ranking:
            third = ResidentRows(self.table, np.searchsorted(self.item_ids, np.asarray(third)))
        else:
            third = torch.FloatTensor(np.asarray(third, dtype=np.float64))
        rated_idx, um = self.collate_csr(u_idx, ignore_ratings)
        return np.array(cands), self.item_ids[rated_idx], candidate_items, ResidentRows(self.table, rated_idx), um, third


class DeviceUserMatrix:
    """CSR of the `(B, I)` user_matrix in DEVICE memory — what K6 (ops.collate_interacted_raw) leaves behind.  Same duck type as
    `SparseUserMatrix` for `AttentionNCF.forward_resident` (`row_ptr`, `col`, `val`, `shape`, `max_row_nnz`)."""

    def __init__(self, row_ptr, col, val, shape, max_row_nnz):
        self.row_ptr, self.col, self.val = row_ptr, col, val
        self.shape = tuple(shape)
        self.max_row_nnz = int(max_row_nnz)

    def to_dense(self) -> torch.Tensor:
        """the reference's dense matrix, on the CSR's device (training path, tests)"""
        um = torch.zeros(self.shape, dtype=torch.float32, device=self.val.device)
        rows = torch.repeat_interleave(torch.arange(self.shape[0], device=self.val.device), (self.row_ptr[1:] - self.row_ptr[:-1]).long())
        um[rows, self.col.long()] = self.val
        return um

    def nbytes(self) -> int:
        return 0                                # nothing crosses PCIe


class LazyHostIds:
    """`rated_items_ids` of the 6-tuple (dynamic_profiles_provider.py:59,73) for a batch collated on the device: the ids are
    downloaded when somebody looks at them (the training / evaluation loops do not; the attention-weight visualisation does)."""

    def __init__(self, item_ids: np.ndarray, rated_rows: torch.Tensor):
        self._item_ids, self._rows, self._host = item_ids, rated_rows, None

    def _get(self):
        if self._host is None:
            self._host = self._item_ids[self._rows.cpu().numpy()]
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self._get()
        return a if dtype is None else a.astype(dtype)

    def __len__(self):
        return int(self._rows.numel())

    def __getitem__(self, k):
        return self._get()[k]


class DeviceCollateProvider(ResidentDynamicProvider):
    """`ResidentDynamicProvider` whose collate itself runs on the device (K6, csrc/collate.cu): besides the
    profile table, every user's rating list lives in HBM (item numbers + the centred rating `rating - (meanRating + 2.5) / 2`, float64
    arithmetic rounded once to fp32 like dynamic_profiles_provider.py:66).  Per batch the host touches O(B) numbers: the user rows go
    up (8 B each), the size of the rated-item union comes back (4 B); sort/unique/multi-hot of the reference's collate are three
    launches of integer work.  Results are bit-identical to `ResidentDynamicProvider.collate_csr` (tests/test_collate_gpu.py)."""

    def __init__(self, item_ids, item_profiles, user_ids, row_ptr, rated_item_idx, rated_rating, device='cuda'):
        super().__init__(item_ids, item_profiles, user_ids, row_ptr, rated_item_idx, rated_rating, device=device)
        dev = self.table.device
        self._list_len = np.diff(self.row_ptr)
        centred = (self.rated_rating - np.repeat((self.mean_rating + 2.5) / 2, self._list_len)).astype(np.float32)
        rows_of = np.repeat(np.arange(len(self._list_len)), self._list_len)
        self._nz_cnt = np.bincount(rows_of[centred != 0.0], minlength=len(self._list_len)).astype(np.int64)
        self.d_list_ptr = torch.from_numpy(self.row_ptr).to(dev)
        self.d_list_item = torch.from_numpy(self.rated_item_idx.astype(np.int32)).to(dev)
        self.d_list_val = torch.from_numpy(centred).to(dev)
        self._count_ring, self._ring_pos = None, 0             # pinned landing slots of K6's counts (at most 8 collates in flight)

    def collate_launch(self, user_idx, ignore_ratings=False):
        """First half of `collate_device`: uploads the user rows, enqueues K6 and the 8-byte copy of its counts on the current stream, records
        an event and returns a handle WITHOUT synchronising — a loader can enqueue the collate of batch k + 1 ahead of the forward of batch
        k and learn I while that forward runs (bench.py `e2e_device_collate`)."""
        from . import ops
        dev = self.table.device
        if torch.is_tensor(user_idx):
            rows = user_idx.to(dev, non_blocking=True)
            user_idx = user_idx.numpy()
        else:
            user_idx = np.ascontiguousarray(user_idx, dtype=np.int64)
            rows = torch.from_numpy(user_idx).to(dev)
        B = len(user_idx)
        lens = self._list_len[user_idx]
        cnt = lens if ignore_ratings else self._nz_cnt[user_idx]
        nnz, total = int(cnt.sum()), int(lens.sum())
        rated, rp, col, val, counts = ops.collate_interacted_raw(rows, self.d_list_ptr, self.d_list_item, None if ignore_ratings else self.d_list_val,
                                                                 self.get_num_items(), rated_capacity=min(self.get_num_items(), total),
                                                                 nnz_capacity=nnz)
        if self._count_ring is None:
            self._count_ring = [torch.empty(2, dtype=torch.int32).pin_memory() for _ in range(8)]
        host_counts = self._count_ring[self._ring_pos % len(self._count_ring)]
        self._ring_pos += 1
        host_counts.copy_(counts, non_blocking=True)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(dev))
        return (rated, rp, col, val, host_counts, done, B, nnz, int(cnt.max()) if B else 0)

    def collate_finish(self, pending):
        """Second half: waits for K6 of that batch only (not for work enqueued after it) and returns (rated rows (I,), `DeviceUserMatrix`)."""
        rated, rp, col, val, host_counts, done, B, nnz, max_row = pending
        done.synchronize()
        n_rated = int(host_counts[0])            # the one thing the host has to learn: I sizes the projection GEMM of the rated rows
        return rated[:n_rated], DeviceUserMatrix(rp, col[:nnz], val[:nnz], (B, n_rated), max_row)

    def collate_device(self, user_idx, ignore_ratings=False):
        """(rated item rows (I,) int64 DEVICE tensor, `DeviceUserMatrix`) for a batch of user INDICES: a host array, or an int64 host
        tensor (pinned memory makes the upload asynchronous)."""
        return self.collate_finish(self.collate_launch(user_idx, ignore_ratings))

    def collate_interacted_items(self, batch, for_ranking: bool, ignore_ratings=False):
        users, cands, third = zip(*batch)
        u_idx = np.searchsorted(self.user_ids, np.asarray(users))
        candidate_items = ResidentRows(self.table, np.searchsorted(self.item_ids, np.asarray(cands)))
        if for_ranking:
            third = ResidentRows(self.table, np.searchsorted(self.item_ids, np.asarray(third)))
        else:
            third = torch.FloatTensor(np.asarray(third, dtype=np.float64))
        rated_rows, um = self.collate_device(u_idx, ignore_ratings)
        return np.array(cands), LazyHostIds(self.item_ids, rated_rows), candidate_items, ResidentRows(self.table, rated_rows), um, third


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
