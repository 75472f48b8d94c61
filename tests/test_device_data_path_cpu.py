"""Host side of the device data path without a GPU: what `DeviceCollateProvider` / `ResidentProfilesProvider` /
`DeviceNegativeSampler` upload, the containers the datasets pass through, and a numpy restatement of K6's three phases (bitmap mark ->
popcount rank -> fill, csrc/collate.cu) held to the golden output of the unmodified reference collate."""
import numpy as np
import pandas as pd
import torch

from deeprecommendation_b200.content_providers import (DeviceCollateProvider, DeviceUserMatrix, LazyHostIds, ResidentProfilesProvider,
                                                       ResidentRows)
from deeprecommendation_b200.neural_collaborative_filtering.datasets.base import RankingDataset
from deeprecommendation_b200.neural_collaborative_filtering.datasets.fixed_datasets import FixedPointwiseDataset, FixedRankingDataset
from tests._golden import load


def _k6_numpy(user_rows, list_ptr, list_item, list_val, n_items):
    """the algorithm of csrc/collate.cu, phase by phase"""
    n_words = (n_items + 31) // 32
    bits = np.zeros(n_words, dtype=np.uint32)
    row_nz = np.zeros(len(user_rows), dtype=np.int64)
    for b, u in enumerate(user_rows):                                       # mark
        seg = slice(list_ptr[u], list_ptr[u + 1])
        np.bitwise_or.at(bits, list_item[seg] >> 5, np.uint32(1) << (list_item[seg] & 31).astype(np.uint32))
        row_nz[b] = np.count_nonzero(list_val[seg] != 0.0)
    pop = np.array([bin(int(w)).count('1') for w in bits], dtype=np.int64)  # rank
    word_rank = np.concatenate([[0], np.cumsum(pop)[:-1]]) if n_words else np.zeros(0, np.int64)
    rated = np.array([w * 32 + k for w in range(n_words) for k in range(32) if (int(bits[w]) >> k) & 1], dtype=np.int64)
    row_ptr = np.concatenate([[0], np.cumsum(row_nz)])
    col, val = [], []
    for u in user_rows:                                                     # fill
        for j in range(list_ptr[u], list_ptr[u + 1]):
            if list_val[j] != 0.0:
                it = int(list_item[j])
                below = int(bits[it >> 5]) & ((1 << (it & 31)) - 1)
                col.append(int(word_rank[it >> 5]) + bin(below).count('1'))
                val.append(list_val[j])
    return rated, row_ptr, np.asarray(col, dtype=np.int64), np.asarray(val, dtype=np.float32)


def test_k6_algorithm_reproduces_the_reference_collate():
    d, _, _ = load('collate')
    n_items, n_users = d['profiles'].shape[0], len(d['mean_rating'])
    p = DeviceCollateProvider(np.arange(n_items), d['profiles'], np.arange(n_users), d['row_ptr'], d['rated_idx'], d['rated_rating'], device='cpu')
    # what the provider keeps resident: the CSR of the lists and the centred rating rounded once, as dynamic_profiles_provider.py:66 does
    assert p.d_list_ptr.dtype == torch.int64 and p.d_list_item.dtype == torch.int32 and p.d_list_val.dtype == torch.float32
    rated, row_ptr, col, val = _k6_numpy(d['batch_users'], p.d_list_ptr.numpy(), p.d_list_item.numpy(), p.d_list_val.numpy(), n_items)
    assert np.array_equal(rated, d['rated_items_idx'])
    um = DeviceUserMatrix(torch.from_numpy(row_ptr.astype(np.int32)), torch.from_numpy(col.astype(np.int32)), torch.from_numpy(val),
                          (len(d['batch_users']), len(rated)), int(np.diff(row_ptr).max()))
    assert np.array_equal(um.to_dense().numpy().view(np.int32), d['user_matrix'].view(np.int32))
    assert um.nbytes() == 0
    # the same CSR as the host collate of the resident provider, and its host-known sizes
    rated_h, um_h = p.collate_csr(d['batch_users'])
    assert np.array_equal(rated_h, rated) and np.array_equal(um_h.col.numpy(), col) and np.array_equal(um_h.val.numpy().view(np.int32), val.view(np.int32))
    assert np.array_equal(p._nz_cnt[d['batch_users']], np.diff(row_ptr)) and um_h.max_row_nnz == um.max_row_nnz


def test_zero_centred_ratings_stay_in_the_union_but_leave_the_matrix():
    ptr = np.array([0, 3, 3, 5], dtype=np.int64)
    items = np.array([1, 40, 70, 2, 40], dtype=np.int64)
    ratings = np.array([2.5, 2.5, 2.5, 1.0, 4.0])                           # user 0: mean 2.5 -> every centred rating is exactly 0.0
    p = DeviceCollateProvider(np.arange(100), np.zeros((100, 4), np.float32), np.arange(3), ptr, items, ratings, device='cpu')
    assert p._nz_cnt.tolist() == [0, 0, 2] and p._list_len.tolist() == [3, 0, 2]
    rated, row_ptr, col, val = _k6_numpy(np.array([0, 1, 2, 0]), ptr, items.astype(np.int32), p.d_list_val.numpy(), 100)
    assert rated.tolist() == [1, 2, 40, 70] and row_ptr.tolist() == [0, 0, 0, 2, 2] and col.tolist() == [1, 2]
    assert val.tolist() == [1.0 - (2.5 + 2.5) / 2, 4.0 - (2.5 + 2.5) / 2]


def test_containers_of_the_resident_paths():
    table = torch.arange(40, dtype=torch.float32).view(10, 4)
    rows = ResidentRows(table, np.array([3, 3, 9]))
    assert rows.shape == (3, 4) and len(rows) == 3 and rows.float() is rows and rows.to('cpu') is rows
    assert torch.equal(rows.dense(), table[[3, 3, 9]])
    dev_rows = ResidentRows(table, torch.tensor([1, 2], dtype=torch.int32))          # K6 hands over a tensor
    assert dev_rows.pos.dtype == torch.int64 and torch.equal(dev_rows.dense(), table[[1, 2]])
    ids = LazyHostIds(np.arange(10) * 11, torch.tensor([0, 4, 9]))
    assert ids._host is None and len(ids) == 3
    assert np.asarray(ids).tolist() == [0, 44, 99] and ids[1] == 44 and ids._host is not None


def test_resident_fixed_profiles_pass_through_the_dataset_contract():
    rng = np.random.default_rng(0)
    nI, nU, B = 50, 9, 16
    items, users = rng.random((nI, 6), dtype=np.float32), rng.random((nU, 6), dtype=np.float32)
    item_ids, user_ids = np.arange(nI) * 3 + 5, np.arange(nU) * 7 + 1
    pu, pi, pj = rng.integers(0, nU, B), rng.integers(0, nI, B), rng.integers(0, nI, B)
    prov = ResidentProfilesProvider(item_ids, items, user_ids, users, device='cpu')
    seen = []

    class Probe(torch.nn.Module):
        def forward(self, xu, xi):
            seen.append((xu.dense().numpy(), xi.dense().numpy()))
            return torch.zeros(len(xu), 1)
    ds = FixedPointwiseDataset(pd.DataFrame({'userId': user_ids[pu], 'movieId': item_ids[pi], 'rating': np.ones(B)}), prov)
    batch = next(iter(torch.utils.data.DataLoader(ds, batch_size=B, collate_fn=ds.use_collate())))
    out, y = FixedPointwiseDataset.do_forward(Probe(), batch, 'cpu')
    assert out.shape == (B, 1) and y.shape == (B,)
    assert np.array_equal(seen[0][0], users[pu]) and np.array_equal(seen[0][1], items[pi])
    trip = (prov.get_user_profile(user_ids[pu]), prov.get_item_profile(item_ids[pi]), prov.get_item_profile(item_ids[pj]))
    FixedRankingDataset.do_forward(Probe(), trip, 'cpu')
    assert np.array_equal(seen[1][1], items[pi]) and np.array_equal(seen[2][1], items[pj]) and np.array_equal(seen[2][0], users[pu])


def test_negative_sampler_uploads_the_lists_as_csr():
    frame = pd.DataFrame({'userId': [4, 5, 6], 'positive_movieId': [1, 2, 3],
                          'negative_movieIds': [np.array([10, 11]), np.array([], dtype=np.int64), [7, 8, 9]],
                          'negative_ratings': [np.array([1.0, 2.0]), np.array([]), [0.5, 0.0, 3.5]]})
    ds = RankingDataset(frame)
    s = ds.device_sampler('cpu', seed=3)
    assert s.neg_ptr.tolist() == [0, 2, 2, 5] and s.neg_item.tolist() == [10, 11, 7, 8, 9]
    assert s.neg_rating.dtype == torch.float32 and s.neg_rating.tolist() == [1.0, 2.0, 0.5, 0.0, 3.5]
    assert s.offset == 0 and s.seed == 3
    try:
        s.sample(np.array([0, 2]))                                          # the draw itself is a CUDA kernel: no CPU path
    except RuntimeError as e:
        assert 'CUDA' in str(e)
    else:
        raise AssertionError('sampling on CPU tensors must fail loudly')


def test_k6_algorithm_equals_host_collate_on_random_ragged_batches():
    """property check of the three phases against `collate_csr` (itself bit-exact against the reference golden): random catalogues around the
    32-bit word boundaries of the bitmap, empty lists, exact-zero centred ratings, repeated users"""
    rng = np.random.default_rng(123)
    for case in range(60):
        n_items = int(rng.choice([1, 31, 32, 33, 64, 100, 1000]))
        n_users = int(rng.integers(1, 12))
        ptr, items, ratings = [0], [], []
        for u in range(n_users):
            k = int(rng.integers(0, min(n_items, 40) + 1))
            it = np.sort(rng.choice(n_items, size=k, replace=False))
            r = rng.integers(1, 11, k) * 0.5
            if u % 3 == 1:
                r[:] = 2.5                                                   # mean 2.5: centred ratings exactly 0.0
            items.append(it)
            ratings.append(r)
            ptr.append(ptr[-1] + k)
        items = np.concatenate(items).astype(np.int64) if ptr[-1] else np.zeros(0, np.int64)
        ratings = np.concatenate(ratings).astype(np.float64) if ptr[-1] else np.zeros(0, np.float64)
        p = DeviceCollateProvider(np.arange(n_items), np.zeros((n_items, 2), np.float32), np.arange(n_users), np.asarray(ptr, np.int64), items, ratings,
                                  device='cpu')
        batch = rng.integers(0, n_users, int(rng.integers(1, 20)))
        rated, row_ptr, col, val = _k6_numpy(batch, p.d_list_ptr.numpy(), p.d_list_item.numpy(), p.d_list_val.numpy(), n_items)
        rated_h, um_h = p.collate_csr(batch)
        assert np.array_equal(rated, rated_h), case
        assert np.array_equal(row_ptr, um_h.row_ptr.numpy()) and np.array_equal(col, um_h.col.numpy()), case
        assert np.array_equal(val.view(np.int32), um_h.val.numpy().view(np.int32)), case
