"""K6 (csrc/collate.cu): the dynamic-profile collate on the device, BIT-EXACT (integer / index work) against
  * the golden output of the unmodified reference collate (src/content_providers/dynamic_profiles_provider.py:30-73, tests/golden/collate.npz),
  * the host restatement `ResidentDynamicProvider.collate_csr` on seeded ragged batches (repeated users, users without ratings, centred
    ratings that are exactly 0.0, ignore_ratings, catalogues on both sides of the shared-memory bitmap limit, scans longer than one pass),
and the scores of AttentionNCF fed by it against the host-collated path (identical bits)."""
import numpy as np
import pytest
import torch

from oracle import synth
from tests._golden import load

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _lists(n_users, n_items, mean_len, seed, empty_every=7, flat_every=5):
    """CSR of rating lists: ascending item numbers per user; every `empty_every`-th user has no ratings, every `flat_every`-th rates
    everything 2.5 (mean 2.5 -> every centred rating is exactly 0.0), one user rates 1.0 / 4.0 alternately (mean 2.5 -> non-zero)."""
    rng = np.random.default_rng(seed)
    ptr, idx, rr = [0], [], []
    for u in range(n_users):
        n = 0 if (empty_every and u % empty_every == 3) else int(min(n_items, max(1, rng.poisson(mean_len))))
        items = np.sort(rng.choice(n_items, size=n, replace=False))
        r = rng.integers(1, 11, n) * 0.5
        if flat_every and u % flat_every == 2:
            r[:] = 2.5
        idx.append(items)
        rr.append(r)
        ptr.append(ptr[-1] + n)
    return np.asarray(ptr, np.int64), np.concatenate(idx).astype(np.int64), np.concatenate(rr).astype(np.float64)


def _providers(n_users, n_items, mean_len, seed, F=8):
    from deeprecommendation_b200.content_providers import DeviceCollateProvider
    ptr, idx, rr = _lists(n_users, n_items, mean_len, seed)
    profiles = np.random.default_rng(seed + 1).random((n_items, F), dtype=np.float32)
    return DeviceCollateProvider(np.arange(n_items), profiles, np.arange(n_users), ptr, idx, rr, device=DEV)


def _check_against_host(p, user_idx, ignore_ratings=False):
    rated_h, um_h = p.collate_csr(user_idx, ignore_ratings)
    rated_d, um_d = p.collate_device(user_idx, ignore_ratings)
    assert rated_d.dtype == torch.int64 and np.array_equal(rated_d.cpu().numpy(), rated_h)
    assert um_d.shape == um_h.shape and um_d.max_row_nnz == um_h.max_row_nnz
    assert np.array_equal(um_d.row_ptr.cpu().numpy(), um_h.row_ptr.numpy())
    assert np.array_equal(um_d.col.cpu().numpy(), um_h.col.numpy())
    assert np.array_equal(um_d.val.cpu().numpy().view(np.int32), um_h.val.numpy().view(np.int32))
    return rated_d, um_d


def test_collate_device_bit_exact_vs_reference_golden():
    from deeprecommendation_b200.content_providers import DeviceCollateProvider
    d, _, _ = load('collate')
    n_items, n_users = d['profiles'].shape[0], len(d['mean_rating'])
    p = DeviceCollateProvider(np.arange(n_items), d['profiles'], np.arange(n_users), d['row_ptr'], d['rated_idx'], d['rated_rating'], device=DEV)
    batch = [(int(u), int(i), 3.5) for u, i in zip(d['batch_users'], d['batch_items'])]
    cand_ids, rated_ids, cand, rated, um, tgt = p.collate_interacted_items(batch, for_ranking=False)
    assert np.array_equal(np.asarray(rated_ids), d['rated_items_idx']) and len(rated_ids) == len(d['rated_items_idx'])
    assert np.array_equal(cand_ids, d['batch_items'])
    assert np.array_equal(um.to_dense().cpu().numpy().view(np.int32), d['user_matrix'].view(np.int32))
    assert np.array_equal(rated.dense().cpu().numpy().view(np.int32), d['rated_items'].view(np.int32))
    assert np.array_equal(cand.dense().cpu().numpy().view(np.int32), d['candidate_items'].view(np.int32))
    assert tgt.tolist() == [3.5] * len(batch)
    out = p.collate_interacted_items([(0, 1, 2), (3, 4, 5)], for_ranking=True)
    assert out[5].dense().shape == (2, d['profiles'].shape[1])


@pytest.mark.parametrize('n_users,n_items,mean_len,B', [
    (300, 2000, 60, 256),           # shared-memory bitmap, one scan pass
    (64, 31, 10, 40),               # catalogue smaller than one bitmap word
    (2500, 70_000, 40, 2300),       # bitmap of 2,188 words and 2,300 rows: both scans take several passes
    (50, 500_000, 3000, 33),        # catalogue above the shared-memory limit: global bitmap path, long rows
])
def test_collate_device_equals_host_collate(n_users, n_items, mean_len, B):
    p = _providers(n_users, n_items, mean_len, seed=n_items % 97)
    rng = np.random.default_rng(B)
    user_idx = rng.integers(0, n_users, B)              # repeats are normal: a user appears once per (user, item) sample of the batch
    _check_against_host(p, user_idx)
    _check_against_host(p, user_idx, ignore_ratings=True)
    _check_against_host(p, user_idx[:1])


def test_collate_device_degenerate_batches():
    p = _providers(40, 300, 12, seed=5)
    empty_user, flat_user = 3, 2                         # (see _lists)
    assert p._list_len[empty_user] == 0 and p._nz_cnt[flat_user] == 0 and p._list_len[flat_user] > 0
    rated, um = _check_against_host(p, np.array([empty_user, empty_user]))
    assert rated.numel() == 0 and um.shape == (2, 0)
    rated, um = _check_against_host(p, np.array([flat_user, empty_user]))      # items in the union, no entry in the matrix
    assert rated.numel() == p._list_len[flat_user] and um.val.numel() == 0
    rated, um = _check_against_host(p, np.zeros(0, dtype=np.int64))
    assert rated.numel() == 0 and um.shape == (0, 0) and um.row_ptr.tolist() == [0]


def test_attention_scores_from_device_collate_equal_host_collate():
    """(user, item) samples -> scores through DynamicPointwiseDataset.do_forward: the device-collated batch gives the bits of the
    host-collated resident batch (itself bit-equal to the reference's dense contract, tests/test_models_gpu.py)."""
    from deeprecommendation_b200.content_providers import DeviceCollateProvider, ResidentDynamicProvider
    from deeprecommendation_b200.neural_collaborative_filtering.datasets.dynamic_datasets import DynamicPointwiseDataset, DynamicRankingDataset
    from deeprecommendation_b200.neural_collaborative_filtering.models import AttentionNCF
    users_raw, items_raw, ratings = synth.interactions_small(200, 3000, 60_000, seed=3)
    _, u = synth.dense_ids(users_raw)
    item_ids, it = synth.dense_ids(items_raw)
    n_items = len(item_ids)
    profiles = synth.item_profiles(n_items, seed=4, f_binary=300, f_dense=300)
    row_ptr, idx, rr, _ = synth.user_rating_lists(u, it, ratings, 200)
    args = (np.arange(n_items), profiles, np.arange(200), row_ptr, idx, rr)
    host, dev = ResidentDynamicProvider(*args, device=DEV), DeviceCollateProvider(*args, device=DEV)
    kw = dict(item_dim=600, item_emb=128, user_emb=128, att_dense=128, mlp_dense_layers=[256, 128], dropout_rate=0.2)
    m = AttentionNCF(**kw).to(DEV).eval()
    m.load_state_dict(synth.to_torch(synth.attention_ncf_weights(seed=9, **kw)))
    pick = np.random.default_rng(0).permutation(len(u))[:256]
    batch = [(int(a), int(b), 3.0) for a, b in zip(u[pick], it[pick])]
    with torch.no_grad():
        out_h, _, _, ids_h, att_h, _ = DynamicPointwiseDataset.do_forward(m, host.collate_interacted_items(batch, False), DEV, True)
        out_d, _, _, ids_d, att_d, _ = DynamicPointwiseDataset.do_forward(m, dev.collate_interacted_items(batch, False), DEV, True)
        assert torch.equal(out_h, out_d) and torch.equal(att_h, att_d) and np.array_equal(np.asarray(ids_d), ids_h)
        rbatch = [(int(a), int(b), int(c)) for a, b, c in zip(u[pick], it[pick], it[pick[::-1]])]
        pos_h, neg_h = DynamicRankingDataset.do_forward(m, host.collate_interacted_items(rbatch, True), DEV)
        pos_d, neg_d = DynamicRankingDataset.do_forward(m, dev.collate_interacted_items(rbatch, True), DEV)
        assert torch.equal(pos_h, pos_d) and torch.equal(neg_h, neg_d)
    # training keeps the dense contract: the device CSR is expanded on the device and gradients flow
    m.train()
    out = DynamicPointwiseDataset.do_forward(m, dev.collate_interacted_items(batch[:32], False), DEV)[0]
    out.sum().backward()
    assert m.ItemEmbeddings[0].weight.grad is not None and torch.isfinite(m.ItemEmbeddings[0].weight.grad).all()


def test_collate_launch_finish_pipelined():
    """several collates in flight (the loader enqueues batch k + 1 ahead of the forward of batch k): every handle resolves to its own batch"""
    p = _providers(300, 5000, 80, seed=11)
    rng = np.random.default_rng(2)
    batches = [torch.from_numpy(rng.integers(0, 300, 128)).pin_memory() for _ in range(6)]
    pending = [p.collate_launch(b) for b in batches]
    for b, h in zip(batches, pending):
        rated_d, um_d = p.collate_finish(h)
        rated_h, um_h = p.collate_csr(b.numpy())
        assert np.array_equal(rated_d.cpu().numpy(), rated_h)
        assert np.array_equal(um_d.row_ptr.cpu().numpy(), um_h.row_ptr.numpy()) and np.array_equal(um_d.col.cpu().numpy(), um_h.col.numpy())
        assert np.array_equal(um_d.val.cpu().numpy().view(np.int32), um_h.val.numpy().view(np.int32))
