"""K1s (csrc/gather_sum.cu): the projection of one-hot / multi-hot profile rows as a gather-sum over W^T, against the dense Linear the
reference runs on the same rows (models/basic_ncf.py:38-39 fed by one_hot_provider.py:17-21 / fixed_profiles_provider.py:49-50) and
against the oracle restatement of BasicNCF.forward.  fp32: max-norm rel <= 1e-5; bf16 table: <= 1e-2."""
import numpy as np
import pytest
import torch

from oracle import restatement as R
from oracle import synth
from tests._golden import maxnorm_rel

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = 1e-5


def _multi_hot(M, K, p, seed, weighted=False):
    rng = np.random.default_rng(seed)
    x = (rng.random((M, K)) < p).astype(np.float32)
    x[M // 3] = 0.0                                              # an empty row
    x[M // 2, :] = 1.0 if K <= 64 else x[M // 2]                 # (a full row on narrow inputs)
    if weighted:
        x *= rng.integers(1, 8, (M, K)).astype(np.float32) * 0.25 - 1.0
    return x


@pytest.mark.parametrize('N', [32, 64, 128, 256])
@pytest.mark.parametrize('weighted', [False, True])
def test_linear_sparse_csr_equals_dense_linear(N, weighted):
    from deeprecommendation_b200 import ops
    M, K = 777, 966
    x = _multi_hot(M, K, 0.012, seed=N, weighted=weighted)
    g = torch.Generator().manual_seed(1)
    w, b = torch.randn(N, K, generator=g) / 30, torch.randn(N, generator=g) / 10
    want = torch.from_numpy(x).double() @ w.double().T + b.double()
    xd = torch.from_numpy(x).to(DEV)
    rp, col, val = ops.dense_to_csr(xd, 0, K)
    r, c = np.nonzero(x)
    assert np.array_equal(col.cpu().numpy(), c) and np.array_equal(val.cpu().numpy(), x[r, c])
    assert np.array_equal(np.diff(rp.cpu().numpy()), np.bincount(r, minlength=M))
    wd, bd = w.to(DEV), b.to(DEV)
    out = ops.linear_sparse_raw(wd, bd, csr=(rp, col, None if not weighted else val))
    assert maxnorm_rel(out, want) < TOL
    assert torch.equal(out, ops.linear_sparse_raw(wd, bd, csr=(rp, col, None if not weighted else val)))      # list order: deterministic
    base = torch.randn(M, N, device=DEV)
    acc = ops.linear_sparse_raw(wd, None, csr=(rp, col, val), out=base.clone(), accumulate=True)
    assert maxnorm_rel(acc, base.double().cpu() + want - b.double()) < TOL
    bf = ops.linear_sparse_raw(wd, bd, csr=(rp, col, val), table_dtype=torch.bfloat16)
    assert maxnorm_rel(bf, want) < 1e-2


def test_linear_sparse_column_window_and_one_hot_ids():
    from deeprecommendation_b200 import ops
    M, K, N = 300, 500, 128
    g = torch.Generator().manual_seed(2)
    w, b = (torch.randn(N, K, generator=g) / 20).to(DEV), (torch.randn(N, generator=g) / 10).to(DEV)
    x = _multi_hot(M, 200, 0.03, seed=5)
    rp, col, val = ops.dense_to_csr(torch.from_numpy(x).to(DEV), 0, 200)
    out = ops.linear_sparse_raw(w, b, csr=(rp, col, None), cols=(100, 300))                 # indices relative to column 100
    want = torch.from_numpy(x).double() @ w[:, 100:300].double().cpu().T + b.double().cpu()
    assert maxnorm_rel(out, want) < TOL
    ids = torch.from_numpy(np.random.default_rng(0).integers(0, K, M)).to(DEV)
    ids[7] = -1                                                                           # empty row -> bias only
    out = ops.linear_sparse_raw(w, b, ids=ids)
    want = w.t()[ids.clamp(min=0)] * (ids >= 0)[:, None] + b
    assert torch.equal(out, want)                                                         # a lookup + one add: exact


@pytest.mark.parametrize('mlp', [[256], [256, 128]])
def test_basic_ncf_mixed_item_profiles_vs_oracle(mlp):
    """item profiles = 966 multi-hot columns (~1 %) + 1128 dense columns handed out as MixedRows; dense user profiles"""
    from deeprecommendation_b200.content_providers import MixedRows
    from deeprecommendation_b200.neural_collaborative_filtering.models import BasicNCF
    kw = dict(item_dim=2094, user_dim=2094, item_emb=128, user_emb=128, mlp_dense_layers=mlp, dropout_rate=0.2)
    sd = synth.to_torch(synth.basic_ncf_weights(seed=1, **kw))
    m = BasicNCF(**kw).to(DEV).eval()
    m.load_state_dict(sd)
    xi = synth.item_profiles(512, seed=2)
    xu = (synth.item_profiles(512, seed=3) - 0.3) * 0.1
    ref = R.basic_ncf_forward(sd, torch.from_numpy(xu), torch.from_numpy(xi))
    mixed = MixedRows.from_dense(xi, 966)
    assert mixed.val is None and mixed.col.numel() < 0.02 * 512 * 966                     # binary, ~1 % dense
    with torch.no_grad():
        out = m(torch.from_numpy(xu).to(DEV), mixed.float().to(DEV))
        dense = m(torch.from_numpy(xu).to(DEV), torch.from_numpy(xi).to(DEV))
    assert maxnorm_rel(out, ref) < TOL and maxnorm_rel(dense, ref) < TOL


def test_basic_ncf_one_hot_rows_through_the_dataset_contract_vs_oracle():
    from deeprecommendation_b200.content_providers import OneHotArrayProvider
    from deeprecommendation_b200.neural_collaborative_filtering.datasets.fixed_datasets import FixedPointwiseDataset
    from deeprecommendation_b200.neural_collaborative_filtering.models import BasicNCF
    n_items, n_users = 1174, 5000                                                         # the reference's dataset sizes (thesis slide 28)
    kw = dict(item_dim=n_items, user_dim=n_users, item_emb=128, user_emb=128, mlp_dense_layers=[256, 128], dropout_rate=0.2)
    g = torch.Generator().manual_seed(4)
    m = BasicNCF(**kw).to(DEV).eval()
    sd = {k: torch.randn(v.shape, generator=g) * 0.05 for k, v in m.state_dict().items()}
    m.load_state_dict(sd)
    rng = np.random.default_rng(1)
    item_ids, user_ids = np.sort(rng.choice(10**6, n_items, replace=False)), np.sort(rng.choice(10**6, n_users, replace=False))
    batch = [(int(user_ids[rng.integers(n_users)]), int(item_ids[rng.integers(n_items)]), 3.5) for _ in range(257)]
    outs = {}
    for sparse in (True, False):
        cp = OneHotArrayProvider(item_ids, user_ids, sparse=sparse)
        users, items, targets = zip(*batch)
        uv, iv = cp.get_user_profile(np.array(users)), cp.get_item_profile(np.array(items))
        from deeprecommendation_b200.neural_collaborative_filtering.datasets.fixed_datasets import _tensor
        with torch.no_grad():
            outs[sparse], _ = FixedPointwiseDataset.do_forward(m, (_tensor(uv), _tensor(iv), torch.zeros(len(batch))), DEV)
        if not sparse:
            ref = R.basic_ncf_forward(sd, torch.from_numpy(uv), torch.from_numpy(iv))
    assert maxnorm_rel(outs[True], ref) < TOL and maxnorm_rel(outs[False], ref) < TOL


def test_sparse_projection_gradients_equal_dense_path():
    from deeprecommendation_b200.content_providers import MixedRows, OneHotRows
    from deeprecommendation_b200.neural_collaborative_filtering.models import BasicNCF
    kw = dict(item_dim=600, user_dim=300, item_emb=64, user_emb=64, mlp_dense_layers=[64], dropout_rate=None)
    xi = synth.item_profiles(200, seed=6, f_binary=280, f_dense=320)
    uid = np.random.default_rng(3).integers(0, 300, 200)
    xu = np.zeros((200, 300), dtype=np.float32)
    xu[np.arange(200), uid] = 1.0
    y = torch.randn(200, 1, device=DEV)
    grads = {}
    for mode in ('sparse', 'dense'):
        torch.manual_seed(0)
        m = BasicNCF(**kw).to(DEV).train()
        if mode == 'sparse':
            out = m(OneHotRows(uid, 300).to(DEV), MixedRows.from_dense(xi, 280).to(DEV))
        else:
            out = m(torch.from_numpy(xu).to(DEV), torch.from_numpy(xi).to(DEV))
        (out - y).square().sum().backward()
        grads[mode] = {k: p.grad.clone() for k, p in m.named_parameters()}
    for k in grads['dense']:
        assert maxnorm_rel(grads['sparse'][k], grads['dense'][k]) < 1e-4, k
