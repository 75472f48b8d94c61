"""K5 — all-pairs scoring + per-user top-k (csrc/allpairs.cu) against fp64 torch and the CPU oracle, through the C ABI.

Tolerances (north_star): fp32 mode (bf16 hi/lo split operands) max-norm relative error <= 1e-5, bf16 mode <= 1e-2.
The top-k selection is checked bit-exactly against a stable sort of the kernel's OWN score matrix (selection logic), and
within tolerance against the oracle's scores (arithmetic)."""
import numpy as np
import pytest
import torch

from tests._golden import maxnorm_rel

pytestmark = pytest.mark.gpu
TOL = {'fp32': 1e-5, 'bf16': 1e-2}


@pytest.fixture(scope='module')
def dev():
    return torch.device('cuda:0')


def _problem(nU, nI, H1, H2, seed):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(nU, H1, generator=g)
    B = torch.randn(nI, H1, generator=g)
    W2 = (torch.rand(H2, H1, generator=g) * 2 - 1) / H1 ** 0.5
    b2 = (torch.rand(H2, generator=g) * 2 - 1) / H1 ** 0.5
    w3 = (torch.rand(H2, generator=g) * 2 - 1) / H2 ** 0.5
    b3 = torch.randn(1, generator=g) * 0.1
    return A, B, W2, b2, w3, b3


def _ref_scores(A, B, W2, b2, w3, b3):
    h1 = torch.relu(A.double()[:, None, :] + B.double()[None, :, :])
    h2 = torch.relu(h1 @ W2.double().T + b2.double())
    return h2 @ w3.double() + b3.double()


@pytest.mark.parametrize('nU,nI,H1,H2,k,splits', [(5, 70, 64, 16, 10, 0), (33, 1000, 256, 128, 10, 0), (100, 257, 128, 100, 20, 3),
                                                   (1, 9724, 256, 128, 10, 0), (70, 33, 192, 128, 64, 1), (3, 5, 64, 8, 10, 0)])
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_allpairs_kernel_vs_fp64(dev, nU, nI, H1, H2, k, splits, precision):
    from deeprecommendation_b200 import _lib as L, ops
    from oracle import restatement as R
    A, B, W2, b2, w3, b3 = _problem(nU, nI, H1, H2, seed=nU * 31 + nI)
    mode = {'fp32': L.AP_BF16X2, 'bf16': L.AP_BF16}[precision]
    packed = ops.allpairs_pack(W2.to(dev), b2.to(dev), w3.to(dev), b3.to(dev), H1, mode)
    val, idx, scores = ops.allpairs_topk_raw(A.to(dev), B.to(dev), packed, mode, k, return_scores=True, n_splits=splits)
    ref = _ref_scores(A, B, W2, b2, w3, b3)
    assert scores.shape == (nU, nI)
    assert maxnorm_rel(scores, ref) < TOL[precision]
    # selection: exactly the stable top-k of the kernel's own scores
    sv, si = R.topk_stable(scores.cpu(), k)
    assert torch.equal(idx.cpu(), si)
    assert torch.equal(val.cpu(), sv)
    # and without materialising the scores the result is the same
    val2, idx2, none = ops.allpairs_topk_raw(A.to(dev), B.to(dev), packed, mode, k, n_splits=splits)
    assert none is None and torch.equal(idx2, idx) and torch.equal(val2, val)


def test_allpairs_seen_items_are_skipped(dev):
    from deeprecommendation_b200 import _lib as L, ops
    from oracle import restatement as R
    nU, nI, k = 37, 500, 10
    A, B, W2, b2, w3, b3 = _problem(nU, nI, 128, 64, seed=5)
    rng = np.random.default_rng(0)
    seen = [np.sort(rng.choice(nI, size=rng.integers(0, 400), replace=False)) for _ in range(nU)]
    seen[3] = np.arange(nI)                                  # everything seen -> all (-inf, -1)
    seen[4] = np.arange(nI - 4)                              # 4 candidates left < k
    ptr = np.concatenate(([0], np.cumsum([len(s) for s in seen]))).astype(np.int32)
    flat = np.concatenate(seen).astype(np.int32)
    packed = ops.allpairs_pack(W2.to(dev), b2.to(dev), w3.to(dev), b3.to(dev), 128, L.AP_BF16X2)
    val, idx, scores = ops.allpairs_topk_raw(A.to(dev), B.to(dev), packed, L.AP_BF16X2, k, return_scores=True,
                                             seen=(torch.from_numpy(ptr).to(dev), torch.from_numpy(flat).to(dev)))
    sv, si = R.topk_stable(scores.cpu(), k, seen=seen)
    assert torch.equal(idx.cpu(), si) and torch.equal(val.cpu(), sv)
    assert (idx[3] == -1).all() and (idx[4, 4:] == -1).all() and (idx[4, :4] >= nI - 4).all()


@pytest.mark.parametrize('mlp', [[256, 128], [256], [100, 40]])
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_basic_ncf_recommend_vs_oracle(dev, mlp, precision):
    """BasicNCF.recommend == reference forward on every pair + stable top-k (oracle/restatement.py)"""
    from deeprecommendation_b200 import synth
    from deeprecommendation_b200.neural_collaborative_filtering.models import BasicNCF
    from oracle import restatement as R
    kw = dict(item_dim=300, user_dim=300, item_emb=128, user_emb=128, mlp_dense_layers=mlp, dropout_rate=0.2)
    sd = synth.to_torch(synth.basic_ncf_weights(seed=11, **kw))
    m = BasicNCF(**kw).to(dev).eval()
    m.load_state_dict(sd)
    xi = torch.from_numpy(synth.item_profiles(333, seed=2, f_binary=150, f_dense=150))
    xu = torch.from_numpy((synth.item_profiles(21, seed=3, f_binary=150, f_dense=150) - 0.3) * 0.1)
    ref = R.basic_ncf_all_pairs(sd, xu, xi)
    val, idx, scores = m.recommend(xu.to(dev), xi.to(dev), k=10, precision=precision, return_scores=True)
    assert maxnorm_rel(scores, ref) < TOL[precision]
    rv, ri = R.topk_stable(ref, 10)
    # the k-th best reference score is reproduced; items may swap only between scores closer than the tolerance
    assert float((val.cpu() - rv).abs().max() / ref.abs().max()) < TOL[precision]      # same normalisation as the scores
    picked = torch.gather(ref, 1, idx.cpu())
    assert float((picked - rv).abs().max() / ref.abs().max()) < 2 * TOL[precision]
    if precision == 'fp32':
        assert (idx.cpu() == ri).float().mean() > 0.97


def test_recommend_matches_forward_on_pairs(dev):
    """the all-pairs path and the per-pair forward (K1a + K1b) agree on the same pairs"""
    from deeprecommendation_b200 import synth
    from deeprecommendation_b200.neural_collaborative_filtering.models import BasicNCF
    kw = dict(item_dim=200, user_dim=200, item_emb=128, user_emb=128, mlp_dense_layers=[256, 128], dropout_rate=0.2)
    m = BasicNCF(**kw).to(dev).eval()
    m.load_state_dict(synth.to_torch(synth.basic_ncf_weights(seed=5, **kw)))
    xi = torch.from_numpy(synth.item_profiles(90, seed=2, f_binary=100, f_dense=100)).to(dev)
    xu = torch.from_numpy(synth.item_profiles(40, seed=3, f_binary=100, f_dense=100) * 0.1).to(dev)
    _, _, scores = m.recommend(xu, xi, k=5, return_scores=True)
    with torch.no_grad():
        fwd = m(xu.repeat_interleave(90, 0), xi.repeat(40, 1)).view(40, 90)
    assert maxnorm_rel(scores, fwd) < 1e-5


def test_graph_ncf_recommend_vs_oracle(dev):
    """GraphNCF.recommend == oracle forward on every (user, item) pair + stable top-k, connected items left out"""
    from deeprecommendation_b200 import synth
    from deeprecommendation_b200.graph import IdTable, create_graph
    from deeprecommendation_b200.neural_collaborative_filtering.models import GraphNCF
    from oracle import restatement as R
    users, items, ratings = synth.interactions_zipf(120, 90, 2500, seed=7)
    nU, nI, F_ = 120, 90, 32
    kw = dict(item_dim=F_, user_dim=F_, num_gnn_layers=2, hetero=True, node_emb=64, mlp_dense_layers=[128, 64], dropout_rate=0.2)
    sd = synth.to_torch(synth.graph_ncf_weights(seed=8, **kw))
    rng = np.random.default_rng(3)
    fi, fu = rng.standard_normal((nI, F_)).astype(np.float32), rng.standard_normal((nU, F_)).astype(np.float32)
    g = R.create_graph(users, items, ratings, np.arange(nU), np.arange(nI))
    gd = {k: (torch.from_numpy(v) if v is not None else None) for k, v in g.items()}
    gd['item_features'], gd['user_features'] = torch.from_numpy(fi), torch.from_numpy(fu)
    uu = torch.arange(nI, nI + nU).repeat_interleave(nI)
    ii = torch.arange(nI).repeat(nU)
    with torch.no_grad():
        ref = R.graph_ncf_forward(sd, gd, uu, ii, 2).view(nU, nI)
    graph = create_graph(torch.from_numpy(users).to(dev), torch.from_numpy(items).to(dev), torch.from_numpy(ratings).to(dev),
                         torch.from_numpy(fi).to(dev), torch.from_numpy(fu).to(dev),
                         IdTable(torch.arange(nU, device=dev)), IdTable(torch.arange(nI, device=dev)))
    m = GraphNCF(**kw).to(dev).eval()
    m.load_state_dict(sd)
    val, idx, scores = m.recommend(graph, k=5, return_scores=True)
    assert maxnorm_rel(scores, ref) < 1e-5
    e = g['user2item_edge_index']
    seen = [np.unique(e[1][e[0] == nI + u]) for u in range(nU)]
    sv, si = R.topk_stable(scores.cpu(), 5, seen=seen)
    assert torch.equal(idx.cpu(), si) and torch.equal(val.cpu(), sv)
    rv, _ = R.topk_stable(ref, 5, seen=seen)
    assert float((val.cpu() - rv).abs().max() / ref.abs().max()) < 1e-5


def test_attention_recommend_for_user_vs_oracle(dev):
    """AttentionNCF.recommend_for_user reproduces webapp/backend.py:78-121 computed with the oracle forward"""
    from deeprecommendation_b200 import synth
    from deeprecommendation_b200.neural_collaborative_filtering.models import AttentionNCF
    from oracle import restatement as R
    kw = dict(item_dim=256, item_emb=128, user_emb=128, att_dense=128, mlp_dense_layers=[256, 128], dropout_rate=0.2)
    sd = synth.to_torch(synth.attention_ncf_weights(seed=4, **kw))
    m = AttentionNCF(**kw).to(dev).eval()
    m.load_state_dict(sd)
    prof = torch.from_numpy(synth.item_profiles(700, seed=5, f_binary=128, f_dense=128))
    rng = np.random.default_rng(1)
    rated = rng.choice(700, size=23, replace=False)
    ratings = rng.integers(1, 11, size=23) * 0.5
    out = m.recommend_for_user(prof.to(dev), torch.from_numpy(rated), torch.from_numpy(ratings), k=10)
    # oracle: the reference's own steps
    rs = np.sort(rated)
    r_sorted = ratings[np.argsort(rated)]
    cand = np.setdiff1d(np.arange(700), rs)
    um = np.repeat((r_sorted - (r_sorted.mean() + 2.5) / 2)[None, :], len(cand), 0).astype(np.float32)
    with torch.no_grad():
        y, att = R.attention_ncf_forward(sd, prof[cand], prof[rs], torch.from_numpy(um), return_attention_weights=True)
    y = y.view(-1)
    order = np.argsort(-y.numpy(), kind='stable')[:10]
    assert np.array_equal(out['items'].cpu().numpy(), cand[order])
    assert maxnorm_rel(out['scores'], y[order]) < 1e-5
    thr = 1.5 / 23 + 0.025
    for j, o in enumerate(order):
        mask = att[o].numpy() > thr
        assert np.array_equal(out['because'][j].cpu().numpy(), rs[mask])
        assert np.allclose(out['attention'][j].cpu().numpy(), att[o].numpy()[mask], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize('nU,nI,H1,k,splits', [(70, 1000, 256, 10, 0), (5, 129, 64, 3, 2), (300, 333, 192, 64, 3), (64, 128, 128, 1, 1)])
def test_relu_dot_kernel_is_exact_fp32(dev, nU, nI, H1, k, splits):
    """b200rec_allpairs_relu_dot_topk (one-hidden-layer MLPs in all-pairs mode, FP32 pipes): scores against float64 to fp32 rounding,
    top-k bit-equal to a stable sort of its own scores, `seen` pairs skipped, split lists merged."""
    from deeprecommendation_b200 import ops
    from oracle import restatement as R
    g = torch.Generator().manual_seed(nU + nI)
    A, B = torch.randn(nU, H1, generator=g), torch.randn(nI, H1, generator=g)
    w2, b2 = torch.randn(H1, generator=g) / H1 ** 0.5, torch.randn(1, generator=g)
    ref = (A.double()[:, None, :] + B.double()[None]).relu() @ w2.double() + b2.double()
    val, idx, sc = ops.allpairs_relu_dot_raw(A.to(dev), B.to(dev), w2.to(dev), b2.to(dev), k, return_scores=True, n_splits=splits)
    assert maxnorm_rel(sc, ref) < 2e-6
    sv, si = R.topk_stable(sc.cpu(), k)
    assert torch.equal(idx.cpu(), si) and torch.equal(val.cpu(), sv)
    # seen: user 0 has already interacted with its two best items
    ptr = torch.zeros(nU + 1, dtype=torch.int32)
    ptr[1:] = 2
    seen_items = torch.sort(si[0, :2]).values.int()
    v2, i2, _ = ops.allpairs_relu_dot_raw(A.to(dev), B.to(dev), w2.to(dev), b2.to(dev), k, seen=(ptr.to(dev), seen_items.to(dev)), n_splits=splits)
    assert not set(i2[0].cpu().tolist()) & set(seen_items.tolist())
    assert torch.equal(i2[1:].cpu(), si[1:])
