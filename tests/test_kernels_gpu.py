"""Kernel-level parity through the C ABI (libb200rec.so) against the CPU oracle / numpy, on a B200.

Tolerances: integer / index outputs bit-exact; fp32 mode max-norm relative error <= 1e-5 (north_star); bf16-table
mode <= 1e-2."""
import numpy as np
import pytest
import torch

from oracle import restatement as R
from tests._golden import maxnorm_rel

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-5
BF16_TOL = 1e-2


@pytest.fixture(scope='module')
def dev():
    return torch.device('cuda:0')


# ---- K1a linear ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('M,K,N', [(1, 7, 1), (37, 33, 20), (512, 2094, 128), (1000, 2094, 256), (130, 64, 130),
                                   (4096, 128, 128), (20000, 128, 128), (300, 2093, 17), (3, 4000, 512)])
@pytest.mark.parametrize('bias,scale,relu', [(True, False, False), (True, True, True), (False, False, True)])
def test_linear_fp32(dev, M, K, N, bias, scale, relu):
    from deeprecommendation_b200 import ops
    g = torch.Generator().manual_seed(M * 7 + K)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) if bias else None
    s = torch.rand(M, generator=g) + 0.5 if scale else None
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double() if bias else None)
    if scale:
        ref = ref * s.double()[:, None]
    if relu:
        ref = ref.relu()
    y = ops.linear_raw(x.to(dev), w.to(dev), b.to(dev) if bias else None, s.to(dev) if scale else None, relu)
    assert y.shape == (M, N)
    assert maxnorm_rel(y, ref) < FP32_TOL


def test_linear_strided_views_and_out(dev):
    """weight column slices (AttentionNet halves), input column slices and writing into a slice of a wider buffer"""
    from deeprecommendation_b200 import ops
    g = torch.Generator().manual_seed(1)
    x_full = torch.randn(200, 96, generator=g).to(dev)
    w_full = torch.randn(40, 128, generator=g).to(dev)
    x, w = x_full[:, 32:96], w_full[:, 64:]
    out_full = torch.full((200, 100), 7.0, device=dev)
    ops.linear_raw(x, w, None, None, False, out=out_full[:, 20:60])
    ref = x.double().cpu() @ w.double().cpu().T
    assert maxnorm_rel(out_full[:, 20:60], ref) < FP32_TOL
    assert torch.all(out_full[:, :20] == 7.0) and torch.all(out_full[:, 60:] == 7.0)


def test_linear_bf16_out(dev):
    from deeprecommendation_b200 import ops
    g = torch.Generator().manual_seed(2)
    x, w = torch.randn(300, 128, generator=g), torch.randn(64, 128, generator=g) / 11
    y = ops.linear_raw(x.to(dev), w.to(dev), None, None, False, out_dtype=torch.bfloat16)
    assert y.dtype == torch.bfloat16 and maxnorm_rel(y.float(), x.double() @ w.double().T) < BF16_TOL


# ---- K1b MLP tower -----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('B,E0,E1,layers', [(37, 16, 24, [32, 16, 1]), (512, 128, 128, [256, 1]), (5000, 128, 128, [256, 128, 1]),
                                            (100, 8, 8, [20, 1]), (64, 12, 4, [24, 12, 6, 1]), (33, 64, 0, [40, 3])])
def test_mlp_tower(dev, B, E0, E1, layers):
    from deeprecommendation_b200 import ops
    g = torch.Generator().manual_seed(B)
    a, b = torch.randn(B, E0, generator=g), (torch.randn(B, E1, generator=g) if E1 else None)
    ws, bs, prev = [], [], E0 + E1
    for h in layers:
        ws.append(torch.randn(h, prev, generator=g) / prev ** 0.5)
        bs.append(torch.randn(h, generator=g))
        prev = h
    x = torch.cat((a, b), 1).double() if E1 else a.double()
    for n, (w, bb) in enumerate(zip(ws, bs)):
        if n:
            x = x.relu()
        x = x @ w.double().T + bb.double()
    y = ops.mlp_tower_raw(a.to(dev), b.to(dev) if E1 else None, [w.to(dev) for w in ws], [t.to(dev) for t in bs])
    assert y.shape == x.shape and maxnorm_rel(y, x) < FP32_TOL


def test_mlp_tower_gather(dev):
    from deeprecommendation_b200 import ops
    g = torch.Generator().manual_seed(5)
    table = torch.randn(300, 32, generator=g)
    i0, i1 = torch.randint(0, 300, (77,), generator=g), torch.randint(0, 300, (77,), generator=g)
    w1, b1, w2, b2 = torch.randn(48, 64, generator=g) / 8, torch.randn(48, generator=g), torch.randn(1, 48, generator=g) / 7, torch.randn(1, generator=g)
    ref = (torch.cat((table[i0], table[i1]), 1).double() @ w1.double().T + b1.double()).relu() @ w2.double().T + b2.double()
    t = table.to(dev)
    y = ops.mlp_tower_raw(t, t, [w1.to(dev), w2.to(dev)], [b1.to(dev), b2.to(dev)], i0.to(dev), i1.to(dev))
    assert maxnorm_rel(y, ref) < FP32_TOL
    d = ops.rowdot(t, t, i0.to(dev), i1.to(dev))
    assert maxnorm_rel(d.view(-1), (table[i0].double() * table[i1].double()).sum(1)) < FP32_TOL


# ---- scan / sort (K4 building blocks) -----------------------------------------------------------------------------------
@pytest.mark.parametrize('n', [0, 1, 5, 2047, 2048, 2049, 100_000, 5_000_000])
def test_exclusive_scan(dev, n):
    import ctypes as C
    from deeprecommendation_b200 import _lib as L
    lib = L.lib()
    rng = np.random.default_rng(n)
    a = rng.integers(0, 5, size=n).astype(np.int32)
    t = torch.from_numpy(a).to(dev)
    out = torch.empty(n + 1, dtype=torch.int32, device=dev)
    ws = torch.empty(max(lib.b200rec_scan_workspace(n), 16), dtype=torch.uint8, device=dev)
    L.check(lib.b200rec_exclusive_scan_i32(C.c_void_p(t.data_ptr()), n, C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()),
                                           ws.numel(), None), 'scan')
    ref = np.concatenate([[0], np.cumsum(a, dtype=np.int64)]).astype(np.int32)
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize('n,bits', [(1, 3), (1000, 8), (4097, 11), (300_000, 18), (3_000_000, 24), (100_000, 31)])
def test_stable_sort_pairs(dev, n, bits):
    import ctypes as C
    from deeprecommendation_b200 import _lib as L
    lib = L.lib()
    rng = np.random.default_rng(bits)
    k = rng.integers(0, 2 ** bits if bits < 31 else 2 ** 31 - 1, size=n).astype(np.int32)
    if n > 100:
        k[: n // 2] = k[n // 2: 2 * (n // 2)] % 97            # many duplicates: stability matters
    v = np.arange(n, dtype=np.int32)
    kt, vt = torch.from_numpy(k.copy()).to(dev), torch.from_numpy(v.copy()).to(dev)
    ws = torch.empty(lib.b200rec_sort_pairs_workspace(n), dtype=torch.uint8, device=dev)
    L.check(lib.b200rec_sort_pairs_i32(C.c_void_p(kt.data_ptr()), C.c_void_p(vt.data_ptr()), n, bits, C.c_void_p(ws.data_ptr()),
                                       ws.numel(), None), 'sort')
    order = np.argsort(k, kind='stable')
    assert np.array_equal(kt.cpu().numpy(), k[order])
    assert np.array_equal(vt.cpu().numpy(), v[order])          # bit-exact permutation == numpy's stable argsort


# ---- row-wise top-k (serving path: webapp/backend.py:113-121, BASELINE configs[3]) -----------------------------------
@pytest.mark.parametrize('R,C,k', [(1, 9724, 10), (37, 1000, 5), (8, 100_000, 20), (3, 7, 10), (5, 300, 64), (2, 1, 1)])
def test_topk_rows_matches_stable_sort(dev, R, C, k):
    """indices bit-exact against a stable descending sort (ties -> lower column, as pandas' mergesort-free
    `sort_values(ascending=False)` would only guarantee for distinct scores; we pin the stable order)"""
    from deeprecommendation_b200 import ops
    g = torch.Generator().manual_seed(R * 131 + C)
    s = torch.randn(R, C, generator=g)
    s[:, ::7] = s[:, :1].clone()                          # many exact ties
    if C > 3:
        s[0, 2] = float('nan')                            # never selected
    val, idx = ops.topk_rows(s.to(dev), k)
    sn = s.numpy().copy()
    sn[np.isnan(sn)] = -np.inf
    order = np.argsort(-sn, axis=1, kind='stable')[:, :k]
    kk = min(k, C)
    ref_idx = np.full((R, k), -1, dtype=np.int64)
    ref_val = np.full((R, k), -np.inf, dtype=np.float32)
    ref_idx[:, :kk] = order[:, :kk]
    ref_val[:, :kk] = np.take_along_axis(sn, order[:, :kk], 1)
    nanpos = np.isinf(ref_val) & (ref_idx >= 0) & np.isnan(np.take_along_axis(s.numpy(), np.maximum(ref_idx, 0), 1))
    ref_idx[nanpos] = -1                                  # a NaN column is reported as "no entry"
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert np.array_equal(val.cpu().numpy(), ref_val)


# ---- K2 backward ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('mode,B,I,H,U', [(0, 70, 900, 128, 128), (0, 9, 300, 132, 64), (1, 40, 500, 128, 128), (0, 5, 64, 8, 260), (1, 3, 50, 4, 12)])
def test_attention_pool_backward_vs_float64_autograd(dev, mode, B, I, H, U):
    """csrc/attention_pool_bwd.cu against autograd of a float64 dense restatement: ragged rows (one empty, one full), exact-zero
    ratings, score_scale != 1.  Tolerance 1e-4 of each gradient's max norm (fp32 kernel, hardware-ordered vector reductions)."""
    from deeprecommendation_b200 import _lib as L
    from deeprecommendation_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + I + H)
    Pc, Pr, Q = torch.randn(B, H, generator=g) * 0.5, torch.randn(I, H, generator=g) * 0.5, torch.randn(I, U, generator=g)
    a2, a20, bU = torch.randn(H, generator=g) * 0.3, torch.randn(1, generator=g), torch.randn(U, generator=g)
    dens = torch.rand(B, 1, generator=g) * 0.4
    um = (torch.randint(1, 11, (B, I), generator=g).float() * 0.5 - 2.75) * (torch.rand(B, I, generator=g) < dens)
    um[0] = 0.0                                             # nothing rated
    um[1] = torch.randint(1, 11, (I,), generator=g).float() * 0.5 - 2.75    # everything rated
    gout = torch.randn(B, U, generator=g)
    scale = 1.25
    leaves = [t.double().requires_grad_(True) for t in (Pc, Pr, Q, a2, a20, bU)]
    ref, _ = R.attention_pool_factorised(*leaves, um.double(), um != 0, net=mode == 0, scale=scale)     # oracle/restatement.py
    ref_grads = torch.autograd.grad(ref, leaves, gout.double(), allow_unused=True)
    d = [t.to(dev).requires_grad_(True) for t in (Pc, Pr, Q, a2, a20, bU)]
    out = ops.attention_pool(d[0], d[1], d[2], mode=L.ATT_NET if mode == 0 else L.ATT_DOT, a2=d[3] if mode == 0 else None,
                             a20=d[4] if mode == 0 else None, bU=d[5], user_matrix=um.to(dev), score_scale=scale)
    assert maxnorm_rel(out.detach(), ref.detach()) < FP32_TOL
    out.backward(gout.to(dev))
    names = ['Pc', 'Pr', 'Q', 'a2', 'a20', 'bU']
    for k, (t, rg) in enumerate(zip(d, ref_grads)):
        if mode == 1 and names[k] in ('a2', 'a20'):
            assert t.grad is None
            continue
        if names[k] == 'a20':                               # shifts every score of a row: true gradient 0 (softmax shift invariance)
            assert float(t.grad.abs().max()) < 1e-4 * float(ref_grads[3].abs().max())
            continue
        assert t.grad.shape == rg.shape
        assert maxnorm_rel(t.grad, rg) < 1e-4, names[k]
