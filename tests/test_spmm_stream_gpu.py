"""K3s (csrc/spmm_stream.cu): the edge-balanced inference SpMM against (1) a float64 torch restatement of one propagation step
(x' = dinv ∘ (A_w · t), gnn_ncf.py:39-94 after transform-before-gather) and (2) the row-owner kernel K3 on the same index.
Segment sizes down to 32 entries make every code path common: rows cut by 1..n segment boundaries, rows that end exactly on a
boundary, runs of one-entry rows inside a batch of 8, empty rows, an empty graph."""
import numpy as np
import pytest
import torch

from tests._golden import maxnorm_rel

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


class _Idx:
    """bare local index (what parallel._LocalIndex / GraphIndex expose to the SpMM wrappers)"""

    def __init__(self, deg, n_src, seed, binary=False):
        from deeprecommendation_b200.parallel import _LocalIndex
        g = torch.Generator().manual_seed(seed)
        deg = torch.as_tensor(deg, dtype=torch.int64)
        rp = torch.zeros(deg.numel() + 1, dtype=torch.int32)
        rp[1:] = torch.cumsum(deg, 0)
        nnz = int(rp[-1])
        col = torch.randint(0, n_src, (max(nnz, 1),), generator=g, dtype=torch.int64)[:nnz].int()
        w = None if binary else (torch.randint(1, 11, (nnz,), generator=g).float() * 0.5 - 2.75)
        dinv = torch.rand(deg.numel(), generator=g) + 0.1
        self.index = _LocalIndex(rp.to(DEV), col.to(DEV), None if w is None else w.to(DEV), torch.zeros(nnz, dtype=torch.int32, device=DEV),
                                 dinv.to(DEV), 256)
        self.rp, self.col, self.w, self.dinv = rp, col, w, dinv

    def reference(self, t):
        t = t.double().cpu()
        n = self.rp.numel() - 1
        out = torch.zeros(n, t.shape[1], dtype=torch.float64)
        rows = torch.repeat_interleave(torch.arange(n), (self.rp[1:] - self.rp[:-1]).long())
        vals = t[self.col.long()] * (1.0 if self.w is None else self.w.double()[:, None])
        out.index_add_(0, rows, vals)
        return out * self.dinv.double()[:, None]


def _degrees(kind, rng):
    if kind == 'zipf':
        d = (rng.zipf(1.3, 600) % 900).astype(np.int64)
        d[rng.integers(0, 600, 60)] = 0
        return d
    if kind == 'ones':                                   # one-entry rows only: a flag on every entry
        return np.ones(500, dtype=np.int64)
    if kind == 'boundary':                               # rows that end exactly on segment boundaries, then a row spanning many
        return np.array([32, 32, 64, 0, 128, 1, 31, 0, 0, 1000, 32, 5], dtype=np.int64)
    if kind == 'one_long':
        return np.array([0, 5000, 0], dtype=np.int64)
    raise ValueError(kind)


@pytest.mark.parametrize('kind', ['zipf', 'ones', 'boundary', 'one_long'])
@pytest.mark.parametrize('seg', [32, 64, 256])
@pytest.mark.parametrize('d,dtype', [(128, torch.float32), (96, torch.float32), (128, torch.bfloat16)])
def test_stream_matches_float64_and_row_owner_kernel(kind, seg, d, dtype):
    from deeprecommendation_b200 import ops
    from deeprecommendation_b200.graph import StreamPlan
    rng = np.random.default_rng(hash((kind, seg, d)) % 2 ** 31)
    ix = _Idx(_degrees(kind, rng), n_src=777, seed=seg + d, binary=(kind == 'ones'))
    n = ix.rp.numel() - 1
    t32 = torch.randn(777, d, generator=torch.Generator().manual_seed(1))
    t = t32.to(DEV).to(dtype)
    want = ix.reference(t.float())
    plan = StreamPlan(ix.index.row_ptr, ix.index.col, ix.index.w, ix.index.dinv, seg)
    xs = torch.full((n, d), float('nan'), device=DEV)
    ops.spmm_stream_raw(plan, t, x_next=xs)
    assert maxnorm_rel(xs, want) < 2e-6
    xo = torch.full((n, d), float('nan'), device=DEV)
    ops.spmm_raw(ix.index, t, w=ix.index.w, dinv=ix.index.dinv, x_next=xo)
    assert maxnorm_rel(xs, xo) < 2e-6
    assert torch.all(xs[(ix.rp[1:] == ix.rp[:-1]).to(DEV)] == 0)                        # rows without entries are zero-filled
    xs2 = torch.empty_like(xs)
    ops.spmm_stream_raw(plan, t, x_next=xs2)
    assert torch.equal(xs, xs2)                                                         # deterministic
    # fused running mean: acc_out = (acc_in + x') * scale, x_next optional
    acc_in = torch.randn(n, d, device=DEV)
    acc = torch.empty_like(acc_in)
    ops.spmm_stream_raw(plan, t, acc_in=acc_in, acc_out=acc, acc_scale=0.25)
    assert maxnorm_rel(acc, (acc_in.double().cpu() + want) * 0.25) < 2e-6


def test_stream_empty_graph_and_push_epilogue():
    from deeprecommendation_b200 import ops
    from deeprecommendation_b200.graph import StreamPlan
    ix = _Idx(np.zeros(40, dtype=np.int64), n_src=10, seed=0)
    plan = StreamPlan(ix.index.row_ptr, ix.index.col, ix.index.w, ix.index.dinv, 64)
    x = torch.full((40, 128), float('nan'), device=DEV)
    ops.spmm_stream_raw(plan, torch.randn(10, 128, device=DEV), x_next=x)
    assert torch.all(x == 0)
    # push: 3 "owners" of 50 rows each inside one buffer per owner, slot 1 of 2
    rng = np.random.default_rng(5)
    deg = (rng.zipf(1.4, 140) % 300).astype(np.int64)
    deg[::7] = 0
    ix = _Idx(deg, n_src=300, seed=3)
    t = torch.randn(300, 128, device=DEV)
    want = ix.reference(t)
    rpp, d = 50, 128
    bufs = [torch.zeros(2 * rpp * d, device=DEV) for _ in range(3)]
    import ctypes as C
    dst = (C.c_void_p * 3)(*[b.data_ptr() for b in bufs])
    for seg in (32, 128):
        for b in bufs:
            b.zero_()
        plan = StreamPlan(ix.index.row_ptr, ix.index.col, ix.index.w, ix.index.dinv, seg)
        ops.spmm_stream_raw(plan, t, push=(dst, 3, rpp, 1 * rpp * d, d))
        torch.cuda.synchronize()
        got = torch.cat([b.view(2, rpp, d)[1] for b in bufs])[:140]
        assert maxnorm_rel(got, want) < 2e-6 and all(torch.all(b.view(2, rpp, d)[0] == 0) for b in bufs)
