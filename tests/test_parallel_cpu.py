"""Host-side logic of the multi-GPU GraphNCF path on CPU: nnz-balanced row partition, slot addressing of the gathered
feature buffer and the all-gather layout, with world_size-2 (and 3) gloo process groups.  The SpMM arithmetic is emulated
with torch index_add_ here (the CUDA kernel is covered by the -m gpu tests and tests/mp_graph_check.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deeprecommendation_b200 import synth
from deeprecommendation_b200.parallel import RowPartition, column_slice_csr, split_rows
from oracle import restatement as R


def _graph(seed=5, n_users=300, n_items=200, n=9000):
    users, items, ratings = synth.interactions_zipf(n_users, n_items, n, seed=seed)
    g = R.create_graph(users, items, ratings, np.arange(n_users), np.arange(n_items))
    total = np.concatenate([g['user2item_edge_index'], g['item2user_edge_index']], axis=1)
    N = n_users + n_items
    row_ptr, col, perm = R.csr_by_destination(total, N)
    w = np.concatenate([g['user2item_edge_attr'], g['item2user_edge_attr']])[perm]
    return torch.from_numpy(row_ptr), torch.from_numpy(col), torch.from_numpy(w), N


def _spmm(row_ptr, col, w, feats, n_rows):
    dst = torch.repeat_interleave(torch.arange(n_rows), row_ptr[1:] - row_ptr[:-1])
    return torch.zeros(n_rows, feats.shape[1], dtype=torch.float64).index_add_(0, dst, w.double()[:, None] * feats.double()[col])


def test_split_rows_balances_edges():
    row_ptr, col, w, N = _graph()
    nnz = int(row_ptr[-1])
    for P in (1, 2, 3, 4, 8):
        s = split_rows(row_ptr, P)
        assert s[0] == 0 and s[-1] == N and all(a <= b for a, b in zip(s[:-1], s[1:]))
        per = [int(row_ptr[b] - row_ptr[a]) for a, b in zip(s[:-1], s[1:])]
        assert sum(per) == nnz
        max_deg = int((row_ptr[1:] - row_ptr[:-1]).max())
        assert max(per) <= nnz / P + max_deg          # no part exceeds its share by more than one row
    # degenerate: more parts than rows with edges
    tiny = torch.tensor([0, 0, 5, 5])
    assert split_rows(tiny, 4)[-1] == 3


def test_slot_index_round_trip():
    part = RowPartition([0, 7, 7, 20, 31], rank=2)
    ids = torch.arange(31)
    slots = part.slot_index(ids)
    assert part.max_rows == 16 and len(torch.unique(slots)) == 31
    owner = slots // part.max_rows
    assert owner.tolist() == [0] * 7 + [2] * 13 + [3] * 11
    assert (slots - owner * part.max_rows).tolist() == list(range(7)) + list(range(13)) + list(range(11))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        row_ptr, col, w, N = _graph()
        d = 8
        feats = torch.from_numpy(np.random.default_rng(1).standard_normal((N, d)).astype(np.float32))
        part = RowPartition(split_rows(row_ptr, world), rank)
        lrp, lcol, lw = part.local_csr(row_ptr, col, w)
        # every rank contributes only ITS rows of the features; all-gather into padded slots
        mine = torch.zeros(part.max_rows, d)
        mine[:part.rows] = feats[part.r0:part.r1]
        slots = [torch.zeros(part.max_rows, d) for _ in range(world)]
        dist.all_gather(slots, mine)
        gathered = torch.cat(slots)
        out_own = _spmm(lrp, lcol, lw, gathered, part.rows)
        # reassemble on every rank and compare with the single-process product
        pad = torch.zeros(part.max_rows, d, dtype=torch.float64)
        pad[:part.rows] = out_own
        outs = [torch.zeros(part.max_rows, d, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(outs, pad)
        full = torch.cat([o[:b - a] for o, a, b in zip(outs, part.splits[:-1], part.splits[1:])])
        ref = _spmm(row_ptr, col, w, feats, N)
        ret[rank] = bool(torch.equal(full, ref)) and int(lcol.numel()) > 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_partitioned_propagation_matches_single_process(world):
    ret = mp.get_context('spawn').Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world and all(ret.values()), dict(ret)


# ---- scheme 'reduce': users partitioned, items replicated, one all-reduce of the item partials per layer ---------------------
def test_column_slice_csr_keeps_order_and_counts():
    row_ptr = torch.tensor([0, 3, 3, 7, 8])
    col = torch.tensor([5, 1, 9, 2, 6, 7, 3, 6])
    w = torch.arange(8.0)
    rp, c, ww = column_slice_csr(row_ptr, col, 5, 8, w)
    assert rp.tolist() == [0, 1, 1, 3, 4] and c.tolist() == [0, 1, 2, 1] and ww.tolist() == [0.0, 4.0, 5.0, 7.0]
    rp, c, none = column_slice_csr(row_ptr, col, 100, 200, None)
    assert rp.tolist() == [0, 0, 0, 0, 0] and c.numel() == 0 and none is None


def _worker_reduce(rank, world, port, ret):
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        n_users, n_items = 300, 200
        row_ptr, col, w, N = _graph(n_users=n_users, n_items=n_items)
        nI, d = n_items, 8
        feats = torch.from_numpy(np.random.default_rng(1).standard_normal((N, d)).astype(np.float32))
        k_items = int(row_ptr[nI])
        part = RowPartition(split_rows(row_ptr[nI:] - k_items, world), rank)        # users by their own-row nnz
        # (a) partial item rows from the owned users' edges, completed by an all-reduce
        lrp, lcol, lw = column_slice_csr(row_ptr[:nI + 1], col[:k_items], nI + part.r0, nI + part.r1, w[:k_items])
        own_user_feats = feats[nI + part.r0: nI + part.r1]
        partial = _spmm(lrp, lcol, lw, own_user_feats, nI)
        dist.all_reduce(partial)
        # (b) owned user rows from the replicated item table
        k0, k1 = int(row_ptr[nI + part.r0]), int(row_ptr[nI + part.r1])
        urp = row_ptr[nI + part.r0: nI + part.r1 + 1] - k0
        own_users = _spmm(urp, col[k0:k1], w[k0:k1], feats[:nI], part.rows)
        ref = _spmm(row_ptr, col, w, feats, N)
        counts = torch.tensor([int(lcol.numel())])
        dist.all_reduce(counts)
        ok = torch.allclose(partial, ref[:nI], rtol=1e-12, atol=1e-12) and torch.equal(own_users, ref[nI + part.r0: nI + part.r1])
        ret[rank] = bool(ok) and int(counts) == k_items
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_user_partitioned_reduce_scheme_matches_single_process(world):
    ret = mp.get_context('spawn').Manager().dict()
    mp.spawn(_worker_reduce, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world and all(ret.values()), dict(ret)


# ---- host logic of round 2's partitioned path (pure torch: runs without a GPU) ----------------------------------------------------------
def _walk_stream_plan(plan, t, d):
    """CPU model of csrc/spmm_stream.cu + the fix-up pass: every 'warp' walks its segment of the entry stream exactly as the kernel does
    (flag = last entry of a row, head / tail partial slots, rows_ne numbering, empty rows zero-filled)."""
    n_rows = plan.n_rows
    out = torch.full((n_rows, d), float('nan'), dtype=torch.float64)
    partials = torch.zeros((max(plan.n_slots, 1), d), dtype=torch.float64)
    colf, wd = plan.colf.tolist(), plan.wd.double()
    for s in range(plan.n_segs):
        k0, kend = s * plan.seg, min((s + 1) * plan.seg, plan.nnz)
        j = int(plan.seg_first_j[s])
        head = int(plan.seg_head_slot[s])
        open_is_head = head >= 0
        acc = torch.zeros(d, dtype=torch.float64)
        last_flag = True
        for k in range(k0, kend):
            c = colf[k]
            flag = c < 0
            acc = acc + wd[k] * t[c & 0x7fffffff]
            last_flag = flag
            if flag:
                if open_is_head:
                    partials[head] = acc
                    open_is_head = False
                else:
                    out[int(plan.rows_ne[j])] = acc
                acc = torch.zeros(d, dtype=torch.float64)
                j += 1
        if not last_flag:
            partials[head if open_is_head else int(plan.seg_tail_slot[s])] = acc
    for e in plan.rows_empty.tolist()[:plan.n_empty]:
        out[e] = 0.0
    for m in range(plan.n_multi):
        r, f, n = int(plan.multi_row[m]), int(plan.multi_first_slot[m]), int(plan.multi_n_slots[m])
        out[r] = partials[f:f + n].sum(0)
    return out


def test_stream_plan_walk_reproduces_the_spmm_on_cpu():
    from deeprecommendation_b200.graph import StreamPlan
    g = torch.Generator().manual_seed(0)
    for seg, deg in ((32, [3, 0, 0, 70, 1, 1, 0, 40, 5, 0]), (32, [32, 32, 64, 0, 128, 1, 31, 0, 0, 200, 32, 5]), (64, [0, 700, 0]), (32, [1] * 100),
                     (128, torch.randint(0, 90, (60,), generator=g).tolist())):
        deg_t = torch.tensor(deg)
        n = len(deg)
        rp = torch.zeros(n + 1, dtype=torch.int32)
        rp[1:] = torch.cumsum(deg_t, 0)
        nnz = int(rp[-1])
        col = torch.randint(0, 50, (nnz,), generator=g).int()
        w = torch.randn(nnz, generator=g)
        dinv = torch.rand(n, generator=g) + 0.1
        t = torch.randn(50, 8, generator=g).double()
        plan = StreamPlan(rp, col, w, dinv, seg)
        assert plan.n_segs * seg >= nnz and int((plan.colf < 0).sum()) == int((deg_t > 0).sum())
        got = _walk_stream_plan(plan, t, 8)
        rows = torch.repeat_interleave(torch.arange(n), deg_t)
        want = torch.zeros(n, 8, dtype=torch.float64).index_add_(0, rows, t[col.long()] * (w.double() * dinv.double()[rows])[:, None])
        assert not torch.isnan(got).any() and float((got - want).abs().max() / want.abs().max()) < 1e-6, (seg, deg)      # (wd is fp32)


def test_peer_arena_layout_is_symmetric_and_ordered():
    """the arena layout depends only on (world, rows per owner, d_max, batch_max) — identical on every rank — and its regions do not overlap"""
    from deeprecommendation_b200.peer import PeerShard, _rows_per_part
    for world, nI, d, bmax in ((8, 62423, 128, 512), (2, 11, 64, 16), (8, 1_000_000, 64, 512)):
        off = PeerShard.layout(world, _rows_per_part(nI, world), d, bmax)
        names = ['recv0', 'recv1', 'T0', 'T1', 'rows', 'total']
        vals = [off[k] for k in names]
        assert vals == sorted(vals) and vals[0] >= 4096 and all(v % 256 == 0 for v in vals)
        slab = world * _rows_per_part(nI, world) * d * 4
        assert all(b - a >= slab for a, b in zip(vals[:4], vals[1:5])) and off['total'] - off['rows'] >= 2 * bmax * d * 4
        assert world * _rows_per_part(nI, world) >= nI


def _peer_partition_worker(rank, world, port, q):
    import torch.distributed as dist
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    from deeprecommendation_b200.parallel import RowPartition, column_slice_csr, split_rows
    g = torch.Generator().manual_seed(5)                      # every rank derives the same CSR, keeps its own user range
    nI, nU = 40, 90
    deg = torch.randint(0, 30, (nU,), generator=g)
    rp = torch.zeros(nU + 1, dtype=torch.int32)
    rp[1:] = torch.cumsum(deg, 0)
    users = RowPartition(split_rows(rp, world), rank)
    edges_own = int(rp[users.r1] - rp[users.r0])
    tot = torch.tensor([edges_own, users.rows])
    dist.all_reduce(tot)
    # item rows restricted to my users' columns: the column slices of all ranks tile the item CSR exactly
    item_deg = torch.randint(0, 50, (nI,), generator=g)
    irp = torch.zeros(nI + 1, dtype=torch.int32)
    irp[1:] = torch.cumsum(item_deg, 0)
    icol = torch.randint(nI, nI + nU, (int(irp[-1]),), generator=g).int()
    lrp, lcol = column_slice_csr(irp, icol, nI + users.r0, nI + users.r1)[:2]
    cnt = torch.tensor([int(lcol.numel())])
    dist.all_reduce(cnt)
    q.put((rank, int(tot[0]) == int(rp[-1]), int(tot[1]) == nU, int(cnt[0]) == int(icol.numel()),
           bool(((lcol >= 0) & (lcol < users.rows)).all()) if lcol.numel() else True))
    dist.destroy_process_group()


def test_peer_user_partition_tiles_the_graph_world2_gloo():
    """world-size-2 gloo: the users' nnz-balanced ranges and the per-rank column slices of the item rows (the A_l operand of peer.py) tile the graph"""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_peer_partition_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(all(r[1:]) for r in res), res
