"""Multi-GPU GraphNCF propagation: 1-D row partition of the node embeddings with a per-layer all-gather.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Destination rows of the neighbour index are split
into P contiguous ranges with (nearly) equal numbers of EDGES — degrees are heavy-tailed, equal row counts would not
balance.  Rank r owns rows [splits[r], splits[r+1]) of x, of the CSR and of every output.  Per layer:

    t_own   = dinv ∘ (W_type x_own + b_type)          K1a, written straight into rank r's slot of the gather buffer
    T       = all_gather(t_own)                        (P, max_rows, d) padded slots; NCCL, in place
    x'_own  = dinv ∘ (A_own · T)                       K3 over the local rows; `col` was remapped once to slot addressing

BasicNCF / AttentionNCF need no collective: pairs are sharded across ranks (bench.py).

The partition arithmetic (`split_rows`, `RowPartition`) is pure torch so that world_size-2 gloo tests cover it on CPU
(tests/test_parallel_cpu.py); the propagation itself only runs on CUDA.
"""
from __future__ import annotations

import torch

from . import ops
from .graph import GraphIndex, get_index


def split_rows(row_ptr: torch.Tensor, parts: int) -> list:
    """P+1 row boundaries such that every part holds ~nnz/P entries (ties broken towards more rows on the left)."""
    n_rows = int(row_ptr.numel() - 1)
    nnz = int(row_ptr[-1])
    targets = torch.tensor([(nnz * k) // parts for k in range(1, parts)], dtype=row_ptr.dtype, device=row_ptr.device)
    cuts = torch.searchsorted(row_ptr[1:].contiguous(), targets, right=False) + 1 if parts > 1 else targets
    splits = [0] + [min(int(c), n_rows) for c in cuts.tolist()] + [n_rows]
    for k in range(1, len(splits)):                      # monotone even for degenerate graphs
        splits[k] = max(splits[k], splits[k - 1])
    return splits


class RowPartition:
    """Ownership map of the N node rows and the slot addressing of the gathered feature buffer."""

    def __init__(self, splits, rank: int):
        self.splits, self.rank, self.parts = list(splits), rank, len(splits) - 1
        self.r0, self.r1 = self.splits[rank], self.splits[rank + 1]
        self.rows = self.r1 - self.r0
        self.max_rows = max(b - a for a, b in zip(self.splits[:-1], self.splits[1:]))
        self.max_rows = max(1, (self.max_rows + 3) // 4 * 4)

    def slot_index(self, ids: torch.Tensor) -> torch.Tensor:
        """global node id -> row of the (P * max_rows, d) gathered buffer"""
        bounds = torch.tensor(self.splits[1:-1], dtype=ids.dtype, device=ids.device)
        owner = torch.searchsorted(bounds, ids, right=True) if self.parts > 1 else torch.zeros_like(ids)
        starts = torch.tensor(self.splits[:-1], dtype=ids.dtype, device=ids.device)
        return owner * self.max_rows + (ids - starts[owner])

    def local_csr(self, row_ptr, col, *per_edge):
        """slices of a full CSR for the owned rows; `col` comes back in slot addressing"""
        k0, k1 = int(row_ptr[self.r0]), int(row_ptr[self.r1])
        lrp = (row_ptr[self.r0:self.r1 + 1] - k0).contiguous()
        lcol = self.slot_index(col[k0:k1].long()).to(col.dtype).contiguous()
        return (lrp, lcol) + tuple(None if a is None else a[k0:k1].contiguous() for a in per_edge)


class _LocalIndex(GraphIndex):
    """GraphIndex over the owned rows only (no edge lists / hash: inference-side object)."""

    def __init__(self, row_ptr, col, w, pos, dinv_own, chunk):
        self.row_ptr, self.col, self.w, self.pos = row_ptr, col, w, pos
        self.dinv = dinv_own
        self.num_nodes = int(row_ptr.numel() - 1)
        self.e1, self.e2 = int(col.numel()), 0
        self.chunk_size = chunk
        self.symmetric, self.w_bwd, self._hash = False, None, None
        self._build_plan()


class PartitionedGraph:
    """What rank `rank` keeps of a graph: its rows of the index and the slot map.  Build with `partition_graph`."""

    def __init__(self, graph, group=None, rank=None, world=None):
        import torch.distributed as dist
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        full = get_index(graph)                      # every rank builds the (cheap, ~10 ms at 50M edges) full index, keeps a slice
        self.nI = int(graph.item_features.shape[0])
        self.N = full.num_nodes
        self.part = RowPartition(split_rows(full.row_ptr, self.world), self.rank)
        lrp, lcol, lw, lpos = self.part.local_csr(full.row_ptr, full.col, full.w, full.pos)
        self.dinv_own = full.dinv[self.part.r0:self.part.r1].contiguous()
        self.index = _LocalIndex(lrp, lcol, lw, lpos, self.dinv_own, full.chunk_size)
        self.edges_total = full.e1 + full.e2
        self.edges_own = int(lcol.numel())
        p = self.part
        # own rows split by node type: items are global rows [0, nI), users [nI, N)
        self.items_own = (p.r0, min(p.r1, self.nI)) if p.r0 < self.nI else (p.r0, p.r0)
        self.users_own = (max(p.r0, self.nI), p.r1) if p.r1 > self.nI else (p.r1, p.r1)
        self.item_features = graph.item_features[self.items_own[0]:self.items_own[1]]
        self.user_features = graph.user_features[self.users_own[0] - self.nI:self.users_own[1] - self.nI]
        self._tg = None

    def gather_buffer(self, d, device):
        if self._tg is None or self._tg.shape[1] != d:
            self._tg = torch.zeros((self.world * self.part.max_rows, d), dtype=torch.float32, device=device)
        return self._tg


def partition_graph(graph, group=None) -> PartitionedGraph:
    pg = PartitionedGraph(graph, group)
    try:
        object.__setattr__(graph, '_b200rec_partition', pg)
    except Exception:
        pass
    return pg


def encode_partitioned(model, pg: PartitionedGraph):
    """Propagated + combined embeddings of ALL nodes in slot addressing, (P*max_rows, width), identical on every rank.
    Inference only (the training path of GraphNCF is single-GPU for now)."""
    import torch.distributed as dist
    if torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters()):
        raise NotImplementedError('partitioned GraphNCF propagation is inference-only; wrap the call in torch.no_grad()')
    if model.concat:
        raise NotImplementedError('concat=True is not supported by the partitioned path yet')
    p, idx = pg.part, pg.index
    dev = pg.dinv_own.device
    L_ = len(model.gnn_convs)
    d = model.item_embeddings[0].weight.shape[0]
    ie, ue = model.item_embeddings[0], model.user_embeddings[0]
    ni = pg.items_own[1] - pg.items_own[0]              # owned rows = [items | users], items first
    x0 = torch.empty((p.rows, d), dtype=torch.float32, device=dev)
    if ni:
        ops.linear_raw(pg.item_features, ie.weight, ie.bias, out=x0[:ni])
    if p.rows - ni:
        ops.linear_raw(pg.user_features, ue.weight, ue.bias, out=x0[ni:])
    tg = pg.gather_buffer(d, dev)
    mine = tg[p.rank * p.max_rows: p.rank * p.max_rows + p.rows]
    slot = tg[p.rank * p.max_rows: (p.rank + 1) * p.max_rows]
    acc = torch.empty((p.rows, d), dtype=torch.float32, device=dev)
    x, spare = x0, (torch.empty((p.rows, d), dtype=torch.float32, device=dev) if L_ > 1 else None)
    if L_:
        lin_u, lin_i, _ = model.gnn_convs[0].typed()
    for l in range(L_):
        if ni:
            ops.linear_raw(x[:ni], lin_i.weight, lin_i.bias, row_scale=pg.dinv_own[:ni], out=mine[:ni])
        if p.rows - ni:
            ops.linear_raw(x[ni:], lin_u.weight, lin_u.bias, row_scale=pg.dinv_own[ni:], out=mine[ni:])
        if pg.world > 1:
            dist.all_gather_into_tensor(tg, slot, group=pg.group)         # in place: rank r's slot is already in position
        last = l == L_ - 1
        xn = None if last else spare
        ops.spmm_raw(idx, tg, w=idx.w, dinv=pg.dinv_own, x_next=xn, acc_in=x0 if l == 0 else acc, acc_out=acc,
                     acc_scale=1.0 / (L_ + 1) if last else 1.0)
        x = xn
    if L_ == 0:
        acc = x0
    # combined embeddings of every node, for the batch gather
    out = torch.zeros((pg.world * p.max_rows, d), dtype=torch.float32, device=dev)
    out[p.rank * p.max_rows: p.rank * p.max_rows + p.rows] = acc
    if pg.world > 1:
        dist.all_gather_into_tensor(out, out[p.rank * p.max_rows: (p.rank + 1) * p.max_rows], group=pg.group)
    return out


def forward_partitioned(model, pg: PartitionedGraph, userIds, itemIds):
    """GraphNCF.forward on a partitioned graph: every rank propagates its rows, then scores its shard of the batch and
    the (B, 1) scores are all-gathered (NCF batches are data-parallel)."""
    import torch.distributed as dist
    comb = encode_partitioned(model, pg)
    B = userIds.shape[0]
    W = pg.world
    per = (B + W - 1) // W
    lo, hi = min(B, pg.rank * per), min(B, (pg.rank + 1) * per)
    u_slot, i_slot = pg.part.slot_index(userIds.long()), pg.part.slot_index(itemIds.long())
    from .neural_collaborative_filtering.util import run_mlp
    if model.MLP is None:
        mine = ops.rowdot(comb, comb, u_slot[lo:hi], i_slot[lo:hi])
    else:
        mine = run_mlp(model.MLP, comb, comb, idx0=i_slot[lo:hi], idx1=u_slot[lo:hi], training=False)
    if W == 1:
        return mine
    buf = torch.zeros((W * per, 1), dtype=torch.float32, device=comb.device)
    buf[pg.rank * per: pg.rank * per + (hi - lo)] = mine
    dist.all_gather_into_tensor(buf, buf[pg.rank * per:(pg.rank + 1) * per], group=pg.group)
    return buf[:B]
