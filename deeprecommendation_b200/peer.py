"""Partitioned GraphNCF propagation over peer-mapped memory (scheme 'peer'; SURVEY.md §8e, csrc/peer.cu).

Replaces, on the P GPUs of one NVSwitch box, the whole-graph loop of the reference (models/gnn_ncf.py:336-345, recomputed for
every batch) WITHOUT a library collective on the data path.  Ownership (same as scheme 'reduce' of parallel.py):

    users   nnz-balanced contiguous ranges; a rank holds x / acc / t of ITS users only
    items   equal contiguous ranges of `rpp` rows; a rank holds x / acc of its items, and a copy T of the transformed features
            t_items of ALL items (the table its user rows gather from)

Per layer l on rank r (one stream, `par = l & 1` selects one of two copies of every exchange buffer):

    A_l   K3 over ALL item rows restricted to the edges from r's users (sources: local t_users).  Its epilogue stores each
          partial row straight into the OWNER's receive slot recv[par][r] over NVLink  — reduce-scatter fused into the SpMM.
          signal(A)
    B_l   wait(T);  K3 over r's user rows from the table T[par]  (+ the fused running mean)    — hides A_l's stores in flight
    C_l   wait(A);  x_items' = sum of the P receive slots in rank order (deterministic), running mean;
          K1c transform of x_items' whose epilogue writes t_items into T[1-par] of EVERY rank  — all-gather fused into the GEMM.
          signal(T);  t_users for the next layer (local)

After the last layer the owners push the 2B batch rows into every rank's `rows` buffer (peer_gather_rows) and every rank runs the
MLP on the batch.  signal / wait are epoch flags inside the arenas (release / acquire at system scope; waits time out into an
error flag instead of hanging the GPU), counted on the device so that a captured CUDA graph can be replayed.

The arenas are cudaMalloc'd by libb200rec and exchanged as CUDA IPC handles through torch.distributed (plumbing).
`emulated_shards` builds all P shards inside ONE process (plain device buffers, the same kernels and addresses) so that the
exchange logic is covered by the single-GPU test-suite; there the ranks advance in lock step, segment by segment.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import torch

from . import _lib as L
from . import ops
from .graph import get_index
from .ops import _ptr, _stream

CH_T0, CH_T1, CH_A0, CH_A1, CH_R, CH_E = 0, 1, 2, 3, 4, 5
_CTRL_BYTES = 4096                 # flags uint32[16][16] (1 KB) | counters uint32[32] | err int32
_OFF_COUNTERS = L.PEER_CHANNELS * L.PEER_MAX * 4
_OFF_ERR = _OFF_COUNTERS + 2 * L.PEER_CHANNELS * 4
WAIT_TIMEOUT_NS = 10_000_000_000
# the item side of a layer (wait for the partials, slot reduction, transform + broadcast) runs on a second stream under the user-row SpMM
OVERLAP = os.environ.get('B200REC_PEER_OVERLAP', '1') == '1'
TRACE = os.environ.get('B200REC_PEER_TRACE', '0') == '1'


class _Raw:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {'shape': (int(nbytes),), 'typestr': '|u1', 'data': (int(ptr), False), 'version': 3, 'strides': None}


def _align(n, a=256):
    return (int(n) + a - 1) // a * a


class PeerArena:
    """`nbytes` of device memory on this rank, addressable from every rank of the group.  bases[q] = address of rank q's arena in
    THIS process (bases[rank] is the local one); `buf` is a uint8 torch view of the local arena."""

    def __init__(self, rank, world, bases, buf, owner=None):
        self.rank, self.world, self.bases, self.buf = rank, world, [int(b) for b in bases], buf
        self._owner = owner                                            # keeps the memory alive (emulation) / frees it (IPC)

    @classmethod
    def ipc(cls, nbytes, group=None, device=None):
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        if world > L.PEER_MAX:
            raise NotImplementedError(f'peer exchange addresses at most {L.PEER_MAX} GPUs of one box')
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        lib = L.lib()
        ptr = C.c_void_p()
        handle = C.create_string_buffer(L.PEER_HANDLE_BYTES)
        with torch.cuda.device(dev):
            L.check(lib.b200rec_peer_alloc(nbytes, C.byref(ptr), handle), 'peer_alloc')
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        bases, opened = [], []
        with torch.cuda.device(dev):
            for q in range(world):
                if q == rank:
                    bases.append(ptr.value)
                    continue
                p = C.c_void_p()
                L.check(lib.b200rec_peer_open(handles[q], C.byref(p)), 'peer_open')
                bases.append(p.value)
                opened.append(p.value)
        buf = torch.as_tensor(_Raw(ptr.value, nbytes), device=dev)

        class _Owner:
            pass
        owner = _Owner()

        def _release(local=ptr.value, opened=tuple(opened), dev=dev):
            try:
                with torch.cuda.device(dev):
                    for p in opened:
                        lib.b200rec_peer_close(C.c_void_p(p))
                    lib.b200rec_peer_free(C.c_void_p(local))
            except Exception:
                pass
        weakref.finalize(owner, _release)
        dist.barrier(group=group)                                      # nobody frees or writes before every rank has mapped everything
        return cls(rank, world, bases, buf, owner)

    @classmethod
    def emulated(cls, nbytes, world, device):
        """`world` arenas inside one process (tests): ordinary device buffers, every 'rank' sees every base address"""
        bufs = [torch.zeros(int(nbytes), dtype=torch.uint8, device=device) for _ in range(world)]
        bases = [b.data_ptr() for b in bufs]
        return [cls(q, world, bases, bufs[q], owner=bufs) for q in range(world)]

    def view(self, offset, shape, dtype=torch.float32):
        n = 1
        for s in shape:
            n *= int(s)
        nb = n * torch.empty((), dtype=dtype).element_size()
        return self.buf[offset: offset + nb].view(dtype).view(*shape)

    def ptrs(self, offset=0, first=None):
        """host array of the P addresses `base[q] + offset` (optionally rotated so that rank `first` comes first)"""
        order = list(range(self.world))
        if first is not None:
            order = [first] + [q for q in order if q != first]
        arr = (C.c_void_p * self.world)()
        for k, q in enumerate(order):
            arr[k] = self.bases[q] + offset
        return arr


class PeerShard:
    """What rank `rank` keeps of a user-partitioned graph + the arena layout (identical on every rank)."""

    def __init__(self, *, rank, world, nI, nU, d_max, rpp, users_r0, users_rows, index_users, index_items, dinv_users, dinv_items_all,
                 item_features_own, user_features_own, arena, batch_max, edges_total, edges_own):
        self.rank, self.world, self.nI, self.nU, self.rpp = rank, world, int(nI), int(nU), int(rpp)
        self.N = self.nI + self.nU
        self.users_r0, self.users_rows = int(users_r0), int(users_rows)
        self.it_r0 = min(rank * self.rpp, self.nI)
        self.it_rows = min(self.it_r0 + self.rpp, self.nI) - self.it_r0
        self.index_users, self.index_items = index_users, index_items
        self.dinv_users, self.dinv_items_all = dinv_users, dinv_items_all
        self.dinv_items_own = dinv_items_all[self.it_r0: self.it_r0 + self.it_rows]
        self.item_features_own, self.user_features_own = item_features_own, user_features_own
        self.arena, self.d_max, self.batch_max = arena, int(d_max), int(batch_max)
        self.edges_total, self.edges_own = int(edges_total), int(edges_own)
        self.off = self.layout(world, self.rpp, self.d_max, self.batch_max)
        self.device = arena.buf.device
        self._flag_ptrs = arena.ptrs(0)
        self._counters = arena.buf[_OFF_COUNTERS:].data_ptr()
        self._err = arena.view(_OFF_ERR, (1,), torch.int32)
        self.table_rows = world * self.rpp

    @staticmethod
    def layout(world, rpp, d_max, batch_max):
        off, cur = {}, _CTRL_BYTES
        slab = _align(world * rpp * d_max * 4)
        for name in ('recv0', 'recv1', 'T0', 'T1'):
            off[name] = cur
            cur += slab
        off['rows'] = cur
        cur += _align(2 * batch_max * d_max * 4)
        off['total'] = cur
        return off

    # ---- exchange primitives ---------------------------------------------------------------------------------------------
    def signal(self, channel):
        with torch.cuda.device(self.device):
            L.check(L.lib().b200rec_peer_signal(self._flag_ptrs, self.world, self.rank, channel, C.c_void_p(self._counters), _stream()), 'peer_signal')

    def wait(self, channel):
        with torch.cuda.device(self.device):
            L.check(L.lib().b200rec_peer_wait(C.c_void_p(self.arena.bases[self.rank]), self.world, channel, C.c_void_p(self._counters),
                                              WAIT_TIMEOUT_NS, _ptr(self._err), _stream()), 'peer_wait')

    def check(self):
        """raises if a wait of this rank ever timed out (call after a synchronize)"""
        e = int(self._err.item())
        if e:
            raise L.B200RecError(f'peer exchange: rank {self.rank} timed out waiting for rank {e - 1}')

    def table(self, par, d, dtype):
        """local copy of the transformed item features of ALL items: (P * rpp, d)"""
        return self.arena.view(self.off[f'T{par}'], (self.table_rows, d), dtype)

    def push_spec(self, par, d):
        """(destination arenas, parts, rows per part, element offset of MY receive slot, leading dimension) for K3's epilogue"""
        return (self.arena.ptrs(self.off[f'recv{par}']), self.world, self.rpp, self.rank * self.rpp * d, d)

    def reduce(self, par, d, *, x_next, acc_in, acc_out, acc_scale):
        if self.it_rows == 0:
            return
        recv = self.arena.view(self.off[f'recv{par}'], (self.world * self.rpp, d))
        with torch.cuda.device(self.device), ops._timed('peer_reduce', (self.it_rows, d, self.world)):
            L.check(L.lib().b200rec_peer_reduce(_ptr(recv), self.world, self.rpp * d, d, self.it_rows, d, _ptr(x_next), d if x_next is not None else 0,
                                               _ptr(acc_in), _ptr(acc_out), d, float(acc_scale), _stream()), 'peer_reduce')

    def push_transform(self, x, lin, par, t_dtype):
        """T[par][it_r0 : it_r0 + rows] = dinv ∘ (W x + b) on EVERY rank — K1c with the all-gather in its epilogue.
        `lin` = a Linear module or a (weight, bias) pair."""
        M, K = x.shape
        if M == 0:
            return
        w, b = lin if isinstance(lin, tuple) else (lin.weight, lin.bias)
        N = w.shape[0]
        esz = 2 if t_dtype == torch.bfloat16 else 4
        dst = self.arena.ptrs(self.off[f'T{par}'], first=self.rank)
        y_off = self.it_r0 * N
        if K <= 128 and K % 32 == 0 and N <= 128 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0 and x.stride(1) == 1:
            wc, ldw = ops._row_major(w)
            packed = ops._packed_weight(wc, ldw, L.TC_TF32X3)
            bias = None if b is None else b.detach().contiguous().float()
            with torch.cuda.device(self.device), ops._timed('linear_shortk_push', (M, K, N, self.world)):
                L.check(L.lib().b200rec_linear_shortk_push(_ptr(x), M, K, x.stride(0), _ptr(packed), N, _ptr(bias), _ptr(self.dinv_items_own), 0,
                                                          dst, self.world, y_off, N, ops._dtype_code(t_dtype), _stream()), 'linear_shortk_push')
            return
        # widths the persistent kernel does not take: plain GEMM into the local table, then a copy kernel to the peers
        local = self.table(par, N, t_dtype)[self.it_r0: self.it_r0 + M]
        ops.linear_raw(x, w, b, row_scale=self.dinv_items_own, out=local)
        if self.world > 1:
            if (N * esz) % 16:
                raise NotImplementedError('the peer path needs rows of the message table that are multiples of 16 bytes')
            n32 = N * esz // 4                          # the copy kernel moves 128-bit words: a bf16 row is N/2 32-bit columns
            others = (C.c_void_p * (self.world - 1))(*[self.arena.bases[q] + self.off[f'T{par}'] for q in range(self.world) if q != self.rank])
            with torch.cuda.device(self.device):
                L.check(L.lib().b200rec_peer_push_rows(_ptr(local), n32, M, n32, others, self.world - 1, self.it_r0 * n32, n32, _stream()), 'peer_push_rows')

    def gather_rows(self, table, row0, rows, ids, dst_row, d, scale=1.0):
        dst = self.arena.ptrs(self.off['rows'])
        with torch.cuda.device(self.device):
            L.check(L.lib().b200rec_peer_gather_rows(_ptr(table) if rows else None, d, row0, rows, _ptr(ids), ids.numel(), d, float(scale), dst,
                                                    self.world, dst_row * d, d, _stream()), 'peer_gather_rows')

    # ---- device-side trace (B200REC_PEER_TRACE=1): %globaltimer markers between the kernels of a forward, readable after a graph replay
    def stamp(self, tag):
        if not TRACE:
            return
        if getattr(self, '_trace_buf', None) is None:
            self._trace_buf = torch.zeros(256, dtype=torch.int64, device=self.device)
            self._trace_tags = []
        if self._trace_reset:
            self._trace_tags, self._trace_reset = [], False
        k = len(self._trace_tags)
        if k >= 256:
            return
        self._trace_tags.append(tag)
        with torch.cuda.device(self.device):
            L.check(L.lib().b200rec_device_timestamp(C.c_void_p(self._trace_buf.data_ptr() + 8 * k), _stream()), 'device_timestamp')

    _trace_reset = True

    def trace(self):
        """[(tag, microseconds since the first marker)] of the last forward (after a synchronize)"""
        if getattr(self, '_trace_buf', None) is None:
            return []
        t = self._trace_buf[:len(self._trace_tags)].cpu().tolist()
        return [(tag, round((x - t[0]) / 1e3, 1)) for tag, x in zip(self._trace_tags, t)]

    def side_stream(self):
        if getattr(self, '_side', None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def rows_view(self, n, d):
        return self.arena.view(self.off['rows'], (n, d))


def _rows_per_part(nI, world):
    return max(1, (int(nI) + world - 1) // world)


def shard_from_full(graph, rank, world, arena, d_max, batch_max=8192):
    """Shard of rank `rank` cut out of the full neighbour index (every rank builds the ~10 ms full index and keeps its slices;
    graphs that do not fit one GPU are built per rank instead: `bench.py --workload graph5`)."""
    from .parallel import RowPartition, _LocalIndex, column_slice_csr, split_rows
    full = get_index(graph)
    nI, N = int(graph.item_features.shape[0]), full.num_nodes
    rp = full.row_ptr
    k_items = int(rp[nI])
    users = RowPartition(split_rows(rp[nI:] - k_items, world), rank)
    k0, k1 = int(rp[nI + users.r0]), int(rp[nI + users.r1])
    cut = lambda a: None if a is None else a[k0:k1].contiguous()
    dinv_users = full.dinv[nI + users.r0: nI + users.r1].contiguous()
    dinv_items = full.dinv[:nI].contiguous()
    index_users = _LocalIndex((rp[nI + users.r0: nI + users.r1 + 1] - k0).contiguous(), full.col[k0:k1].contiguous(), cut(full.w), cut(full.pos),
                              dinv_users, full.chunk_size)
    head = lambda a: None if a is None else a[:k_items]
    lrp, lcol, lw, lpos = column_slice_csr(rp[:nI + 1], full.col[:k_items], nI + users.r0, nI + users.r1, head(full.w), head(full.pos))
    index_items = _LocalIndex(lrp, lcol, lw, lpos, dinv_items, full.chunk_size)
    rpp = _rows_per_part(nI, world)
    i0 = min(rank * rpp, nI)
    return PeerShard(rank=rank, world=world, nI=nI, nU=N - nI, d_max=d_max, rpp=rpp, users_r0=users.r0, users_rows=users.rows,
                     index_users=index_users, index_items=index_items, dinv_users=dinv_users, dinv_items_all=dinv_items,
                     item_features_own=graph.item_features[i0: min(i0 + rpp, nI)], user_features_own=graph.user_features[users.r0:users.r1],
                     arena=arena, batch_max=batch_max, edges_total=full.e1 + full.e2, edges_own=int(index_users.col.numel() + lcol.numel()))


def partition_graph_peer(graph, group=None, d_max=128, batch_max=8192):
    """One process per GPU: allocates + exchanges the arenas and cuts this rank's shard.  `d_max` = largest node_emb served."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    nI = int(graph.item_features.shape[0])
    nbytes = PeerShard.layout(world, _rows_per_part(nI, world), d_max, batch_max)['total']
    arena = PeerArena.ipc(nbytes, group, device=graph.item_features.device)
    return shard_from_full(graph, dist.get_rank(group), world, arena, d_max, batch_max)


def emulated_shards(graph, world, d_max=128, batch_max=8192):
    """all P shards in ONE process (single-GPU tests of the exchange logic; run with `forward_emulated`)"""
    nI = int(graph.item_features.shape[0])
    nbytes = PeerShard.layout(world, _rows_per_part(nI, world), d_max, batch_max)['total']
    arenas = PeerArena.emulated(nbytes, world, graph.item_features.device)
    return [shard_from_full(graph, q, world, arenas[q], d_max, batch_max) for q in range(world)]


# ----------------------------------------------------------------------------------------------------------------------
def _steps(model, sh: PeerShard, userIds, itemIds, keep=None):
    """Generator: GraphNCF.forward on shard `sh` (inference).  It yields right before every cross-rank wait, so that an emulation can
    advance all ranks in lock step; a real rank just runs it to the end.  Returns the (B, 1) scores."""
    from .neural_collaborative_filtering.util import run_mlp
    if torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters()):
        raise NotImplementedError('partitioned GraphNCF propagation is inference-only; wrap the call in torch.no_grad()')
    if model.concat or model.convType != 'LightGCN':
        raise NotImplementedError('the partitioned path covers LightGCN with mean combine (concat / LightGAT: single GPU)')
    dev = sh.device
    L_ = len(model.gnn_convs)
    d = model.item_embeddings[0].weight.shape[0]
    if d > sh.d_max or d % 4:
        raise ValueError(f'node_emb {d} does not fit the arena of this partition (d_max {sh.d_max}) or is not a multiple of 4')
    B = int(userIds.shape[0])
    if 2 * B > 2 * sh.batch_max:
        raise ValueError(f'batch of {B} pairs exceeds the partition\'s batch_max {sh.batch_max}')
    ie, ue = model.item_embeddings[0], model.user_embeddings[0]
    nu, ni = sh.users_rows, sh.it_rows
    iid, uid = itemIds.long().contiguous(), userIds.long().contiguous()
    empty = lambda n: torch.empty((n, d), dtype=torch.float32, device=dev)
    sh._trace_reset = True
    sh.stamp('start')
    x0_items, x0_users = empty(ni), empty(nu)

    def embed():                                                                     # x0 = node embeddings of the owned rows (:300-301)
        if ni:
            ops.linear_raw(sh.item_features_own, ie.weight, ie.bias, out=x0_items)
        if nu:
            ops.linear_raw(sh.user_features_own, ue.weight, ue.bias, out=x0_users)

    acc_items, acc_users = x0_items, x0_users
    if L_ == 0:
        embed()
    if L_ > 0:
        lin_u, lin_i, _ = model.gnn_convs[0].typed()
        if model.message_dtype not in ('fp32', 'bf16'):
            raise ValueError("message_dtype must be 'fp32' or 'bf16'")
        t_dtype = torch.bfloat16 if (model.message_dtype == 'bf16' and d % 32 == 0 and d <= 128) else torch.float32
        t_users = torch.empty((max(nu, 1), d), dtype=t_dtype, device=dev)
        acc_users, acc_items = empty(nu), empty(ni)
        spare_u = empty(nu) if L_ > 1 else None
        xi_next = empty(ni) if L_ > 1 else None
        main = torch.cuda.current_stream(dev)
        side = sh.side_stream() if OVERLAP else None
        joined = None
        # Layer 0's messages straight from the node features: t0 = dinv ∘ (W_t (W_e f + b_e) + b_t) = dinv ∘ ((W_t W_e) f + (W_t b_e + b_t)) — composite
        # weights formed once per weight version in float64.  So the first SpMM waits for ONE transform only; the embeddings x0 (needed from the
        # first running-mean update on) and the broadcast of the item messages run on the side stream underneath it.
        (Wci, bci), (Wcu, bcu) = _layer0_composites(model, ie, ue, lin_i, lin_u)
        if nu:
            ops.linear_raw(sh.user_features_own, Wcu, bcu, row_scale=sh.dinv_users, out=t_users[:nu])

        def layer0_side():
            sh.push_transform(sh.item_features_own, (Wci, bci), 0, t_dtype)
            sh.signal(CH_T0)
            embed()

        x0_ready = None
        if side is not None:
            start = torch.cuda.Event()
            start.record(main)
            side.wait_event(start)
            with torch.cuda.stream(side):
                layer0_side()
                x0_ready = torch.cuda.Event()
                x0_ready.record(side)
        else:
            layer0_side()

        def item_side(l, par, last, scale):                                          # C_l: needs every rank's A_l, nothing of B_l
            sh.stamp(f'C{l} begin')
            sh.wait(CH_A0 + par)
            sh.stamp(f'C{l} partials arrived')
            sh.reduce(par, d, x_next=None if last else xi_next, acc_in=x0_items if l == 0 else acc_items, acc_out=acc_items, acc_scale=scale)
            if not last:
                sh.push_transform(xi_next, lin_i, 1 - par, t_dtype)
                sh.signal(CH_T0 + (1 - par))
            sh.stamp(f'C{l} end')

        for l in range(L_):
            par, last = l & 1, l == L_ - 1
            scale = 1.0 / (L_ + 1) if last else 1.0
            sh.stamp(f'A{l} begin')
            if nu:                                                                   # A_l: partial item rows -> owners' receive slots
                ops.propagate_step(sh.index_items, t_users, dinv=sh.dinv_items_all, push=sh.push_spec(par, d))
            sh.signal(CH_A0 + par)
            sh.stamp(f'A{l} end')
            if side is not None:
                forked = torch.cuda.Event()
                forked.record(main)
            yield
            if side is not None:                                                     # C_l on the side stream, under B_l
                side.wait_event(forked)
                with torch.cuda.stream(side):
                    item_side(l, par, last, scale)
                    joined = torch.cuda.Event()
                    joined.record(side)
            sh.wait(CH_T0 + par)
            if l == 0 and x0_ready is not None:
                main.wait_event(x0_ready)                                            # x0_users is B_0's running-mean input
            sh.stamp(f'B{l} table arrived')
            if nu:                                                                   # B_l: own user rows from the gathered table
                ops.propagate_step(sh.index_users, sh.table(par, d, t_dtype), dinv=sh.dinv_users, x_next=None if last else spare_u,
                                   acc_in=x0_users if l == 0 else acc_users, acc_out=acc_users, acc_scale=scale)
            if side is None:
                yield
                item_side(l, par, last, scale)
            sh.stamp(f'B{l} end')
            if not last and nu:
                ops.linear_raw(spare_u, lin_u.weight, lin_u.bias, row_scale=sh.dinv_users, out=t_users[:nu])
        if joined is not None:
            main.wait_event(joined)
        sh.stamp('layers joined')
    if keep is not None:                        # tests: the owned rows of the combined embedding
        keep['items'], keep['users'] = acc_items, acc_users
    sh.gather_rows(acc_items, sh.it_r0, ni, iid, 0, d)
    sh.gather_rows(acc_users, sh.nI + sh.users_r0, nu, uid, B, d)
    sh.signal(CH_R)
    yield
    sh.wait(CH_R)
    sh.stamp('batch rows arrived')
    rows = sh.rows_view(2 * B, d)
    if model.MLP is None:
        out = ops.rowdot(rows[B:], rows[:B])
    else:
        out = run_mlp(model.MLP, rows[:B], rows[B:], training=False)               # item first (gnn_ncf.py:361)
    sh.stamp('end')
    if L_ == 0:                                 # no layer barrier protected `rows` against the next call's writers
        sh.signal(CH_E)
        yield
        sh.wait(CH_E)
    return out


def _layer0_composites(model, ie, ue, lin_i, lin_u):
    """((W_i2u·W_item, W_i2u·b_item + b_i2u), (W_u2i·W_user, W_u2i·b_user + b_u2i)) as cached fp32 tensors (float64 products, rounded once)"""
    ps = (ie.weight, ie.bias, ue.weight, ue.bias, lin_i.weight, lin_i.bias, lin_u.weight, lin_u.bias)
    key = tuple((p.data_ptr(), p._version) for p in ps if p is not None)
    hit = getattr(model, '_peer_layer0', None)
    if hit is not None and hit[0] == key:
        return hit[1]

    def comp(lin, emb):
        w = (lin.weight.detach().double() @ emb.weight.detach().double()).float().contiguous()
        eb = emb.bias.detach().double() if emb.bias is not None else torch.zeros(emb.weight.shape[0], dtype=torch.float64, device=w.device)
        b = lin.weight.detach().double() @ eb
        if lin.bias is not None:
            b = b + lin.bias.detach().double()
        return w, b.float().contiguous()

    out = (comp(lin_i, ie), comp(lin_u, ue))
    object.__setattr__(model, '_peer_layer0', (key, out))
    return out


def forward_peer(model, sh: PeerShard, userIds, itemIds, keep=None):
    gen = _steps(model, sh, userIds, itemIds, keep)
    try:
        while True:
            next(gen)
    except StopIteration as stop:
        return stop.value


def forward_emulated(model, shards, userIds, itemIds, keep=None):
    """all ranks of `emulated_shards` in lock step on one stream; returns the list of per-rank outputs"""
    keep = keep if keep is not None else [None] * len(shards)
    gens = [_steps(model, sh, userIds, itemIds, keep[q]) for q, sh in enumerate(shards)]
    outs = [None] * len(shards)
    live = list(range(len(shards)))
    while live:
        for q in list(live):
            try:
                next(gens[q])
            except StopIteration as stop:
                outs[q] = stop.value
                live.remove(q)
    return outs
