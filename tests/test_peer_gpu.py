"""Scheme 'peer' (deeprecommendation_b200/peer.py, csrc/peer.cu): the partitioned GraphNCF propagation whose exchange is written as
this library's own kernels over peer-mapped memory — K3 pushing partial rows to their owner, the slot reduction, K1c broadcasting
the transformed rows, epoch flags.  On ONE GPU: (1) P emulated ranks inside one process advance in lock step over the same kernels
and addresses; (2) two real processes share cuda:0 and exchange CUDA IPC handles over gloo (tests/mp_peer_check.py).
Reference = the single-GPU forward (itself pinned to the oracle / reference goldens in tests/test_models_gpu.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import synth
from tests._golden import maxnorm_rel

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _graph(n_users=1500, n_items=900, n=60_000, F=48, d_emb=64, L_=2, mlp=(128,), dot=False, seed=13):
    from deeprecommendation_b200.graph import IdTable, create_graph
    from deeprecommendation_b200.neural_collaborative_filtering.models import GraphNCF
    users, items, ratings = synth.interactions_zipf(n_users, n_items, n, seed=seed)
    rng = np.random.default_rng(2)
    fi, fu = rng.standard_normal((n_items, F)).astype(np.float32), rng.standard_normal((n_users, F)).astype(np.float32)
    kw = dict(item_dim=F, user_dim=F, num_gnn_layers=L_, hetero=True, node_emb=d_emb, mlp_dense_layers=list(mlp), dropout_rate=0.2,
              use_dot_product=dot)
    sd = synth.to_torch(synth.graph_ncf_weights(seed=3, **kw))
    g = create_graph(torch.from_numpy(users).to(DEV), torch.from_numpy(items).to(DEV), torch.from_numpy(ratings).to(DEV),
                     torch.from_numpy(fi).to(DEV), torch.from_numpy(fu).to(DEV),
                     IdTable(torch.arange(n_users, device=DEV)), IdTable(torch.arange(n_items, device=DEV)))
    m = GraphNCF(**kw).to(DEV).eval()
    m.load_state_dict(sd)
    pick = rng.permutation(n)[:300]
    return m, g, g.user2item_edge_index[0][pick].contiguous(), g.user2item_edge_index[1][pick].contiguous()


def _check_world(m, g, uid, iid, P, tol=1e-6, d_max=128):
    from deeprecommendation_b200.graph import get_index
    from deeprecommendation_b200.peer import emulated_shards, forward_emulated
    full = get_index(g)
    nI = g.item_features.shape[0]
    with torch.no_grad():
        ref = m(g, uid, iid, DEV)
        comb = m._encode(g, full, None, full.dinv, False)
        shards = emulated_shards(g, P, d_max=d_max, batch_max=512)
        assert sum(sh.edges_own for sh in shards) == full.e1 + full.e2
        assert sum(sh.users_rows for sh in shards) == full.num_nodes - nI and sum(sh.it_rows for sh in shards) == nI
        for rep in range(2):                         # second pass: the device-side epoch counters advance (CUDA-graph replay contract)
            keep = [dict() for _ in range(P)]
            outs = forward_emulated(m, shards, uid, iid, keep)
            torch.cuda.synchronize()
            for sh, out, k in zip(shards, outs, keep):
                sh.check()
                assert maxnorm_rel(out, ref) < tol, (P, sh.rank, rep)
                assert maxnorm_rel(k['items'], comb[sh.it_r0: sh.it_r0 + sh.it_rows]) < tol if sh.it_rows else True
                assert maxnorm_rel(k['users'], comb[nI + sh.users_r0: nI + sh.users_r0 + sh.users_rows]) < tol if sh.users_rows else True
            assert all(torch.equal(outs[0], o) for o in outs[1:])       # every rank ends with the same bits
    return outs[0], ref


@pytest.mark.parametrize('P', [1, 2, 3, 8])
def test_peer_emulated_ranks_match_single_gpu(P):
    m, g, uid, iid = _graph()
    _check_world(m, g, uid, iid, P)


@pytest.mark.parametrize('d_emb,L_,mlp', [(128, 2, (256, 128)), (64, 3, (128,)), (32, 1, (64,)), (64, 0, (128,))])
def test_peer_emulated_shapes(d_emb, L_, mlp):
    m, g, uid, iid = _graph(n_users=2500, n_items=1200, n=90_000, d_emb=d_emb, L_=L_, mlp=mlp, seed=5)
    _check_world(m, g, uid, iid, 4)


def test_peer_emulated_dot_product_and_more_ranks_than_items_rows():
    m, g, uid, iid = _graph(n_users=400, n_items=11, n=2500, d_emb=64, L_=2, dot=True, seed=7)      # 11 items over 8 ranks: empty item shards
    _check_world(m, g, uid, iid, 8)


def test_peer_emulated_width_outside_the_persistent_gemm():
    """node_emb = 48 is not a shape of K1c: the transform falls back to the plain GEMM + the copy kernel to the peers"""
    m, g, uid, iid = _graph(d_emb=48, L_=2, seed=3)
    _check_world(m, g, uid, iid, 3)


def test_peer_emulated_bf16_messages():
    m, g, uid, iid = _graph(d_emb=128, L_=2, mlp=(256, 128))
    with torch.no_grad():
        ref = m(g, uid, iid, DEV)
    m.message_dtype = 'bf16'
    try:
        out, _ = _check_world(m, g, uid, iid, 4, tol=1e-2)
    finally:
        m.message_dtype = 'fp32'
    assert maxnorm_rel(out, ref) < 1e-2 and not torch.equal(out, ref)


def test_peer_push_epilogue_equals_local_rows():
    """K3 with `push`: every finished row lands in slot (rank) of its owner's receive buffer, bit-equal to the x_next the plain
    epilogue writes (single- and multi-chunk rows)."""
    from deeprecommendation_b200 import ops
    from deeprecommendation_b200.graph import get_index
    from deeprecommendation_b200.peer import emulated_shards
    m, g, uid, iid = _graph(n_users=3000, n_items=700, n=120_000, d_emb=64, seed=11)
    P, d = 4, 64
    shards = emulated_shards(g, P, d_max=d, batch_max=64)
    t = torch.randn(3000, d, device=DEV)
    for sh in shards:
        tu = t[sh.users_r0: sh.users_r0 + sh.users_rows].contiguous()
        ref = torch.zeros(sh.nI, d, device=DEV)
        ops.spmm_raw(sh.index_items, tu, w=sh.index_items.w, dinv=sh.dinv_items_all, x_next=ref)
        assert sh.index_items.n_multi > 0                                   # the fix-up kernel's push path is exercised
        ops.spmm_raw(sh.index_items, tu, w=sh.index_items.w, dinv=sh.dinv_items_all, push=sh.push_spec(1, d))
        torch.cuda.synchronize()
        has_edges = (sh.index_items.row_ptr[1:] > sh.index_items.row_ptr[:-1])
        for o, owner in enumerate(shards):
            recv = owner.arena.view(owner.off['recv1'], (P, owner.rpp, d))[sh.rank][:owner.it_rows]
            want = ref[owner.it_r0: owner.it_r0 + owner.it_rows]
            rows = has_edges[owner.it_r0: owner.it_r0 + owner.it_rows]
            assert torch.equal(recv[rows], want[rows])
            assert torch.all(recv[~rows] == 0)                               # never written: the zero-initialised arena


def test_peer_wait_times_out_into_error_flag_instead_of_hanging():
    from deeprecommendation_b200 import peer
    m, g, uid, iid = _graph(n_users=200, n_items=50, n=2000)
    sh = peer.emulated_shards(g, 2, d_max=64, batch_max=16)[0]
    old = peer.WAIT_TIMEOUT_NS
    peer.WAIT_TIMEOUT_NS = 20_000_000            # 20 ms
    try:
        sh.wait(peer.CH_E)                       # nobody signals channel E
        torch.cuda.synchronize()
    finally:
        peer.WAIT_TIMEOUT_NS = old
    with pytest.raises(RuntimeError, match='timed out'):
        sh.check()


def test_peer_two_processes_share_one_gpu_over_cuda_ipc():
    """two real ranks (gloo rendezvous, both on cuda:0): arenas exchanged as CUDA IPC handles, flags crossing process boundaries"""
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT=str(29600 + os.getpid() % 300), B200REC_PEER_CHECK_SAME_GPU='1')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', env['MASTER_PORT'], os.path.join(ROOT, 'tests', 'mp_peer_check.py')],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert r.returncode == 0 and lines, (r.returncode, r.stdout[-2000:], r.stderr[-3000:])
    res = json.loads(lines[-1])
    assert res['ok'] and res['world'] == 2, res


def test_sharded_build_equals_shard_cut_from_the_full_index():
    """sharded.create_graph_local / shard_from_local (a rank sees only ITS users' interactions; item statistics made global by a
    reduction) against peer.shard_from_full (cut out of the full index): same CSR slices, weights, deg^-1/2 — and the same forward."""
    from deeprecommendation_b200.graph import IdTable, create_graph, get_index
    from deeprecommendation_b200.neural_collaborative_filtering.models import GraphNCF
    from deeprecommendation_b200.parallel import RowPartition, split_rows
    from deeprecommendation_b200.peer import PeerArena, PeerShard, _rows_per_part, emulated_shards, forward_emulated
    from deeprecommendation_b200.sharded import create_graph_local, shard_from_local
    n_users, n_items, n, F, d, P = 2000, 700, 80_000, 64, 64, 4
    users, items, ratings = synth.interactions_zipf(n_users, n_items, n, seed=17)
    order = np.argsort(users, kind='stable')                   # interactions grouped by user (a rank's file holds its users' rows)
    users, items, ratings = users[order], items[order], ratings[order]
    rng = np.random.default_rng(3)
    fi, fu = rng.standard_normal((n_items, F)).astype(np.float32), rng.standard_normal((n_users, F)).astype(np.float32)
    tu, ti, tr = torch.from_numpy(users).to(DEV), torch.from_numpy(items).to(DEV), torch.from_numpy(ratings).to(DEV)
    g = create_graph(tu, ti, tr, torch.from_numpy(fi).to(DEV), torch.from_numpy(fu).to(DEV), IdTable(torch.arange(n_users, device=DEV)),
                     IdTable(torch.arange(n_items, device=DEV)))
    full = get_index(g)
    ref_shards = emulated_shards(g, P, d_max=d, batch_max=512)
    # the same user ranges, built rank by rank from the rank's own interactions
    glob = {'cnt': torch.bincount(ti, minlength=n_items).int(), 'sum': torch.zeros(n_items, dtype=torch.float64, device=DEV).index_add_(0, ti, tr.double()),
            'deg': full.deg[:n_items].clone()}
    reduce_items = lambda t, what: t.copy_(glob[what])
    arenas = PeerArena.emulated(PeerShard.layout(P, _rows_per_part(n_items, P), d, 512)['total'], P, DEV)
    shards = []
    for q, ref in enumerate(ref_shards):
        mine = (tu >= ref.users_r0) & (tu < ref.users_r0 + ref.users_rows)
        gl = create_graph_local(tu[mine] - ref.users_r0, ti[mine], tr[mine], ref.users_rows, n_items, g.item_features,
                                g.user_features[ref.users_r0: ref.users_r0 + ref.users_rows], reduce_items)
        sh = shard_from_local(gl, rank=q, world=P, nI=n_items, nU=n_users, users_r0=ref.users_r0, arena=arenas[q], d_max=d, batch_max=512,
                              reduce_items=reduce_items, edges_total=full.e1 + full.e2)
        for a, b in ((sh.index_users, ref.index_users), (sh.index_items, ref.index_items)):
            assert torch.equal(a.row_ptr, b.row_ptr) and torch.equal(a.col, b.col) and torch.equal(a.w, b.w)
        assert torch.equal(sh.dinv_users, ref.dinv_users) and torch.equal(sh.dinv_items_all, ref.dinv_items_all)
        shards.append(sh)
    kw = dict(item_dim=F, user_dim=F, num_gnn_layers=3, hetero=True, node_emb=d, mlp_dense_layers=[128], dropout_rate=0.2)
    m = GraphNCF(**kw).to(DEV).eval()
    m.load_state_dict(synth.to_torch(synth.graph_ncf_weights(seed=3, **kw)))
    pick = rng.permutation(n)[:300]
    uid, iid = g.user2item_edge_index[0][pick].contiguous(), g.user2item_edge_index[1][pick].contiguous()
    with torch.no_grad():
        ref_out = m(g, uid, iid, DEV)
        outs = forward_emulated(m, shards, uid, iid)
    torch.cuda.synchronize()
    for sh, o in zip(shards, outs):
        sh.check()
        assert maxnorm_rel(o, ref_out) < 1e-6
