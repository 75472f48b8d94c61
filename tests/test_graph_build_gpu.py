"""K4: device-side create_graph / CSR build, bit-exact against the reference's golden output and the oracle."""
import numpy as np
import pytest
import torch

from oracle import restatement as R
from oracle import synth
from tests._golden import load

pytestmark = pytest.mark.gpu


def _build(d, binary, dev):
    from deeprecommendation_b200.graph import IdTable, create_graph
    ut = IdTable(torch.from_numpy(d['all_users']).to(dev))
    it = IdTable(torch.from_numpy(d['all_items']).to(dev))
    g = create_graph(torch.from_numpy(d['user_raw']).to(dev), torch.from_numpy(d['item_raw']).to(dev),
                     torch.from_numpy(d['rating']).to(dev), None, None, ut, it, binary=binary)
    return g, ut, it


@pytest.mark.parametrize('binary', [0, 1])
def test_create_graph_vs_reference_golden(binary):
    dev = torch.device('cuda:0')
    d, _, _ = load(f'graph_build_binary{binary}')
    g, ut, it = _build(d, bool(binary), dev)
    assert np.array_equal(ut.sorted().cpu().numpy(), np.unique(d['all_users']))
    assert np.array_equal(it.sorted().cpu().numpy(), np.unique(d['all_items']))
    assert g.user2item_edge_index.dtype == torch.int64
    assert np.array_equal(g.user2item_edge_index.cpu().numpy(), d['user2item_edge_index'])
    assert np.array_equal(g.item2user_edge_index.cpu().numpy(), d['item2user_edge_index'])
    if not binary:
        assert np.array_equal(g.user2item_edge_attr.cpu().numpy().view(np.int32), d['user2item_edge_attr'].view(np.int32))
        assert np.array_equal(g.item2user_edge_attr.cpu().numpy().view(np.int32), d['item2user_edge_attr'].view(np.int32))
    else:
        assert g.user2item_edge_attr is None and g.item2user_edge_attr is None


def test_unknown_id_raises_keyerror():
    from deeprecommendation_b200.graph import IdTable
    dev = torch.device('cuda:0')
    t = IdTable(torch.tensor([5, 9, 9, 2], device=dev))
    assert t.count == 3 and t.lookup(torch.tensor([9, 2], device=dev), offset=10).tolist() == [12, 10]
    with pytest.raises(KeyError):
        t.lookup(torch.tensor([7], device=dev))


@pytest.mark.parametrize('n_users,n_items,n', [(300, 500, 20_000), (5000, 3000, 400_000)])
def test_index_vs_oracle(n_users, n_items, n):
    """edge lists, attrs, CSR (row_ptr / col / weights / positions), degrees and deg^-1/2: all bit-exact"""
    from deeprecommendation_b200.graph import IdTable, create_graph, GraphIndex
    dev = torch.device('cuda:0')
    users, items, ratings = synth.interactions_zipf(n_users, n_items, n, seed=3)
    users_raw, items_raw = users * 3 + 1, items * 2 + 5
    us, its = R.node_ids(users_raw, items_raw)
    ref = R.create_graph(users_raw, items_raw, ratings, us, its, binary=False)
    ut, it = IdTable(torch.from_numpy(users_raw).to(dev)), IdTable(torch.from_numpy(items_raw).to(dev))
    g = create_graph(torch.from_numpy(users_raw).to(dev), torch.from_numpy(items_raw).to(dev), torch.from_numpy(ratings).to(dev),
                     None, None, ut, it)
    for k in ('user2item_edge_index', 'item2user_edge_index'):
        assert np.array_equal(getattr(g, k).cpu().numpy(), ref[k])
    for k in ('user2item_edge_attr', 'item2user_edge_attr'):
        assert np.array_equal(getattr(g, k).cpu().numpy().view(np.int32), ref[k].view(np.int32))
    N = len(us) + len(its)
    idx = GraphIndex(g.user2item_edge_index, g.item2user_edge_index, g.user2item_edge_attr, g.item2user_edge_attr, N, chunk=64)
    total = np.concatenate([ref['user2item_edge_index'], ref['item2user_edge_index']], axis=1)
    row_ptr, src, perm = R.csr_by_destination(total, N)
    assert np.array_equal(idx.row_ptr.cpu().numpy(), row_ptr.astype(np.int32))
    assert np.array_equal(idx.col.cpu().numpy(), src.astype(np.int32))
    e1 = ref['user2item_edge_index'].shape[1]
    assert np.array_equal(idx.pos.cpu().numpy(), np.where(perm < e1, perm, perm - e1).astype(np.int32))
    w_all = np.concatenate([ref['user2item_edge_attr'], ref['item2user_edge_attr']])
    assert np.array_equal(idx.w.cpu().numpy().view(np.int32), w_all[perm].view(np.int32))
    deg = np.diff(row_ptr)
    assert np.array_equal(idx.deg.cpu().numpy(), deg.astype(np.int32))
    dinv = torch.from_numpy(deg.astype(np.float32)).pow(-0.5)
    dinv[dinv == float('inf')] = 0
    assert np.array_equal(idx.dinv.cpu().numpy().view(np.int32), dinv.numpy().view(np.int32))     # gnn_ncf.py:48-50, bit-exact
    # chunk plan covers every row exactly once, in order
    cr, cs, sl = idx.chunk_row.cpu().numpy(), idx.chunk_start.cpu().numpy(), idx.chunk_slot.cpu().numpy()
    assert idx.n_chunks == np.maximum(1, -(-deg // 64)).sum()
    assert np.array_equal(np.unique(cr), np.arange(N))
    first = np.r_[True, cr[1:] != cr[:-1]]
    assert np.array_equal(cs[first], row_ptr[:-1][cr[first]])
    assert (sl >= 0).sum() == idx.n_slots and idx.n_multi == (deg > 64).sum()
    # position lookup == pos_df
    pick = np.random.default_rng(0).permutation(n)[:500]
    u_nodes = torch.from_numpy(ref['user2item_edge_index'][0][pick]).to(dev)
    i_nodes = torch.from_numpy(ref['user2item_edge_index'][1][pick]).to(dev)
    assert np.array_equal(idx.positions(u_nodes, i_nodes).cpu().numpy(), pick)
    with pytest.raises(KeyError):                       # a pair that is not an edge (gnn_ncf.py:370 raises KeyError)
        idx.positions(torch.tensor([len(its)], device=dev), torch.tensor([N + 5], device=dev))
    # target masking: bitmap + degree adjustment, duplicates counted once
    pos = torch.from_numpy(np.r_[pick[:50], pick[:10]]).to(dev)
    skip, dinv_m = idx.masked(pos)
    bits = np.zeros(n, dtype=bool)
    bits[pick[:50]] = True
    got = np.unpackbits(skip.cpu().numpy().view(np.uint8), bitorder='little')[:n].astype(bool)
    assert np.array_equal(got, bits)
    deg_m = deg.copy()
    np.subtract.at(deg_m, ref['user2item_edge_index'][1][pick[:50]], 1)
    np.subtract.at(deg_m, ref['item2user_edge_index'][1][pick[:50]], 1)
    ref_dinv = torch.from_numpy(deg_m.astype(np.float32)).pow(-0.5)
    ref_dinv[ref_dinv == float('inf')] = 0
    assert np.array_equal(dinv_m.cpu().numpy().view(np.int32), ref_dinv.numpy().view(np.int32))
