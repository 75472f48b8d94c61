"""CPU restatement of the reference's scoring hot path.  TEST INFRASTRUCTURE (oracle/__init__.py).

Every function follows the reference's own operation order on torch CPU tensors (fp32 arithmetic,
int64 indices), so that (a) it reproduces the reference's outputs on identical inputs — checked
against tests/golden/*.npz, which were produced by the unmodified reference (make_golden.py) — and
(b) timing it measures the reference's algorithm, not an optimised rewrite (`cpu_baseline.kind =
"port"` in bench.py).  Paths below are relative to /root/reference/src/.

Weights are passed as a plain `state_dict`-style mapping with the reference's key names
(SURVEY.md §8b), so shipped checkpoints can be fed directly.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------
# a-1  build_MLP_layers — neural_collaborative_filtering/util.py:5-18
# --------------------------------------------------------------------------------------------------
def mlp_linear_indices(sd, prefix: str = 'MLP.'):
    """Indices k of `MLP.{k}.weight` in ascending order (0,3,6 with dropout; 0,2,4 without)."""
    ks = sorted({int(k[len(prefix):].split('.')[0]) for k in sd if k.startswith(prefix) and k.endswith('.weight')})
    return ks


def mlp_forward(x: torch.Tensor, sd, prefix: str = 'MLP.') -> torch.Tensor:
    """Linear, then (ReLU, [Dropout inert in eval], Linear)* — util.py:13-17."""
    ks = mlp_linear_indices(sd, prefix)
    for n, k in enumerate(ks):
        if n > 0:
            x = torch.relu(x)
        x = F.linear(x, sd[f'{prefix}{k}.weight'], sd[f'{prefix}{k}.bias'])
    return x


# --------------------------------------------------------------------------------------------------
# a-2  BasicNCF.forward — models/basic_ncf.py:37-42
# --------------------------------------------------------------------------------------------------
def basic_ncf_forward(sd, X_user: torch.Tensor, X_item: torch.Tensor) -> torch.Tensor:
    user_emb = F.linear(X_user, sd['user_embeddings.0.weight'], sd['user_embeddings.0.bias'])   # :38
    item_emb = F.linear(X_item, sd['item_embeddings.0.weight'], sd['item_embeddings.0.bias'])   # :39
    combined = torch.cat((user_emb, item_emb), dim=1)                                           # :40 user first
    return mlp_forward(combined, sd)                                                            # :41


def basic_ncf_all_pairs(sd, X_users: torch.Tensor, X_items: torch.Tensor) -> torch.Tensor:
    """(nU, nI) scores: the reference forward on every (user, item) pair, users major — what BASELINE configs[3] and the
    webapp's score-every-candidate loop (webapp/backend.py:96-99) evaluate."""
    nU, nI = X_users.shape[0], X_items.shape[0]
    out = torch.empty((nU, nI), dtype=torch.float32)
    for u in range(nU):
        out[u] = basic_ncf_forward(sd, X_users[u:u + 1].expand(nI, -1), X_items).view(-1)
    return out


def topk_stable(scores: torch.Tensor, k: int, seen=None):
    """k best columns per row, descending, ties towards the lower column (`sort_values(by='score', ascending=False).iloc[:k]`,
    webapp/backend.py:113-121, pinned to a stable order); `seen[u]` = columns left out (`ignore_seen`, :85).
    Returns (values (nU, k) fp32, indices (nU, k) int64); rows with fewer than k candidates are padded with (-inf, -1)."""
    s = scores.detach().cpu().numpy().astype(np.float32).copy()
    nU, nI = s.shape
    val = np.full((nU, k), -np.inf, dtype=np.float32)
    idx = np.full((nU, k), -1, dtype=np.int64)
    for u in range(nU):
        cols = np.arange(nI)
        if seen is not None:
            cols = np.setdiff1d(cols, np.asarray(seen[u], dtype=np.int64))
        cols = cols[~np.isnan(s[u, cols])]
        order = cols[np.argsort(-s[u, cols], kind='stable')][:k]
        val[u, :len(order)] = s[u, order]
        idx[u, :len(order)] = order
    return torch.from_numpy(val), torch.from_numpy(idx)


# --------------------------------------------------------------------------------------------------
# a-3  AttentionNCF.forward — models/attention_ncf.py:136-224
# --------------------------------------------------------------------------------------------------
def attention_net(x: torch.Tensor, sd) -> torch.Tensor:
    """AttentionNet: Linear(2E, att_dense), ReLU, Dropout, Linear(att_dense, 1)  (:112-117), or a single
    Linear(2E, 1) when att_dense is None (:120-122)."""
    if 'AttentionNet.3.weight' in sd:
        h = torch.relu(F.linear(x, sd['AttentionNet.0.weight'], sd['AttentionNet.0.bias']))
        return F.linear(h, sd['AttentionNet.3.weight'], sd['AttentionNet.3.bias'])
    return F.linear(x, sd['AttentionNet.0.weight'], sd['AttentionNet.0.bias'])


def attention_ncf_forward(sd, candidate_items, rated_items, user_matrix, *, use_cos_sim_instead=False,
                          training=False, return_attention_weights=False):
    """Literal op order of attention_ncf.py:146-224 (eval mode, or train mode with every dropout inert)."""
    I = rated_items.shape[0]
    B = candidate_items.shape[0]
    cand_emb = F.linear(candidate_items, sd['ItemEmbeddings.0.weight'], sd['ItemEmbeddings.0.bias'])   # :150
    rated_emb = F.linear(rated_items, sd['ItemEmbeddings.0.weight'], sd['ItemEmbeddings.0.bias'])      # :151
    cand_full = cand_emb.repeat_interleave(I, dim=0)                                                   # :154
    rated_full = rated_emb.repeat(B, 1)                                                                # :155
    nz = user_matrix != 0
    cand_sel = cand_full.view(B, I, -1)[nz]                                                            # :158
    rated_sel = rated_full.view(B, I, -1)[nz]                                                          # :159
    if use_cos_sim_instead:                                                                            # :162-173
        a = F.normalize(cand_sel, p=2, dim=1)
        b = F.normalize(rated_sel, p=2, dim=1)
        att_out = torch.bmm(a.unsqueeze(1), b.unsqueeze(2)).view(-1)
    else:                                                                                              # :176-179
        att_out = attention_net(torch.cat((cand_sel, rated_sel), dim=1), sd).view(-1)
    scores = -float('inf') * torch.ones((B, I), dtype=torch.float32)                                   # :182
    scores[nz] = att_out                                                                               # :192
    if training:                                                                                       # :195-203
        mask = torch.isclose(cand_full, rated_full, atol=1e-5).all(dim=1).view(B, I)
        scores[mask] = -float('inf')
    scores = F.softmax(scores, dim=1)                                                                  # :208
    scores = scores.nan_to_num(nan=0.0, posinf=0.0, neginf=0.0)                                        # :209
    attended = torch.mul(scores, user_matrix)                                                          # :212
    user_feat = torch.matmul(attended, rated_items)                                                    # :213
    user_emb = F.linear(user_feat, sd['UserEmbeddings.0.weight'], sd['UserEmbeddings.0.bias'])         # :216
    combined = torch.cat((cand_emb, user_emb), dim=1)                                                  # :219 item first
    out = mlp_forward(combined, sd)                                                                    # :222
    return (out, scores) if return_attention_weights else out


def attention_ncf_forward_blocked(sd, candidate_items, rated_items, user_matrix, block: int = 64, **kw):
    """Same arithmetic as `attention_ncf_forward`, evaluated `block` candidate rows at a time so that the
    (B*I, E) materialisation of attention_ncf.py:154-155 fits in host memory at large I.  Each row of the
    output depends only on its own row of `candidate_items`/`user_matrix`, so blocking does not change
    any value (asserted in tests/test_oracle.py)."""
    outs, atts = [], []
    want_att = kw.get('return_attention_weights', False)
    for s in range(0, candidate_items.shape[0], block):
        r = attention_ncf_forward(sd, candidate_items[s:s + block], rated_items, user_matrix[s:s + block], **kw)
        if want_att:
            outs.append(r[0]); atts.append(r[1])
        else:
            outs.append(r)
    return (torch.cat(outs), torch.cat(atts)) if want_att else torch.cat(outs)


# --------------------------------------------------------------------------------------------------
# f-1  gradient of the attention pooling (attention_ncf.py:154-216 under NCF/train.py:99-105), in the factorised form
#      the CUDA path computes (csrc/attention_pool.cu): s = scale·(a2·ReLU(Pc[b]+Pr[i]) + a20) or scale·<Pc[b],Pr[i]>,
#      alpha = softmax over the kept pairs, out = (alpha∘um)·Q + bU
# --------------------------------------------------------------------------------------------------
def attention_pool_factorised(Pc, Pr, Q, a2, a20, bU, um, keep, *, net=True, scale=1.0):
    """Dense restatement of the factorised forward; `keep` (B, I) bool = pairs that take part.  Returns (out, alpha)."""
    if net:
        s = (torch.relu(Pc[:, None, :] + Pr[None, :, :]) * a2).sum(-1) + a20
    else:
        s = Pc @ Pr.T
    s = (s * scale).masked_fill(~keep, float('-inf'))
    alpha = torch.nan_to_num(torch.softmax(s, dim=1), nan=0.0)          # rows with nothing kept -> 0 (:208-209)
    return (alpha * um) @ Q + bU, alpha


def attention_pool_backward(Pc, Pr, Q, a2, bU, um, alpha, out, g, *, net=True, scale=1.0):
    """Closed-form gradients of `attention_pool_factorised` from its outputs (what csrc/attention_pool_bwd.cu evaluates per
    non-zero): ds = scale·alpha·(um·<g[b],Q[i]> − <g[b], out[b]−bU>).  Returns dict(Pc, Pr, Q, a2, a20, bU)."""
    gq = g @ Q.T                                                         # (B, I): <g[b], Q[i]>
    gdot = (g * (out - bU)).sum(1, keepdim=True)                         # softmax row sum Σ_j alpha_bj·dalpha_bj
    ds = scale * alpha * (um * gq - gdot)
    grads = {'Q': (alpha * um).T @ g, 'bU': g.sum(0)}
    if net:
        z = Pc[:, None, :] + Pr[None, :, :]                              # (B, I, H)
        t = ds[:, :, None] * a2 * (z > 0)
        grads.update(Pc=t.sum(1), Pr=t.sum(0), a2=(ds[:, :, None] * torch.relu(z)).sum((0, 1)), a20=ds.sum())
    else:
        grads.update(Pc=ds @ Pr, Pr=ds.T @ Pc, a2=None, a20=None)
    return grads


# --------------------------------------------------------------------------------------------------
# a-4  LightGCNConv.forward / message — models/gnn_ncf.py:39-94 (+ PyG propagate/degree, pyg_shim)
# --------------------------------------------------------------------------------------------------
def _propagate_add(x, edge_index, norm, weight, W, b):
    """PyG propagate(aggr='add') with the message of gnn_ncf.py:74-94: per-edge Linear on x_j, scaled by
    weight*norm, scatter-added over edge_index[1]."""
    x_j = x.index_select(0, edge_index[0])
    t = F.linear(x_j, W, b)                                            # W(x_j): per-edge transform (:91,93)
    if weight is not None:
        msg = weight.view(-1, 1) * norm.view(-1, 1) * t               # :91
    else:
        msg = norm.view(-1, 1) * t                                    # :93
    out = torch.zeros((x.size(0), t.size(1)), dtype=t.dtype)
    return out.index_add_(0, edge_index[1], msg)


def lightgcn_conv(x, u2i_index, i2u_index, u2i_attr, i2u_attr, sd, prefix, hetero=True):
    total = torch.cat([u2i_index, i2u_index], dim=1)                                   # :41
    to_ = total[1]
    deg = torch.zeros(x.size(0), dtype=x.dtype).scatter_add_(0, to_, torch.ones(to_.size(0), dtype=x.dtype))  # :48
    dinv = deg.pow(-0.5)                                                               # :49
    dinv[dinv == float('inf')] = 0                                                     # :50
    if hetero:
        n1 = dinv[u2i_index[0]] * dinv[u2i_index[1]]                                   # :54
        o1 = _propagate_add(x, u2i_index, n1, u2i_attr,
                            sd[f'{prefix}user2item_W.0.weight'], sd[f'{prefix}user2item_W.0.bias'])   # :55
        n2 = dinv[i2u_index[0]] * dinv[i2u_index[1]]                                   # :58
        o2 = _propagate_add(x, i2u_index, n2, i2u_attr,
                            sd[f'{prefix}item2user_W.0.weight'], sd[f'{prefix}item2user_W.0.bias'])   # :59
        return o1 + o2                                                                 # :62
    attr = torch.cat([u2i_attr, i2u_attr]) if (u2i_attr is not None and i2u_attr is not None) else None  # :43-44
    norm = dinv[total[0]] * dinv[total[1]]                                             # :66
    return _propagate_add(x, total, norm, attr, sd[f'{prefix}W.0.weight'], sd[f'{prefix}W.0.bias'])     # :69


# --------------------------------------------------------------------------------------------------
# LightGATConv.forward / message — models/gnn_ncf.py:128-177 (queued as NEXT)
# --------------------------------------------------------------------------------------------------
def _segment_softmax(src, index, n):
    """PyG 2.0.4 utils.softmax: exp(src - groupmax) / (groupsum + 1e-16)."""
    idx = index.view(-1, 1).expand_as(src)
    gmax = torch.full((n, src.size(1)), float('-inf')).scatter_reduce(0, idx, src, reduce='amax', include_self=True)
    e = (src - gmax.index_select(0, index)).exp()
    gsum = torch.zeros((n, src.size(1))).scatter_add_(0, idx, e)
    return e / (gsum.index_select(0, index) + 1e-16)


def _propagate_gat(x, edge_index, weight, W, b, A, a0):
    x_j = x.index_select(0, edge_index[0])
    x_i = x.index_select(0, edge_index[1])
    a = F.linear(torch.cat([x_j, x_i], dim=1), A, a0)                  # :158-167
    # NB: PyG infers N = to_index.max()+1 here; identical for every group that has an edge
    a = _segment_softmax(a, edge_index[1], int(edge_index[1].max()) + 1 if edge_index.size(1) else 0)   # :171
    t = F.linear(x_j, W, b)
    msg = weight.view(-1, 1) * a * t if weight is not None else a * t  # :173-176
    return torch.zeros((x.size(0), t.size(1))).index_add_(0, edge_index[1], msg)


def lightgat_conv(x, u2i_index, i2u_index, u2i_attr, i2u_attr, sd, prefix, hetero=True):
    if hetero:
        o1 = _propagate_gat(x, u2i_index, u2i_attr, sd[f'{prefix}user2item_W.0.weight'], sd[f'{prefix}user2item_W.0.bias'],
                            sd[f'{prefix}user2item_AttNet.0.weight'], sd[f'{prefix}user2item_AttNet.0.bias'])
        o2 = _propagate_gat(x, i2u_index, i2u_attr, sd[f'{prefix}item2user_W.0.weight'], sd[f'{prefix}item2user_W.0.bias'],
                            sd[f'{prefix}item2user_AttNet.0.weight'], sd[f'{prefix}item2user_AttNet.0.bias'])
        return o1 + o2
    total = torch.cat([u2i_index, i2u_index], dim=1)
    attr = torch.cat([u2i_attr, i2u_attr]) if (u2i_attr is not None and i2u_attr is not None) else None
    return _propagate_gat(x, total, attr, sd[f'{prefix}W.0.weight'], sd[f'{prefix}W.0.bias'],
                          sd[f'{prefix}AttNet.0.weight'], sd[f'{prefix}AttNet.0.bias'])


# --------------------------------------------------------------------------------------------------
# a-5 / a-6  GraphNCF.forward, _mask_edge_index — models/gnn_ncf.py:298-378
# --------------------------------------------------------------------------------------------------
def mask_target_edges(edge_index, edge_attr, positions):
    """gnn_ncf.py:369-378 with `_pos` already looked up (`pos_df.loc[...]['pos']`)."""
    mask = torch.ones(edge_index.shape[1], dtype=torch.bool)
    mask[positions] = False
    return edge_index[:, mask], (edge_attr[mask] if edge_attr is not None else None)


def graph_encode(sd, graph: dict, num_gnn_layers: int, *, hetero=True, concat=False, convType='LightGCN',
                 masked_positions=None, x0=None):
    """gnn_ncf.py:300-351 — node embed, (training) target-edge masking, L shared-weight convs, mean/concat.

    `graph` holds item_features, user_features, user2item_edge_index, item2user_edge_index and optionally
    the two edge_attr tensors.  `masked_positions` (int64) are the `pos_df` positions of the batch's
    target edges; the reference masks the SAME positions in both lists (:317,:320 both pass `_inp1`).
    `x0` overrides the node embedding (pre-embedded input, SURVEY.md §8d cfg 3/5)."""
    if x0 is None:
        item_emb = F.linear(graph['item_features'], sd['item_embeddings.0.weight'], sd['item_embeddings.0.bias'])  # :300
        user_emb = F.linear(graph['user_features'], sd['user_embeddings.0.weight'], sd['user_embeddings.0.bias'])  # :301
        x = torch.vstack([item_emb, user_emb])                                                                    # :304
    else:
        x = x0
    u2i, i2u = graph['user2item_edge_index'], graph['item2user_edge_index']
    u2i_attr, i2u_attr = graph.get('user2item_edge_attr'), graph.get('item2user_edge_attr')
    if masked_positions is not None:                                                   # :314-320
        u2i, u2i_attr = mask_target_edges(u2i, u2i_attr, masked_positions)
        i2u, i2u_attr = mask_target_edges(i2u, i2u_attr, masked_positions)
    conv = lightgcn_conv if convType == 'LightGCN' else lightgat_conv
    hs = [x]
    for _ in range(num_gnn_layers):                                                    # :337 same weights each layer
        x = conv(x, u2i, i2u, u2i_attr, i2u_attr, sd, 'gnn_convs.0.', hetero)
        hs.append(x)
    if concat:
        return torch.cat(hs, dim=1)                                                    # :349
    return torch.mean(torch.stack(hs, dim=0), dim=0)                                   # :351


def graph_ncf_forward(sd, graph: dict, userIds, itemIds, num_gnn_layers: int, *, hetero=True, concat=False,
                      use_dot_product=False, convType='LightGCN', masked_positions=None, x0=None):
    emb = graph_encode(sd, graph, num_gnn_layers, hetero=hetero, concat=concat, convType=convType,
                       masked_positions=masked_positions, x0=x0)
    item_emb = emb[itemIds]                                                            # :354
    user_emb = emb[userIds]                                                            # :357
    if use_dot_product:
        return torch.bmm(user_emb.unsqueeze(1), item_emb.unsqueeze(2)).view(-1, 1)     # :365
    return mlp_forward(torch.cat((item_emb, user_emb), dim=1), sd)                     # :361-362 item first


# --------------------------------------------------------------------------------------------------
# a-7  create_graph + node-id assignment — content_providers/graph_providers.py:10-66, 76-80
# --------------------------------------------------------------------------------------------------
def node_ids(all_user_ids: np.ndarray, all_item_ids: np.ndarray):
    """Items sorted -> 0..nI-1, users sorted -> nI..nI+nU-1 (graph_providers.py:76-80)."""
    items_sorted = np.unique(all_item_ids)
    users_sorted = np.unique(all_user_ids)
    return users_sorted, items_sorted


def create_graph(user_raw: np.ndarray, item_raw: np.ndarray, rating: np.ndarray,
                 users_sorted: np.ndarray, items_sorted: np.ndarray, binary: bool = False) -> dict:
    """Vectorised restatement of the `iterrows()` loop of graph_providers.py:16-66.

    Per interaction row, in file order: edge [u, i] with attr r - (mean_u + 2.5)/2 and edge [i, u] with attr
    r - (mean_i + 2.5)/2, means taken over THIS interaction list in float64 (pandas groupby.mean, :16-17)
    and the attrs rounded to fp32 at the end (:63-64).  `binary=True` keeps only r >= centre and drops
    attrs.  `pos_*` are the `pos_df` rows: (src, dst) -> running index within its own list (:37,:46).
    """
    nI = items_sorted.shape[0]
    u_idx = np.searchsorted(users_sorted, user_raw)
    i_idx = np.searchsorted(items_sorted, item_raw)
    u_node = (u_idx + nI).astype(np.int64)
    i_node = i_idx.astype(np.int64)
    r = rating.astype(np.float64)
    cnt_u = np.bincount(u_idx, minlength=users_sorted.shape[0])
    cnt_i = np.bincount(i_idx, minlength=nI)
    mean_u = np.bincount(u_idx, weights=r, minlength=users_sorted.shape[0]) / np.maximum(cnt_u, 1)
    mean_i = np.bincount(i_idx, weights=r, minlength=nI) / np.maximum(cnt_i, 1)
    user_avg = (mean_u[u_idx] + 2.5) / 2                                               # :32
    item_avg = (mean_i[i_idx] + 2.5) / 2                                               # :41
    keep_u = np.ones(r.shape[0], dtype=bool) if not binary else (r >= user_avg)        # :33
    keep_i = np.ones(r.shape[0], dtype=bool) if not binary else (r >= item_avg)        # :42
    g = {
        'user2item_edge_index': np.stack([u_node[keep_u], i_node[keep_u]]),            # :34,:61
        'item2user_edge_index': np.stack([i_node[keep_i], u_node[keep_i]]),            # :43,:62
        'user2item_edge_attr': None if binary else (r - user_avg).astype(np.float32),  # :36,:63
        'item2user_edge_attr': None if binary else (r - item_avg).astype(np.float32),  # :45,:64
    }
    g['pos_user2item'] = np.arange(int(keep_u.sum()), dtype=np.int64)                  # :37 running index i
    g['pos_item2user'] = np.arange(int(keep_i.sum()), dtype=np.int64)                  # :46 running index j
    return g


def csr_by_destination(edge_index: np.ndarray, num_nodes: int):
    """Canonical neighbour index used by the CUDA SpMM: edges STABLY sorted by destination node (ties keep
    interactions-file order, the order in which the reference's CPU index_add_ accumulates).
    Returns row_ptr int64[num_nodes+1], src int64[E], perm int64[E] (perm[k] = original edge position)."""
    dst = edge_index[1]
    perm = np.argsort(dst, kind='stable').astype(np.int64)
    row_ptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(dst, minlength=num_nodes), out=row_ptr[1:])
    return row_ptr, edge_index[0][perm].astype(np.int64), perm


# --------------------------------------------------------------------------------------------------
# a-8  DynamicProfilesProvider.collate_interacted_items — content_providers/dynamic_profiles_provider.py:30-73
# --------------------------------------------------------------------------------------------------
def collate_interacted_items(batch_users: np.ndarray, batch_items: np.ndarray, row_ptr: np.ndarray,
                             rated_idx: np.ndarray, rated_rating: np.ndarray, mean_rating: np.ndarray,
                             profiles: np.ndarray, ignore_ratings: bool = False):
    """`row_ptr/rated_idx/rated_rating/mean_rating` are the `user_ratings` frame as CSR with every row sorted
    by item id (the load-bearing sort of :64-66).  Returns (rated_items_ids, candidate_items (B,F),
    rated_items (I,F), user_matrix (B,I)) as float32 arrays, following :46-71."""
    segs = [rated_idx[row_ptr[u]:row_ptr[u + 1]] for u in batch_users]
    rated_ids = np.sort(np.unique(np.concatenate(segs)))                               # :59
    B, I = len(batch_users), rated_ids.shape[0]
    um = np.zeros((B, I), dtype=np.float64)
    for b, u in enumerate(batch_users):                                                # :62 multi-hot
        cols = np.searchsorted(rated_ids, segs[b])
        if ignore_ratings:
            um[b, cols] = 1.0
        else:                                                                          # :66 centred ratings
            um[b, cols] = rated_rating[row_ptr[u]:row_ptr[u + 1]] - (mean_rating[u] + 2.5) / 2
    return (rated_ids, profiles[batch_items].astype(np.float32), profiles[rated_ids].astype(np.float32),
            um.astype(np.float32))


# --------------------------------------------------------------------------------------------------
# f-4  RankingDataset negative sampling — neural_collaborative_filtering/datasets/base.py:57-78, BPR loss :97-98
# --------------------------------------------------------------------------------------------------
def negative_sampling_probs(negative_ratings: np.ndarray, w: float) -> np.ndarray:
    """'sum_dynamic' (:62-66): p = r^w / sum(r^w) with the builtin `sum` of the reference (left-to-right float64)."""
    boosted = np.asarray(negative_ratings, dtype=np.float64) ** w
    return boosted / sum(boosted)


def choice_given_uniform(p: np.ndarray, u: float) -> int:
    """What `np.random.choice(ids, p=p)` (:76) does with ONE uniform of its stream (numpy/random/mtrand.pyx `choice`, branch `p is not None`,
    size None): `cdf = p.cumsum(); cdf /= cdf[-1]; idx = cdf.searchsorted(uniform, side='right')`.  tests/test_oracle_sampling.py pins this
    to the real np.random.choice under a seeded legacy generator."""
    cdf = np.asarray(p, dtype=np.float64).cumsum()
    cdf /= cdf[-1]
    return int(cdf.searchsorted(u, side='right'))


def bpr_loss(out_pos: torch.Tensor, out_neg: torch.Tensor) -> torch.Tensor:
    """:97-98"""
    return torch.sum(-torch.log(torch.sigmoid(out_pos - out_neg)))
