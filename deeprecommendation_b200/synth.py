"""Seeded synthetic inputs shaped like BASELINE.json's configs: data for tests, smoke() and bench.py.

No dataset ships with the reference (`data/` is git-ignored there) and there is no network, so every
parity test and benchmark runs on these generators.  All of them are pure numpy and deterministic
in `seed` (42 = `src/globals.py:26` of the reference).
"""
from __future__ import annotations

import numpy as np

F_BINARY = 966      # 21 genres + 945 personnel multi-hot columns (create_dataset.ipynb cell 15/16)
F_DENSE = 1128      # genome-tag relevances in [0, 1]
F_PROFILE = F_BINARY + F_DENSE   # 2094


def item_profiles(n_items: int, seed: int = 42, f_binary: int = F_BINARY, f_dense: int = F_DENSE,
                  p: float = 0.01) -> np.ndarray:
    """(n_items, f_binary + f_dense) float32: sparse binary block then dense [0,1) block."""
    rng = np.random.default_rng(seed)
    out = np.empty((n_items, f_binary + f_dense), dtype=np.float32)
    out[:, :f_binary] = (rng.random((n_items, f_binary), dtype=np.float32) < p)
    out[:, f_binary:] = rng.random((n_items, f_dense), dtype=np.float32)
    return out


def _counts_lognormal(n_users: int, total: int, lo: int, hi: int, rng) -> np.ndarray:
    raw = rng.lognormal(mean=0.0, sigma=1.15, size=n_users)
    c = raw / raw.sum() * total
    for _ in range(64):                      # clip + rescale until the total matches
        c = np.clip(c, lo, hi)
        free = (c > lo) & (c < hi)
        err = total - c.sum()
        if abs(err) < 0.5 or not free.any():
            break
        c[free] += err * c[free] / c[free].sum()
    c = np.clip(np.floor(c), lo, hi).astype(np.int64)
    i = 0
    order = np.argsort(-c)
    while c.sum() != total:                  # settle the rounding remainder on the largest users
        j = order[i % n_users]
        step = 1 if c.sum() < total else -1
        if lo <= c[j] + step <= hi:
            c[j] += step
        i += 1
    return c


def interactions_small(n_users: int = 610, n_items: int = 9724, n_ratings: int = 100_836,
                       lo: int = 20, hi: int = 2698, seed: int = 42):
    """MovieLens-latest-small-shaped (u, i, r) triplets: unique pairs, file order shuffled.

    Returns int64 userId (1-based, like MovieLens), int64 movieId (sparse ids: 7*k+1, to exercise the
    sorted-unique node-id assignment of graph_providers.py:76-80) and float64 ratings in {0.5..5.0}.
    """
    rng = np.random.default_rng(seed)
    hi = min(hi, n_items)
    lo = min(lo, hi)
    counts = _counts_lognormal(n_users, n_ratings, lo, hi, rng)
    pop = 1.0 / np.arange(1, n_items + 1) ** 0.8          # item popularity ~ Zipf(0.8)
    pop /= pop.sum()
    users, items = [], []
    for u in range(n_users):
        it = rng.choice(n_items, size=int(counts[u]), replace=False, p=pop)
        users.append(np.full(it.shape, u, dtype=np.int64))
        items.append(it.astype(np.int64))
    users = np.concatenate(users)
    items = np.concatenate(items)
    ratings = rng.integers(1, 11, size=users.shape[0]).astype(np.float64) * 0.5
    perm = rng.permutation(users.shape[0])
    return users[perm] + 1, items[perm] * 7 + 1, ratings[perm]


def interactions_zipf(n_users: int, n_items: int, n_ratings: int, seed: int = 42,
                      a_user: float = 0.55, a_item: float = 0.95, active_items: float = 0.946):
    """MovieLens-25M-shaped bipartite edge list with heavy-tailed degrees on both sides.

    Users and items are drawn independently per edge from truncated power laws and de-duplicated, then
    topped up until exactly `n_ratings` unique pairs exist.  `active_items` leaves a fraction of items
    with no rating at all (59,047 of 62,423 in MovieLens-25M), which exercises zero-degree rows.
    Returns dense 0-based int64 ids (users, items) and float64 ratings in {0.5..5.0}, file order random.
    """
    rng = np.random.default_rng(seed)
    n_act = max(1, int(round(n_items * active_items)))
    pu = 1.0 / np.arange(1, n_users + 1) ** a_user
    pi = 1.0 / np.arange(1, n_act + 1) ** a_item
    cu, ci = np.cumsum(pu / pu.sum()), np.cumsum(pi / pi.sum())
    user_perm = rng.permutation(n_users)          # popularity rank -> id
    item_perm = rng.permutation(n_items)[:n_act]
    keys = np.empty(0, dtype=np.int64)
    while keys.shape[0] < n_ratings:
        need = n_ratings - keys.shape[0]
        m = int(need * 1.25) + 1024
        u = np.minimum(np.searchsorted(cu, rng.random(m)), n_users - 1)
        i = np.minimum(np.searchsorted(ci, rng.random(m)), n_act - 1)
        k = user_perm[u].astype(np.int64) * n_items + item_perm[i]
        keys = np.unique(np.concatenate([keys, k]))
        if keys.shape[0] > n_ratings:
            keys = rng.permutation(keys)[:n_ratings]
    keys = rng.permutation(keys)
    users, items = keys // n_items, keys % n_items
    ratings = rng.integers(1, 11, size=n_ratings).astype(np.float64) * 0.5
    return users, items, ratings


def fixed_user_profiles(users: np.ndarray, items: np.ndarray, ratings: np.ndarray,
                        profiles: np.ndarray, n_users: int) -> np.ndarray:
    """Reference formula for fixed user profiles (create_dataset.ipynb cell 45, `create_user_embedding`):
    mean_i[(r_ui - (mean_u + 2.5)/2) * item_profile_i].  `users`/`items` are dense 0-based indices."""
    cnt = np.bincount(users, minlength=n_users).astype(np.float64)
    mean_u = np.bincount(users, weights=ratings, minlength=n_users) / np.maximum(cnt, 1)
    w = ratings - (mean_u[users] + 2.5) / 2
    out = np.zeros((n_users, profiles.shape[1]), dtype=np.float64)
    np.add.at(out, users, w[:, None] * profiles[items].astype(np.float64))
    out /= np.maximum(cnt, 1)[:, None]
    return out.astype(np.float32)


def dense_ids(raw: np.ndarray):
    """sorted-unique raw ids -> dense index (graph_providers.py:76-80 semantics)."""
    uniq, inv = np.unique(raw, return_inverse=True)
    return uniq, inv.astype(np.int64)


def user_rating_lists(users: np.ndarray, items: np.ndarray, ratings: np.ndarray, n_users: int):
    """Per-user (movieId sorted) rating lists = the `user_ratings` frame of create_dataset.ipynb cell 45.

    Returns CSR (row_ptr int64[n_users+1], item_idx int64[nnz] sorted within each row, rating float64[nnz],
    mean_rating float64[n_users])."""
    order = np.lexsort((items, users))
    u, i, r = users[order], items[order], ratings[order]
    cnt = np.bincount(u, minlength=n_users)
    row_ptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(cnt, out=row_ptr[1:])
    mean = np.bincount(u, weights=r, minlength=n_users) / np.maximum(cnt, 1)
    return row_ptr, i, r, mean


# --------------------------------------------------------------------------------------------------
# numpy-seeded weights with the reference's state_dict key names.  numpy's Generator
# stream is stable across machines, unlike torch's default initialisers, so golden outputs at full
# shapes can be stored without storing the weights.
# --------------------------------------------------------------------------------------------------
def _linear(rng, out_f: int, in_f: int, scale: float = 1.0):
    bound = scale / np.sqrt(in_f)
    w = rng.uniform(-bound, bound, size=(out_f, in_f)).astype(np.float32)
    b = rng.uniform(-bound, bound, size=(out_f,)).astype(np.float32)
    return w, b


def _mlp(rng, sd: dict, in_f: int, layers, dropout_rate, prefix='MLP.'):
    sizes = list(layers) + [1]
    step = 2 if dropout_rate is None else 3          # util.py:13-17: (ReLU, [Dropout], Linear)
    prev = in_f
    for n, h in enumerate(sizes):
        sd[f'{prefix}{n * step}.weight'], sd[f'{prefix}{n * step}.bias'] = _linear(rng, h, prev)
        prev = h


def basic_ncf_weights(item_dim, user_dim, item_emb=128, user_emb=128, mlp_dense_layers=(256,),
                      dropout_rate=0.2, seed=0) -> dict:
    rng = np.random.default_rng(seed)
    sd = {}
    sd['item_embeddings.0.weight'], sd['item_embeddings.0.bias'] = _linear(rng, item_emb, item_dim)
    sd['user_embeddings.0.weight'], sd['user_embeddings.0.bias'] = _linear(rng, user_emb, user_dim)
    _mlp(rng, sd, item_emb + user_emb, mlp_dense_layers, dropout_rate)
    return sd


def attention_ncf_weights(item_dim, item_emb=128, user_emb=128, att_dense=128, mlp_dense_layers=(256, 128),
                          dropout_rate=0.2, seed=0) -> dict:
    rng = np.random.default_rng(seed)
    sd = {}
    # scale > 1 on the attention layers so the softmax is far from uniform (a sharper parity test)
    sd['ItemEmbeddings.0.weight'], sd['ItemEmbeddings.0.bias'] = _linear(rng, item_emb, item_dim)
    sd['UserEmbeddings.0.weight'], sd['UserEmbeddings.0.bias'] = _linear(rng, user_emb, item_dim)
    if att_dense is not None:
        sd['AttentionNet.0.weight'], sd['AttentionNet.0.bias'] = _linear(rng, att_dense, 2 * item_emb, 4.0)
        sd['AttentionNet.3.weight'], sd['AttentionNet.3.bias'] = _linear(rng, 1, att_dense, 4.0)
    else:
        sd['AttentionNet.0.weight'], sd['AttentionNet.0.bias'] = _linear(rng, 1, 2 * item_emb, 4.0)
    _mlp(rng, sd, item_emb + user_emb, mlp_dense_layers, dropout_rate)
    return sd


def graph_ncf_weights(item_dim, user_dim, num_gnn_layers, node_emb=64, mlp_dense_layers=(256, 128),
                      dropout_rate=0.2, hetero=True, concat=False, use_dot_product=False,
                      convType='LightGCN', seed=0) -> dict:
    rng = np.random.default_rng(seed)
    sd = {}
    sd['item_embeddings.0.weight'], sd['item_embeddings.0.bias'] = _linear(rng, node_emb, item_dim)
    sd['user_embeddings.0.weight'], sd['user_embeddings.0.bias'] = _linear(rng, node_emb, user_dim)
    names = ['user2item_W', 'item2user_W'] if hetero else ['W']
    conv = {}
    for nm in names:
        conv[f'{nm}.0.weight'], conv[f'{nm}.0.bias'] = _linear(rng, node_emb, node_emb, 3.0)
    if convType == 'LightGAT':
        for nm in (['user2item_AttNet', 'item2user_AttNet'] if hetero else ['AttNet']):
            conv[f'{nm}.0.weight'], conv[f'{nm}.0.bias'] = _linear(rng, 1, 2 * node_emb, 4.0)
    for k in range(num_gnn_layers):                  # gnn_ncf.py:227: ONE conv object aliased L times
        for name, v in conv.items():
            sd[f'gnn_convs.{k}.{name}'] = v
    if not use_dot_product:
        in_f = node_emb * (num_gnn_layers + 1) * 2 if concat else node_emb * 2
        _mlp(rng, sd, in_f, mlp_dense_layers, dropout_rate)
    return sd


def to_torch(sd: dict):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}
