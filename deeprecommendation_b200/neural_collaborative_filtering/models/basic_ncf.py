"""BasicNCF on the B200 path (reference: neural_collaborative_filtering/models/basic_ncf.py:9-48)."""
from __future__ import annotations

from torch import nn

from ... import ops
from ..util import build_MLP_layers, run_mlp
from .base import NCF, _named_like


class BasicNCF(NCF):
    """out = MLP([user_embeddings(X_user), item_embeddings(X_item)])  — user first (basic_ncf.py:40).

    Two CUDA projection GEMMs (K1a) feed the fused MLP tower (K1b); the concatenation is never materialised."""

    def __init__(self, item_dim, user_dim, dropout_rate=0.2, item_emb=256, user_emb=256, mlp_dense_layers=None):
        super().__init__()
        mlp_dense_layers = [256, 128] if mlp_dense_layers is None else mlp_dense_layers
        self.kwargs = dict(item_dim=item_dim, user_dim=user_dim, item_emb=item_emb, user_emb=user_emb,
                           mlp_dense_layers=mlp_dense_layers, dropout_rate=dropout_rate)
        self.item_embeddings = nn.Sequential(nn.Linear(item_dim, item_emb))
        self.user_embeddings = nn.Sequential(nn.Linear(user_dim, user_emb))
        self.MLP = build_MLP_layers(item_emb + user_emb, mlp_dense_layers, dropout_rate=dropout_rate)

    def forward(self, X_user, X_item):
        ue, ie = self.user_embeddings[0], self.item_embeddings[0]
        user_emb = ops.linear(X_user, ue.weight, ue.bias)
        item_emb = ops.linear(X_item, ie.weight, ie.bias)
        return run_mlp(self.MLP, user_emb, item_emb, training=self.training)

    def is_dataset_compatible(self, dataset_class):
        return _named_like(dataset_class, 'FixedPointwiseDataset', 'FixedRankingDataset')
