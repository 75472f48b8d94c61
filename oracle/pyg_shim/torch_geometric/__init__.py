"""Minimal CPU restatement of the five torch_geometric 2.0.4 symbols the reference uses.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference
(`/root/reference/src/neural_collaborative_filtering/models/gnn_ncf.py:4-5`,
`content_providers.py:1`, `src/content_providers/graph_providers.py:3`) imports
`torch_geometric`, which is pinned at pyg=2.0.4 in the reference's `env.yml:122` and is NOT
installable here (no network).  Only these symbols are reached:

  torch_geometric.data.Data                  attribute bag + .to(device)
  torch_geometric.nn.MessagePassing          __init__(aggr='add'), propagate(), message(), update()
  torch_geometric.utils.degree               scatter_add of ones
  torch_geometric.utils.softmax              per-group softmax, denominator + 1e-16
  torch_geometric.utils.subgraph             induced-subgraph edge filter, no relabelling

Semantics restated from the published PyG 2.0.4 behaviour (flow='source_to_target',
node_dim=0).  Parity for anything that goes through this shim is "unpinned" in the sense of
SURVEY.md §8(c): the reference ships no test vectors at this boundary.
"""
__version__ = "2.0.4-shim"
