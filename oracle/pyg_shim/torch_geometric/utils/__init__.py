import torch


def degree(index, num_nodes=None, dtype=None):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    out = torch.zeros((n,), dtype=dtype, device=index.device)
    return out.scatter_add_(0, index, torch.ones((index.size(0),), dtype=out.dtype, device=index.device))


def softmax(src, index=None, ptr=None, num_nodes=None, dim=0):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    shape = (n,) + tuple(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    gmax = torch.full(shape, float('-inf'), dtype=src.dtype, device=src.device)
    gmax = gmax.scatter_reduce(0, idx, src, reduce='amax', include_self=True)
    out = (src - gmax.index_select(0, index)).exp()
    gsum = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(0, idx, out)
    return out / (gsum.index_select(0, index) + 1e-16)


def subgraph(subset, edge_index, edge_attr=None, relabel_nodes=False, num_nodes=None, return_edge_mask=False):
    if relabel_nodes:
        raise NotImplementedError
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    node_mask = torch.zeros(n, dtype=torch.bool, device=edge_index.device)
    node_mask[subset] = True
    edge_mask = node_mask[edge_index[0]] & node_mask[edge_index[1]]
    ei = edge_index[:, edge_mask]
    ea = edge_attr[edge_mask] if edge_attr is not None else None
    return (ei, ea, edge_mask) if return_edge_mask else (ei, ea)
