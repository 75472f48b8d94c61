import inspect

import torch
from torch import nn


class MessagePassing(nn.Module):
    """`propagate` for aggr='add', flow='source_to_target', node_dim=0.

    For every parameter of `message` named `foo_j` (resp. `foo_i`) the keyword `foo` is
    index-selected along dim 0 with edge_index[0] (resp. edge_index[1]); if `foo` is a tuple,
    element 0 feeds `_j` and element 1 feeds `_i`.  All other keywords are forwarded untouched.
    The messages are scatter-added over edge_index[1] into a zero tensor with `size(0)` = number of
    nodes of the `x` argument.  `update` is the identity.
    """

    def __init__(self, aggr='add', flow='source_to_target', node_dim=0, **kwargs):
        super().__init__()
        if aggr != 'add' or flow != 'source_to_target' or node_dim != 0:
            raise NotImplementedError('shim covers only what the reference uses')
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]
        params = list(inspect.signature(self.message).parameters)
        call = {}
        num_dst = None
        for p in params:
            if p.endswith('_j') or p.endswith('_i'):
                base, which = p[:-2], p[-2:]
                data = kwargs[base]
                src_data, dst_data = (data[0], data[1]) if isinstance(data, (tuple, list)) else (data, data)
                if num_dst is None:
                    num_dst = dst_data.size(0)          # output rows = number of destination nodes
                call[p] = src_data.index_select(0, src) if which == '_j' else dst_data.index_select(0, dst)
            else:
                call[p] = kwargs[p]
        num_nodes = num_dst if size is None else size[1]
        msgs = self.message(**call)
        out = torch.zeros((num_nodes,) + tuple(msgs.shape[1:]), dtype=msgs.dtype, device=msgs.device)
        out.index_add_(0, dst, msgs)
        return self.update(out)

    def message(self, x_j):
        return x_j

    def update(self, inputs):
        return inputs
