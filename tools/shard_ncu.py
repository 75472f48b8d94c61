"""Launches the SpMMs of ONE rank of an 8-way partition of the configs[2] graph once each (for `ncu --set full`):
A (partial item rows, push epilogue), B (own user rows), both in fp32 and with bf16 messages, then the full-graph SpMM."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from deeprecommendation_b200 import ops
    from deeprecommendation_b200.graph import get_index
    from deeprecommendation_b200.peer import emulated_shards
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dev = torch.device('cuda:0')
    w = bench.build_graph(dev, 1.0, 1)
    graph, d = w['graph'], w['d']
    full = get_index(graph)
    sh = emulated_shards(graph, world, d_max=d, batch_max=1024)[3]
    nu = sh.users_rows
    tu = torch.randn(nu, d, device=dev)
    T = sh.table(0, d, torch.float32)
    T.normal_()
    xu = torch.empty(nu, d, device=dev)
    tb, Tb = tu.bfloat16(), sh.table(1, d, torch.bfloat16)
    t_full = torch.randn(full.num_nodes, d, device=dev)
    xn = torch.empty(full.num_nodes, d, device=dev)
    from deeprecommendation_b200.graph import StreamPlan
    pa = StreamPlan(sh.index_items.row_ptr, sh.index_items.col, sh.index_items.w, sh.dinv_items_all, 128)
    pb = StreamPlan(sh.index_users.row_ptr, sh.index_users.col, sh.index_users.w, sh.dinv_users, 128)
    pf = StreamPlan(full.row_ptr, full.col, full.w, full.dinv, 256)
    for _ in range(2):
        ops.spmm_stream_raw(pa, tu, push=sh.push_spec(0, d))
        ops.spmm_stream_raw(pb, T, x_next=xu)
        ops.spmm_stream_raw(pb, Tb, x_next=xu)
        ops.spmm_stream_raw(pf, t_full, x_next=xn)
        ops.spmm_raw(sh.index_items, tu, w=sh.index_items.w, dinv=sh.dinv_items_all, push=sh.push_spec(0, d))
        ops.spmm_raw(sh.index_users, T, w=sh.index_users.w, dinv=sh.dinv_users, x_next=xu)
        ops.spmm_raw(sh.index_items, tb, w=sh.index_items.w, dinv=sh.dinv_items_all, push=sh.push_spec(0, d))
        ops.spmm_raw(sh.index_users, Tb, w=sh.index_users.w, dinv=sh.dinv_users, x_next=xu)
        ops.spmm_raw(full, t_full, w=full.w, dinv=full.dinv, x_next=xn)
        torch.cuda.synchronize()


if __name__ == '__main__':
    main()
