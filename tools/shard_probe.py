"""Per-rank kernel times of the partitioned GraphNCF propagation, measured on ONE GPU: the configs[2] graph is cut for `world`
ranks (peer.emulated_shards) and the kernels of single ranks are timed with CUDA events.  Tells what an N-GPU step can reach before
any NVLink effect (python tools/shard_probe.py [world] [scale])."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps=7):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    import bench
    from deeprecommendation_b200 import ops
    from deeprecommendation_b200.graph import get_index
    from deeprecommendation_b200.peer import emulated_shards, forward_emulated
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    dev = torch.device('cuda:0')
    w = bench.build_graph(dev, scale, 1)
    model, graph = w['model'], w['graph']
    full = get_index(graph)
    d = w['d']
    out = {'world': world, 'nU': w['nU'], 'nI': w['nI'], 'E': w['E'], 'd': d}
    t_full = torch.randn(full.num_nodes, d, device=dev)
    xn = torch.empty(full.num_nodes, d, device=dev)
    out['full_spmm_ms'] = timed(lambda: ops.spmm_raw(full, t_full, w=full.w, dinv=full.dinv, x_next=xn))
    from deeprecommendation_b200.graph import StreamPlan
    tb_full = t_full.bfloat16()
    for seg in (128, 256, 512):
        plan = StreamPlan(full.row_ptr, full.col, full.w, full.dinv, seg)
        out[f'full_stream_seg{seg}_ms'] = timed(lambda: ops.spmm_stream_raw(plan, t_full, x_next=xn))
        out[f'full_stream_seg{seg}_bf16_ms'] = timed(lambda: ops.spmm_stream_raw(plan, tb_full, x_next=xn))
        out[f'full_stream_seg{seg}_multi'] = [plan.n_multi, plan.n_slots]
        del plan
    out['full_bf16_ms'] = timed(lambda: ops.spmm_raw(full, tb_full, w=full.w, dinv=full.dinv, x_next=xn))
    shards = emulated_shards(graph, world, d_max=d, batch_max=1024)
    lin_u, lin_i, _ = model.gnn_convs[0].typed()
    per_rank = []
    for sh in shards[:: max(1, world // 4)]:
        nu, ni = sh.users_rows, sh.it_rows
        tu = torch.randn(max(nu, 1), d, device=dev)
        T = sh.table(0, d, torch.float32)
        T.normal_()
        xu, accu = torch.empty(nu, d, device=dev), torch.zeros(nu, d, device=dev)
        xi, acci = torch.empty(ni, d, device=dev), torch.zeros(ni, d, device=dev)
        part = torch.zeros(sh.nI, d, device=dev)
        r = {'rank': sh.rank, 'users': nu, 'items_own': ni, 'edges_A': int(sh.index_items.col.numel()), 'edges_B': int(sh.index_users.col.numel()),
             'chunks_A': sh.index_items.n_chunks, 'multi_A': sh.index_items.n_multi, 'chunks_B': sh.index_users.n_chunks}
        r['A_push_ms'] = timed(lambda: ops.spmm_raw(sh.index_items, tu, w=sh.index_items.w, dinv=sh.dinv_items_all, push=sh.push_spec(0, d)))
        r['A_local_ms'] = timed(lambda: ops.spmm_raw(sh.index_items, tu, w=sh.index_items.w, dinv=sh.dinv_items_all, x_next=part))
        r['B_ms'] = timed(lambda: ops.spmm_raw(sh.index_users, T, w=sh.index_users.w, dinv=sh.dinv_users, x_next=xu, acc_in=accu, acc_out=accu, acc_scale=1.0))
        r['reduce_ms'] = timed(lambda: sh.reduce(0, d, x_next=xi, acc_in=acci, acc_out=acci, acc_scale=1.0))
        r['push_transform_ms'] = timed(lambda: sh.push_transform(xi, lin_i, 1, torch.float32))
        r['t_users_ms'] = timed(lambda: ops.linear_raw(xu, lin_u.weight, lin_u.bias, row_scale=sh.dinv_users, out=tu[:nu]))
        r['signal_wait_ms'] = timed(lambda: ([s2.signal(5) for s2 in shards], sh.wait(5)))      # 8 signal kernels + 1 wait kernel
        tb = torch.randn(max(nu, 1), d, device=dev).bfloat16()
        for seg in (64, 128, 256):
            pa = StreamPlan(sh.index_items.row_ptr, sh.index_items.col, sh.index_items.w, sh.dinv_items_all, seg)
            pb = StreamPlan(sh.index_users.row_ptr, sh.index_users.col, sh.index_users.w, sh.dinv_users, seg)
            r[f'A_stream{seg}_push_ms'] = timed(lambda: ops.spmm_stream_raw(pa, tu, push=sh.push_spec(0, d)))
            r[f'B_stream{seg}_ms'] = timed(lambda: ops.spmm_stream_raw(pb, T, x_next=xu, acc_in=accu, acc_out=accu, acc_scale=1.0))
            r[f'A_stream{seg}_push_bf16_ms'] = timed(lambda: ops.spmm_stream_raw(pa, tb, push=sh.push_spec(0, d)))
            r[f'B_stream{seg}_bf16_ms'] = timed(lambda: ops.spmm_stream_raw(pb, sh.table(1, d, torch.bfloat16), x_next=xu, acc_in=accu, acc_out=accu, acc_scale=1.0))
            r[f'stream{seg}_multi_A_B'] = [pa.n_multi, pb.n_multi]
            del pa, pb
        Tb = sh.table(1, d, torch.bfloat16)
        r['A_push_bf16_ms'] = timed(lambda: ops.spmm_raw(sh.index_items, tb, w=sh.index_items.w, dinv=sh.dinv_items_all, push=sh.push_spec(0, d)))
        r['B_bf16_ms'] = timed(lambda: ops.spmm_raw(sh.index_users, Tb, w=sh.index_users.w, dinv=sh.dinv_users, x_next=xu, acc_in=accu, acc_out=accu, acc_scale=1.0))
        per_rank.append(r)
    out['ranks'] = per_rank
    # whole emulated forward (all ranks on one stream): sum over ranks of everything
    ids = (graph.user2item_edge_index[0][w['pick'][0]].contiguous(), graph.user2item_edge_index[1][w['pick'][0]].contiguous())
    with torch.no_grad():
        ref = model(graph, ids[0], ids[1], dev)
        outs = forward_emulated(model, shards, *ids)
        torch.cuda.synchronize()
        for sh in shards:
            sh.check()
        out['emulated_rel_err'] = float(max((o - ref).abs().max() / ref.abs().max() for o in outs))
        out['single_gpu_step_ms'] = timed(lambda: model(graph, ids[0], ids[1], dev))
        out['emulated_all_ranks_ms'] = timed(lambda: forward_emulated(model, shards, *ids), reps=3)
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', f'shard_probe_w{world}.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
