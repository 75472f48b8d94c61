"""Multi-GPU check (run under torchrun on >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mp_graph_check.py

Every rank builds the same graph, rank-partitions it, runs the partitioned GraphNCF forward (NCCL all-gather per layer) and
compares with the single-GPU forward computed locally.  Prints one JSON line per rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from deeprecommendation_b200 import synth
    from deeprecommendation_b200.graph import IdTable, create_graph
    from deeprecommendation_b200.neural_collaborative_filtering.models import GraphNCF
    from deeprecommendation_b200.parallel import partition_graph
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    n_users, n_items, n, F, d = 20000, 8000, 1_500_000, 64, 128
    users, items, ratings = synth.interactions_zipf(n_users, n_items, n, seed=21)
    rng = np.random.default_rng(4)
    fi, fu = rng.standard_normal((n_items, F)).astype(np.float32), rng.standard_normal((n_users, F)).astype(np.float32)
    kw = dict(item_dim=F, user_dim=F, num_gnn_layers=2, hetero=True, node_emb=d, mlp_dense_layers=[256, 128], dropout_rate=0.2)
    sd = synth.to_torch(synth.graph_ncf_weights(seed=5, **kw))
    g = create_graph(torch.from_numpy(users).to(dev), torch.from_numpy(items).to(dev), torch.from_numpy(ratings).to(dev),
                     torch.from_numpy(fi).to(dev), torch.from_numpy(fu).to(dev),
                     IdTable(torch.arange(n_users, device=dev)), IdTable(torch.arange(n_items, device=dev)))
    m = GraphNCF(**kw).to(dev).eval()
    m.load_state_dict(sd)
    pick = rng.permutation(n)[:1000]
    uid, iid = g.user2item_edge_index[0][pick].contiguous(), g.user2item_edge_index[1][pick].contiguous()
    err, info = 0.0, {}
    with torch.no_grad():
        ref = m(g, uid, iid, dev)
        for scheme in ('gather', 'reduce'):
            pg = partition_graph(g, scheme=scheme)
            out = m(g, uid, iid, dev)
            e = float((out - ref).abs().max() / ref.abs().max())
            info[scheme] = e
            err = max(err, e)
    errs = [None] * world
    dist.all_gather_object(errs, (err, pg.edges_own, pg.users.rows, info))
    if rank == 0:
        print(json.dumps({'world': world, 'max_rel_err_per_rank': [e[0] for e in errs], 'edges_per_rank': [e[1] for e in errs],
                          'user_rows_per_rank': [e[2] for e in errs], 'per_scheme': errs[0][3], 'ok': all(e[0] < 1e-5 for e in errs)}), flush=True)
    dist.destroy_process_group()
    if err >= 1e-5:
        sys.exit(1)


if __name__ == '__main__':
    main()
