"""K6 at the config-2 shape: CUDA-event time of the three collate launches per 512-user batch, and the list bytes they read."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeprecommendation_b200 import ops, synth                               # noqa: E402
from deeprecommendation_b200.content_providers import DeviceCollateProvider  # noqa: E402

dev = torch.device('cuda:0')
users_raw, items_raw, ratings = synth.interactions_small(610, 9724, 100_836, seed=42)
_, u = synth.dense_ids(users_raw)
item_ids, it = synth.dense_ids(items_raw)
row_ptr, idx, rr, _ = synth.user_rating_lists(u, it, ratings, 610)
prov = DeviceCollateProvider(np.arange(len(item_ids)), np.zeros((len(item_ids), 8), np.float32), np.arange(610), row_ptr, idx, rr, device=dev)
rng = np.random.default_rng(1000)
batches = [u[rng.permutation(len(u))[:512]].astype(np.int64) for _ in range(8)]
rows = [torch.from_numpy(b).to(dev) for b in batches]
entries = [int(prov._list_len[b].sum()) for b in batches]
nnz = [int(prov._nz_cnt[b].sum()) for b in batches]


def run(k):
    return ops.collate_interacted_raw(rows[k], prov.d_list_ptr, prov.d_list_item, prov.d_list_val, prov.get_num_items(),
                                      rated_capacity=prov.get_num_items(), nnz_capacity=nnz[k])


for k in range(8):
    run(k)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 50
a.record()
for r in range(reps):
    for k in range(8):
        run(k)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / (reps * 8)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for k in range(8):
        run(k)
g.replay()
torch.cuda.synchronize()
a.record()
for r in range(reps):
    g.replay()
b.record()
torch.cuda.synchronize()
ms_graph = a.elapsed_time(b) / (reps * 8)
print(json.dumps({'what': 'K6 b200rec_collate_interacted, config-2 shape, per 512-user batch', 'ms_eager_stream': round(ms, 4),
                  'ms_cuda_graph_replay': round(ms_graph, 4), 'mean_list_entries': float(np.mean(entries)), 'mean_nnz': float(np.mean(nnz)),
                  'algorithmic_bytes': float(np.mean(entries)) * 8 * 2 + float(np.mean(nnz)) * 8,
                  'gbs_at_graph_time': round((float(np.mean(entries)) * 16 + float(np.mean(nnz)) * 8) / (ms_graph * 1e-3) / 1e9, 1)}))
