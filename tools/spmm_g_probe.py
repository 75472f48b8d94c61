"""K3 row-owner kernel on 512-byte rows with G = 32 / 16 / 8 lanes per edge (B200REC_SPMM_G512, read once per process): whole configs[2]
graph and one rank's shard of an 8-way partition; CUDA events, median of 7.  python tools/spmm_g_probe.py  (spawns itself per G)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch
    import bench
    from deeprecommendation_b200 import ops
    from deeprecommendation_b200.graph import get_index
    from deeprecommendation_b200.peer import emulated_shards
    from tools.shard_probe import timed
    dev = torch.device('cuda:0')
    w = bench.build_graph(dev, 1.0, 1)
    graph, d = w['graph'], w['d']
    full = get_index(graph)
    t = torch.randn(full.num_nodes, d, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    xn = torch.empty(full.num_nodes, d, device=dev)
    out = {'G': os.environ.get('B200REC_SPMM_G512', '32')}
    out['full_ms'] = timed(lambda: ops.spmm_raw(full, t, w=full.w, dinv=full.dinv, x_next=xn))
    out['checksum'] = float(xn.double().sum())
    out['absmax'] = float(xn.abs().max())
    sh = emulated_shards(graph, 8, d_max=d, batch_max=64)[2]
    tu = t[:sh.users_rows].contiguous()
    T = sh.table(0, d, torch.float32)
    T.copy_(t[:T.shape[0]])
    xu = torch.empty(sh.users_rows, d, device=dev)
    out['A_push_ms'] = timed(lambda: ops.spmm_raw(sh.index_items, tu, w=sh.index_items.w, dinv=sh.dinv_items_all, push=sh.push_spec(0, d)))
    out['B_ms'] = timed(lambda: ops.spmm_raw(sh.index_users, T, w=sh.index_users.w, dinv=sh.dinv_users, x_next=xu))
    out['B_checksum'] = float(xu.double().sum())
    print(json.dumps(out))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'child':
        child()
    else:
        res = []
        for g in ('32', '16', '8'):
            env = dict(os.environ, B200REC_SPMM_G512=g)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), 'child'], capture_output=True, text=True, env=env, cwd=ROOT)
            line = [l for l in r.stdout.splitlines() if l.startswith('{')]
            res.append(json.loads(line[-1]) if line else {'G': g, 'error': r.stderr[-500:]})
            print(res[-1], flush=True)
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'spmm_g_probe.json'), 'w'), indent=1)
