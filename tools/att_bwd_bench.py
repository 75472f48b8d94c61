"""Times K2's training pair on one GPU at the config-2 batch shape: forward with attention weights (b200rec_attention_pool) and
the native backward (b200rec_attention_pool_backward), CUDA events, L2 flushed between repeats.

    python tools/att_bwd_bench.py [B] [I]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeprecommendation_b200 import _lib as L   # noqa: E402
from deeprecommendation_b200 import ops        # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    I = int(sys.argv[2]) if len(sys.argv) > 2 else 9461
    H = U = 128
    dev = torch.device('cuda:0')
    rng = np.random.default_rng(0)
    lens = np.clip(rng.lognormal(5.6, 1.0, size=B).astype(int), 20, min(2698, I))     # pairs are sampled by activity: mean ~550 per row
    um = np.zeros((B, I), dtype=np.float32)
    for b in range(B):
        um[b, rng.choice(I, size=lens[b], replace=False)] = rng.integers(1, 11, size=lens[b]) * 0.5 - 2.75
    um = torch.from_numpy(um).to(dev)
    nnz = int((um != 0).sum())
    g = torch.Generator(device=dev).manual_seed(1)
    Pc, Pr, Q = (torch.randn(n, H, device=dev, generator=g) * 0.5 for n in (B, I, I))
    a2, a20, bU = torch.randn(H, device=dev, generator=g), torch.zeros(1, device=dev), torch.zeros(U, device=dev)
    gout = torch.randn(B, U, device=dev, generator=g)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out, att = ops.attention_pool_raw(Pc, Pr, Q, mode=L.ATT_NET, a2=a2, a20=a20, bU=bU, user_matrix=um, return_attention_weights=True)

    def timed(fn, reps=10):
        ts = []
        for r in range(reps + 2):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            if r >= 2:
                ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    fwd = timed(lambda: ops.attention_pool_raw(Pc, Pr, Q, mode=L.ATT_NET, a2=a2, a20=a20, bU=bU, user_matrix=um, return_attention_weights=True))
    bwd = timed(lambda: ops.attention_pool_backward_raw(Pc, Pr, Q, a2, bU, um, att, out, gout, L.ATT_NET))
    # bytes the backward must move: alpha + um rows once, per non-zero one Pr and one Q row in and one dPr and one dQ row reduced
    alg = 2 * B * I * 4 + nnz * 4 * (H + U) * 4 // 2 + nnz * 0
    alg = 2 * B * I * 4 + nnz * (2 * (H + U) * 4)
    print(json.dumps({'B': B, 'I': I, 'nnz': nnz, 'forward_with_weights_ms': round(fwd, 4), 'backward_ms': round(bwd, 4),
                      'backward_algorithmic_bytes': alg, 'backward_GBs': round(alg / (bwd * 1e-3) / 1e9, 1),
                      'note': 'tables 2 x 4.8 MB are L2-resident at this shape; includes the host-side launch path of ops.py'}))


if __name__ == '__main__':
    main()
