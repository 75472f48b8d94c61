"""Micro-benchmark of K5 (all-pairs scoring + top-k) on a slab of BASELINE configs[3]: nU users x 100k items, MLP [256,128].
CUDA events, inputs (A, B tables: nU + nI rows of 1 KB) far smaller than the work, top-k only (scores never materialised)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeprecommendation_b200 import _lib as L, ops

nU = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nI = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
H1, H2, k = 256, 128, 10
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn(nU, H1, device=dev, generator=g)
B = torch.randn(nI, H1, device=dev, generator=g)
W2 = (torch.rand(H2, H1, device=dev, generator=g) * 2 - 1) / 16
b2 = (torch.rand(H2, device=dev, generator=g) * 2 - 1) / 16
w3 = (torch.rand(H2, device=dev, generator=g) * 2 - 1) / 11
b3 = torch.zeros(1, device=dev)
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json'))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')) else {'bf16_tflops': 1590.0}
out = {}
for name, mode in (('bf16', L.AP_BF16), ('bf16x2', L.AP_BF16X2)):
    packed = ops.allpairs_pack(W2, b2, w3, b3, H1, mode)
    ts = []
    for r in range(5):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        val, idx, _ = ops.allpairs_topk_raw(A, B, packed, mode, k)
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    pairs = nU * nI
    flop = pairs * (2.0 * H1 * H2 + 2 * H2)
    mma = flop * (3 if mode == L.AP_BF16X2 else 1)
    out[name] = {'ms': round(ms, 3), 'pairs_per_s': pairs / (ms * 1e-3), 'alg_tflops': flop / (ms * 1e-3) / 1e12,
                 'mma_tflops_issued': mma / (ms * 1e-3) / 1e12, 'frac_of_bf16_peak_issued': mma / (ms * 1e-3) / 1e12 / peaks['bf16_tflops']}
    print(name, out[name], flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'allpairs_bench.json'), 'w'), indent=1)
