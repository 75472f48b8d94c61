"""Where the HOST time of an eager resident forward goes (cProfile over the bench's own batches): python tools/host_profile.py [n_iters]"""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device('cuda:0')
torch.cuda.set_device(dev)
w = bench.build_attention(dev, 0, n_batches=4)
model, rf = w['model'], w['resident_form']
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100


def loop(k):
    for i in range(k):
        c, r, um = rf[i % len(rf)]
        with torch.no_grad():
            out = model.forward_resident(c, r, um)
    torch.cuda.synchronize()
    return out


loop(20)
t0 = time.perf_counter()
loop(n)
wall = (time.perf_counter() - t0) / n * 1e3
pr = cProfile.Profile()
pr.enable()
loop(n)
pr.disable()
s = io.StringIO()
st = pstats.Stats(pr, stream=s)
st.sort_stats('tottime').print_stats(45)
print(f'wall per forward_resident (no profiler): {wall:.4f} ms')
print(s.getvalue())
