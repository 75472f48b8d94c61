"""Helpers shared by the parity tests: load tests/golden/*.npz (made by oracle/make_golden.py)."""
import ast
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    sd = {k[3:]: torch.from_numpy(v.copy()) for k, v in d.items() if k.startswith('w::')}
    kwargs = ast.literal_eval(str(d['kwargs'])) if 'kwargs' in d else None
    data = {k: v for k, v in d.items() if not k.startswith('w::') and k != 'kwargs'}
    return data, sd, kwargs


def maxnorm_rel(a, b):
    """max|a-b| / max|b| — the tolerance metric of SURVEY.md §7.3 (outputs cross zero)."""
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
