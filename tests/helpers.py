"""Shared test helpers (CPU side).  The oracle is the checker; nothing here is product code."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def layer_case(blob, name):
    pre = name + '/'
    return {k[len(pre):]: blob[k] for k in blob.files if k.startswith(pre)}


def rel_err(a, b):
    """max-abs-diff / max-abs-ref (SURVEY.md section 8(d) parity definition)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(np.abs(b).max() if b.size else 0.0, 1e-30)
    return float(np.abs(a - b).max() / denom) if b.size else 0.0


def t(x, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(x))
    return x.to(dtype) if dtype is not None else x
