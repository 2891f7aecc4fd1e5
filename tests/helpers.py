"""Shared helpers for the test-suite: golden fixture loading and weight dictionaries."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name, device="cpu", dtype=None):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for k in z.files:
        a = z[k]
        if a.dtype.kind in "US":
            out[k] = str(a)
        else:
            t = torch.from_numpy(a.copy())
            if dtype is not None and t.is_floating_point():
                t = t.to(dtype)
            out[k] = t.to(device)
    return out


def weights_from(fix, names=("sdf", "feature", "deformation")):
    return {n: [fix[f"w_{n}_{i}"] for i in range(3)] for n in names if f"w_{n}_0" in fix}


def max_abs(a, b):
    return (a.double() - b.double()).abs().max().item()


def rel_err(a, b):
    """max |a-b| / (max |b| + tiny): scale-aware error for gradient tensors."""
    return (a.double() - b.double()).abs().max().item() / (b.double().abs().max().item() + 1e-30)
