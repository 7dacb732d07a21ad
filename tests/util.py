"""Shared helpers for the test-suite (golden loading, comparison metrics)."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    out = {"raw": g}
    for pre in ("p0", "g", "p1", "j0", "jg", "j1"):
        out[pre] = {k[len(pre) + 1:]: torch.from_numpy(v) for k, v in g.items() if k.startswith(pre + "/")}
    n = int(g["n_ys"])
    out["ys"] = [g[f"ys_{i}"] for i in range(n)]
    if "jys_0" in g:
        out["jys"] = [g[f"jys_{i}"] for i in range(n)]
    return out


def cosine(a, b):
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    na, nb = a.norm(), b.norm()
    if na == 0 and nb == 0:
        return 1.0
    return float(a @ b / (na * nb + 1e-300))


def rel_err(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))
