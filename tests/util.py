"""Shared helpers for the test-suite (golden loading, comparison metrics, package import)."""
import importlib
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PKG_NAME = "semi-supervised-asr_b200"


def pkg(sub=None):
    """The product package (its directory name has a hyphen, so no plain `import`)."""
    return importlib.import_module(PKG_NAME + ("." + sub if sub else ""))


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
    a = torch.as_tensor(a).detach().cpu().double().flatten()
    b = torch.as_tensor(b).detach().cpu().double().flatten()
    na, nb = a.norm(), b.norm()
    if na == 0 and nb == 0:
        return 1.0
    return float(a @ b / (na * nb + 1e-300))


def rel_err(a, b):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def e2e_from_golden(G, dropout_rate=0.0):
    """Build the product E2E with the golden case's hyper-parameters and load its p0 state_dict."""
    M = pkg("model")
    g, p0 = G["raw"], G["p0"]
    sub = g["subsample"].tolist()
    H = p0["encoder.enc2.layers.0.weight_hh_l0"].shape[1]
    D = p0["encoder.enc2.layers.0.weight_ih_l0"].shape[1]
    V, E = p0["decoder.embedding.weight"].shape
    A = p0["attention.mlp_enc.weight"].shape[0]
    C, _, _, ksz = p0["attention.loc_conv.weight"].shape
    Hd = p0["decoder.LSTMCell.weight_hh"].shape[1]
    O = p0["attention.mlp_o.weight"].shape[0]
    m = M.E2E(input_dim=D, enc_hidden_dim=H, enc_n_layers=len(sub), subsample=sub, dropout_rate=dropout_rate,
              dec_hidden_dim=Hd, att_dim=A, conv_channels=C, conv_kernel_size=(ksz - 1) // 2, att_odim=O,
              embedding_dim=E, output_dim=V, ls_weight=float(g["ls_weight"]), labeldist=g["labeldist"])
    missing = m.load_state_dict(p0, strict=True)
    return m.cuda()
