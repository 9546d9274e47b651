"""Shared helpers for the parity tests."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def T(a, device="cpu", grad=False):
    t = torch.from_numpy(np.asarray(a)).to(device)
    if grad:
        t.requires_grad_(True)
    return t


def rel_err(a, b):
    """max |a-b| / max(|b|) -- the norm-wise relative error the 1e-3 fp32 tolerance is stated in."""
    a = a.detach().cpu().double() if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a)).double()
    b = b.detach().cpu().double() if isinstance(b, torch.Tensor) else torch.as_tensor(np.asarray(b)).double()
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def assert_close(a, b, tol, what=""):
    e = rel_err(a, b)
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e}"
    return e


def nested_sd(arrs, prefix, device="cpu", grad=False):
    """'prefix<module>/<param>' -> {module: {param: tensor}} preserving file order."""
    out = {}
    for k, v in arrs.items():
        if k.startswith(prefix):
            mod, name = k[len(prefix):].split("/", 1)
            out.setdefault(mod, {})[name] = T(v, device, grad)
    return out


def cfg_of(arrs):
    return json.loads(str(arrs["cfg"]))
