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


def det_uniform(name, shape, bound):
    """Machine-independent stand-in for a large weight matrix: uniform(-bound, bound) from numpy's MT19937 seeded by the
    parameter name.  Lets a full-size fixture (tests/golden/fullsize_*.npz) hold scalars only."""
    import zlib
    rng = np.random.RandomState(zlib.crc32(name.encode()))
    return ((rng.random_sample(tuple(shape)) * 2.0 - 1.0) * bound).astype(np.float32)


def det_normal(index, shape):
    """The index-th Gaussian draw of a full-size fixture run (replaces torch.randn_like on both sides)."""
    return np.random.RandomState(70000 + index).standard_normal(tuple(shape)).astype(np.float32)
