"""Shared helpers of the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance for fp32 features and gradients
RTOL = 1e-5


def golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def params_of(g, prefix="param."):
    return {k[len(prefix):]: torch.from_numpy(v.copy()) for k, v in g.items() if k.startswith(prefix)}


def to_np(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


def assert_parity(actual, expected, what="", rtol=RTOL, scale_atol=RTOL):
    """fp32 parity: |a-e| <= rtol*|e| + scale_atol*max|e| elementwise, and normwise relative error
    <= rtol.  (A pure relative bound is meaningless for entries that cancel to ~0; the absolute
    term is tied to the tensor's own scale, not a free constant.)"""
    a = to_np(actual).astype(np.float64)
    e = to_np(expected).astype(np.float64)
    assert a.shape == e.shape, f"{what}: shape {a.shape} vs {e.shape}"
    if e.size == 0:
        return
    assert np.isfinite(a).all(), f"{what}: non-finite values"
    scale = np.abs(e).max()
    err = np.abs(a - e)
    bound = rtol * np.abs(e) + scale_atol * scale
    bad = err > bound
    nrm = np.linalg.norm(e.ravel())
    rel = np.linalg.norm((a - e).ravel()) / nrm if nrm > 0 else np.linalg.norm(a.ravel())
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{e.size} entries outside tolerance, "
                           f"max err {err.max():.3e} (scale {scale:.3e}), normwise rel {rel:.3e}")
    assert rel <= max(rtol, 1e-12), f"{what}: normwise relative error {rel:.3e} > {rtol}"


def assert_bitexact(actual, expected, what=""):
    a, e = to_np(actual), to_np(expected)
    assert a.shape == e.shape, f"{what}: shape {a.shape} vs {e.shape}"
    assert a.dtype == e.dtype or a.dtype.kind == e.dtype.kind, f"{what}: dtype {a.dtype} vs {e.dtype}"
    if a.dtype.kind == "f":
        same = a.view(np.uint32 if a.dtype == np.float32 else np.uint64) == \
            e.astype(a.dtype).view(np.uint32 if a.dtype == np.float32 else np.uint64)
    else:
        same = a == e
    assert same.all(), f"{what}: {int((~same).sum())}/{a.size} entries differ bitwise"
