"""Comparison helpers shared by the CPU (oracle vs reference fixtures) and GPU (CUDA vs oracle / fixtures) tests."""
import os

import numpy as np

from pathed_b200._binding import ISECT_DTYPE, REPO_ROOT, rays_array

GOLDEN = os.path.join(REPO_ROOT, "tests", "golden")
REL = 1e-5  # north_star: BSDF eval/pdf/sample and t within 1e-5 relative


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_err(a, b, floor=1e-6):
    """elementwise |a-b| / max(|b|, floor*scale); scale = magnitude of the whole tuple for vectors"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    scale = np.maximum(np.abs(b), floor)
    if b.ndim > 1:
        scale = np.maximum(scale, np.linalg.norm(b, axis=-1, keepdims=True))
    both_nan = np.isnan(a) & np.isnan(b)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    err = np.abs(a - b) / scale
    return np.where(both_nan | both_inf, 0.0, err)


def frac_within(a, b, tol=REL, floor=1e-6):
    e = rel_err(a, b, floor)
    if e.ndim > 1:
        e = e.max(axis=-1)
    return float((e <= tol).mean()), e


def make_isects(wo, ng, ns, uv, material=0):
    out = np.zeros(len(wo), ISECT_DTYPE)
    out["hit"] = 1; out["t"] = 1.0
    out["wo"] = wo; out["normal"] = ng; out["shading_normal"] = ns; out["uv"] = uv; out["material"] = material
    return out


def to_rays(arr6):
    return rays_array(arr6[:, :3], arr6[:, 3:])


def rel_mse(x, r):
    """mean over pixels and channels of (x-r)^2 / (r^2 + 1e-2)"""
    x = np.asarray(x, np.float64); r = np.asarray(r, np.float64)
    return float(np.mean((x - r) ** 2 / (r ** 2 + 1e-2)))
