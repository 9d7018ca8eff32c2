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


def check_volumetric_queries(o, g):
    occ, ne, et, em = o.occluded_volumetric(to_rays(g["shadow_rays"]), g["shadow_max_t"])
    assert (occ == g["vshadow_occluded"]).mean() >= 0.999
    same = occ == g["vshadow_occluded"]
    assert (ne[same] == g["vshadow_n_events"][same]).mean() >= 0.999, (ne[same] == g["vshadow_n_events"][same]).mean()
    both = same & (ne == g["vshadow_n_events"])
    assert frac_within(et[both], g["vshadow_event_t"][both], floor=1e-4)[0] >= 0.999
    assert ((em[both] != 0xFFFFFFFF).sum(1) == np.minimum(ne[both], et.shape[1])).all()
    for tag in ("vcam", "vsec"):
        rays = to_rays(g["cam_rays"] if tag == "vcam" else g["sec_rays"])
        isects, ne, et, em = o.intersect_volumetric(rays)
        agree = (isects["hit"] == g[tag + "_hit"])
        assert agree.mean() >= 0.9999
        hit = agree & (isects["hit"] == 1)
        assert (rel_err(isects["t"][hit], g[tag + "_t"][hit]) <= REL).mean() >= 0.9999
        assert frac_within(isects["point"][hit], g[tag + "_point"][hit], tol=5e-5)[0] >= 0.9999  # points near the origin: 1 ulp of t
        # events behind the final hit depend on Embree's traversal order (only the camera-ray branch of samplePixel reads
        # them): the restatement keeps those in front of the hit, which is what Embree reports on almost every ray
        same_n = ne == g[tag + "_n_events"]
        assert same_n[agree].mean() >= 0.99, (tag, same_n[agree].mean())
        both = agree & same_n
        # on a miss the reference returns the events in traversal order (it sorts only in the hit branch, src/scene.cpp:337-343
        # vs :347-351); rayTransmission only uses |t1 - t0| of a pair, so the comparison sorts them
        want = g[tag + "_event_t"].copy()
        for i in np.where(g[tag + "_hit"] == 0)[0]:
            k = min(int(g[tag + "_n_events"][i]), want.shape[1])
            want[i, :k] = np.sort(want[i, :k])
        assert frac_within(et[both], want[both], floor=1e-4)[0] >= 0.999
