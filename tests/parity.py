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


def check_container_known_answer(api):
    """scenes/cornell-medium.json, the ray down the view axis from the camera: the container box (z = +-0.9, Passthrough + medium)
    is an ordinary hit for Scene::testIntersect (t = 5.9) and is skipped by the volumetric queries, which report the glass sphere
    (radius 0.3 at the origin of the xz plane: t = 6.5) and the two container faces as volume events (t = 5.9 and 7.7: both lie in
    front of the back wall, the closest TRIANGLE hit, which is all Embree's per-type traversal knows when it meets them)."""
    rays = rays_array([[0.0, 1.0, 6.8]], [[0.0, 0.0, -1.0]])
    assert abs(float(api.intersect_full(rays)["t"][0]) - 5.9) < 1e-5
    isects, ne, et, em = api.intersect_volumetric(rays)
    assert isects["hit"][0] == 1 and abs(float(isects["t"][0]) - 6.5) < 1e-5
    assert ne[0] == 2 and abs(float(et[0, 0]) - 5.9) < 1e-5 and abs(float(et[0, 1]) - 7.7) < 1e-5 and (em[0, :2] == 0).all()
    occ, ne, et, em = api.occluded_volumetric(rays, np.array([6.0], np.float32))
    assert occ[0] == 0 and ne[0] == 1 and abs(float(et[0, 0]) - 5.9) < 1e-5      # up to just before the sphere: one event
    occ, ne, et, em = api.occluded_volumetric(rays, np.array([7.0], np.float32))
    assert occ[0] == 1 and ne[0] == 0                                              # the glass sphere occludes
    assert api.occluded(rays, np.array([6.0], np.float32))[0] == 0               # Scene::testOcclusion skips the container too


def half_like_reference(x):
    """float32 -> float16 the way the reference's EXR writer does it (tinyexr float_to_half_full, vendor/tinyexr.h:7164-7199): truncate the
    mantissa, add one when the highest dropped bit is set (ties go up), flush float denormals; returned as float32 like read_exr gives"""
    x = np.ascontiguousarray(x, np.float32)
    bits = x.view(np.uint32).astype(np.int64)
    sign = (bits >> 16) & 0x8000
    biased = (bits >> 23) & 0xFF
    mant = bits & 0x7FFFFF
    exp = biased - 127 + 15
    normal = (exp << 10 | (mant >> 13)) + ((mant >> 12) & 1)
    full = mant | 0x800000
    shift = np.clip(14 - exp, 0, 40)
    under = np.where(shift <= 24, (full >> shift) + ((full >> np.clip(shift - 1, 0, 40)) & 1), 0)
    half = np.where(exp >= 31, 0x7C00, np.where(exp <= 0, under, normal))
    half = np.where(biased == 0, 0, np.where(biased == 0xFF, 0x7C00 | np.where(mant != 0, 0x200, 0), half))
    return (sign | half).astype(np.uint16).view(np.float16).astype(np.float32).reshape(x.shape)
