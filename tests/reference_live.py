"""Large seeded batches for the parity gates of SURVEY 8(d): 2^20 rays per class and scene, 2^16 BSDF tuples per material, and the
2^16 adversarial rays through mesh vertices and edge midpoints.

The checker is the compiled, UNMODIFIED reference (oracle/_ref/libpathed_ref_probe.so: Embree 3.6.0 + the reference's Scene / Material
classes) wherever it exists -- it is a built artefact like the product's own .so files and travels to the GPU box -- and the pinned CPU
oracle (oracle/liboracle.so) otherwise.  Test infrastructure only: nothing in pathed_b200/ imports this module.

The reference keeps its Embree scene and its Job in process globals, so every scene batch runs in its own python process
(`python tests/reference_live.py scene <name> <n> <out.npz>`).
"""
import ctypes
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from golden_inputs import BSDF_CONFIGS, bsdf_inputs, material_params, uniform_floats  # noqa: E402

PROBE = os.path.join(ROOT, "oracle", "_ref", "libpathed_ref_probe.so")

# BASELINE.json's five configurations at their own resolutions (C1..C5 of SURVEY 8)
FULL_SCENES = {
    "cornell": dict(scene="scenes/cornell.json", width=512, height=512, last_bounce=10, seed=211),
    "cornell_glass": dict(scene="scenes/cornell-glass.json", width=512, height=512, last_bounce=10, seed=212),
    "dragon": dict(scene="scenes/dragon.json", width=1024, height=1024, last_bounce=10, seed=213),
    "mis": dict(scene="scenes/mis-pbrt.json", width=768, height=512, last_bounce=10, seed=214),
    "teapot": dict(scene="scenes/teapot.json", width=1920, height=1080, last_bounce=10, seed=215),
}


def have_probe():
    return os.path.exists(PROBE)


def fptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def film_positions(cfg, n):
    """(row, col) film positions: pixel centres of a scan over the whole image plus the box-filter jitter"""
    u = uniform_floats(cfg["seed"], (n, 2))
    rc = np.stack([u[:, 0] * np.float32(cfg["height"]) - np.float32(0.5), u[:, 1] * np.float32(cfg["width"]) - np.float32(0.5)], 1)
    return np.ascontiguousarray(rc.astype(np.float32))


def secondary_rays(cfg, cam, hit, point, shading_normal):
    """cosine-hemisphere directions about the shading normal from the camera hits (class (ii)); misses repeat the camera ray"""
    n = len(cam)
    xi = uniform_floats(cfg["seed"] + 17, (n, 2))
    ns = shading_normal.astype(np.float32)
    other = np.where(np.abs(ns[:, :1]) > 0.9, np.array([[0, 1, 0]], np.float32), np.array([[1, 0, 0]], np.float32))
    tx = np.cross(ns, other)
    tx /= np.maximum(np.linalg.norm(tx, axis=1, keepdims=True), 1e-20)
    tz = np.cross(ns, tx)
    r = np.sqrt(xi[:, :1])
    phi = 2 * np.pi * xi[:, 1:]
    d = (r * np.cos(phi)) * tx + np.sqrt(1 - xi[:, :1]) * ns + (r * np.sin(phi)) * tz
    sec = np.zeros((n, 6), np.float32)
    sec[:, :3] = point
    sec[:, 3:] = d
    sec[~hit] = cam[~hit]
    return np.ascontiguousarray(sec.astype(np.float32))


def shadow_segments(cam, hit, point, light_point):
    """class (iii), what directSampleLights traces: from a surface point to the light point the scene's own light sampling drew for
    it (src/path_tracer.cpp:113-151); rays without a surface hit repeat the camera ray with a 50-unit interval"""
    n = len(cam)
    seg = light_point.astype(np.float32) - point.astype(np.float32)
    dist = np.linalg.norm(seg, axis=1)
    ok = hit & (dist > 1e-2) & np.isfinite(dist)
    sh = np.zeros((n, 6), np.float32)
    sh[:, :3] = point
    sh[:, 3:] = seg / np.maximum(dist[:, None], 1e-20)
    sh[~ok] = cam[~ok]
    dist = np.where(ok, dist, 50.0).astype(np.float32)
    return np.ascontiguousarray(sh.astype(np.float32)), dist


def adversarial_rays(scene_file, origin, n, seed):
    """class (iv): rays from the camera position exactly through mesh vertices (first half) and edge midpoints (second half)"""
    pts = []
    counts = scene_file.counts()
    for g in range(counts["geometries"]):
        geo = scene_file.geometry(g)
        if geo is None:
            continue
        pos, idx, _ = geo
        if len(idx) == 0:
            continue
        pts.append((pos, idx))
    if not pts:
        return np.zeros((0, 6), np.float32)
    half = n // 2
    pick = (uniform_floats(seed, (n, 3)))
    sizes = np.array([len(i) for _, i in pts], np.float64)
    cum = np.cumsum(sizes) / sizes.sum()
    out = np.zeros((n, 6), np.float32)
    out[:, :3] = np.asarray(origin, np.float32)
    which = np.searchsorted(cum, pick[:, 0].astype(np.float64), side="right").clip(0, len(pts) - 1)
    for m, (pos, idx) in enumerate(pts):
        sel = np.where(which == m)[0]
        if len(sel) == 0:
            continue
        tri = idx[(pick[sel, 1] * len(idx)).astype(np.int64).clip(0, len(idx) - 1)]
        corner = (pick[sel, 2] * 3).astype(np.int64).clip(0, 2)
        v0 = pos[tri[np.arange(len(sel)), corner]]
        v1 = pos[tri[np.arange(len(sel)), (corner + 1) % 3]]
        target = np.where((sel < half)[:, None], v0, (v0 + v1) * np.float32(0.5))
        d = target - out[sel, :3]
        out[sel, 3:] = d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20)
    return np.ascontiguousarray(out.astype(np.float32))


# ------------------------------------------------------------------------------------------------ reference side (probe)
def _probe():
    lib = ctypes.CDLL(PROBE)
    lib.ref_material_new.restype = ctypes.c_void_p
    lib.ref_material_new_textured.restype = ctypes.c_void_p
    return lib


def reference_bsdf(name, n, texture_png=None):
    """f, pdf, sample_wi, sample_pdf, sample_throughput of the reference's Material for bsdf_inputs(name, n)"""
    lib = _probe()
    cfg = BSDF_CONFIGS[name]
    wo, ng, ns, uv, wi, xi = bsdf_inputs(name, n)
    params = material_params(cfg)
    if cfg.get("textured"):
        mat = ctypes.c_void_p(lib.ref_material_new_textured(ctypes.c_int(cfg["type"]), fptr(params), texture_png.encode()))
    else:
        mat = ctypes.c_void_p(lib.ref_material_new(ctypes.c_int(cfg["type"]), fptr(params)))
    f = np.zeros((n, 3), np.float32); pdf = np.zeros(n, np.float32)
    lib.ref_bsdf_eval(mat, n, fptr(wo), fptr(ng), fptr(ns), fptr(uv), fptr(wi), fptr(f), fptr(pdf))
    swi = np.zeros((n, 3), np.float32); spdf = np.zeros(n, np.float32); sthr = np.zeros((n, 3), np.float32)
    used = np.zeros(n, np.int32)
    lib.ref_bsdf_sample(mat, n, fptr(wo), fptr(ng), fptr(ns), fptr(uv), fptr(xi), fptr(swi), fptr(spdf), fptr(sthr), fptr(used))
    return dict(f=f, pdf=pdf, sample_wi=swi, sample_pdf=spdf, sample_throughput=sthr)


def _trace(lib, rays):
    m = len(rays)
    t = np.zeros(m, np.float32); g = np.zeros(m, np.uint32); p = np.zeros(m, np.uint32)
    uv = np.zeros((m, 2), np.float32); ng = np.zeros((m, 3), np.float32)
    lib.ref_intersect_raw(m, fptr(rays), fptr(t), fptr(g), fptr(p), fptr(uv), fptr(ng))
    hit = np.zeros(m, np.int32); t2 = np.zeros(m, np.float32); pt = np.zeros((m, 3), np.float32)
    nn = np.zeros((m, 3), np.float32); ns = np.zeros((m, 3), np.float32); tuv = np.zeros((m, 2), np.float32)
    em = np.zeros((m, 3), np.float32); dl = np.zeros(m, np.int32)
    lib.ref_intersect(m, fptr(rays), fptr(hit), fptr(t2), fptr(pt), fptr(nn), fptr(ns), fptr(tuv), fptr(em), fptr(dl))
    return dict(t=t, geom=g, prim=p, hit=hit == 1, point=pt, shading_normal=ns)


def _scene_worker(name, n, out_path, adversarial_path=None):
    cfg = FULL_SCENES[name]
    lib = _probe()
    rc = lib.ref_init(ROOT.encode(), cfg["scene"].encode(), cfg["width"], cfg["height"], 0, cfg["last_bounce"])
    assert rc == 0, rc
    out = {}
    cam = np.zeros((n, 6), np.float32)
    lib.ref_camera_rays(n, fptr(film_positions(cfg, n)), fptr(cam))
    first = _trace(lib, cam)
    sec = secondary_rays(cfg, cam, first["hit"], first["point"], first["shading_normal"])
    second = _trace(lib, sec)
    lp = np.zeros((n, 3), np.float32); nr = np.zeros((n, 3), np.float32); inv = np.zeros(n, np.float32)
    meas = np.zeros(n, np.int32); sap = np.zeros(n, np.float32); em = np.zeros((n, 3), np.float32)
    ref_pts = np.ascontiguousarray(first["point"])
    lib.ref_scene_sample_direct_lights(n, fptr(ref_pts), fptr(uniform_floats(cfg["seed"] + 29, (n, 3))), fptr(lp), fptr(nr), fptr(inv),
                                       fptr(meas), fptr(sap), fptr(em))
    sh, dist = shadow_segments(cam, first["hit"], first["point"], lp)
    occ = np.zeros(n, np.uint8)
    lib.ref_occluded(n, fptr(sh), fptr(dist), fptr(occ))
    for tag, rays, res in (("cam", cam, first), ("sec", sec, second)):
        out[tag + "_rays"] = rays
        for k in ("t", "geom", "prim"):
            out[tag + "_" + k] = res[k]
    out["shadow_rays"] = sh; out["shadow_max_t"] = dist; out["shadow_occluded"] = occ
    if adversarial_path:
        adv = np.load(adversarial_path)
        res = _trace(lib, adv)
        out["adv_rays"] = adv
        for k in ("t", "geom", "prim"):
            out["adv_" + k] = res[k]
    np.savez(out_path, **out)


def reference_scene_batch(name, n, adversarial=None):
    """Embree's answers for the three ray classes (and the adversarial batch) of one scene; None without the probe"""
    if not have_probe():
        return None
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "batch.npz")
        cmd = [sys.executable, os.path.abspath(__file__), "scene", name, str(n), out]
        if adversarial is not None and len(adversarial):
            adv = os.path.join(tmp, "adv.npy")
            np.save(adv, adversarial)
            cmd.append(adv)
        subprocess.check_call(cmd, cwd=ROOT)
        with np.load(out) as z:
            return {k: z[k] for k in z.files}


# ------------------------------------------------------------------------------------------------ oracle side (fallback checker)
def oracle_scene_batch(name, n, adversarial=None):
    """the same batch answered by the pinned CPU oracle (binned-BVH traversal in Embree's arithmetic)"""
    from oracle_binding import oracle_scene
    from parity import to_rays
    cfg = FULL_SCENES[name]
    o = oracle_scene(cfg["scene"], cfg["width"], cfg["height"])
    o.set_option("brute_force", 0)

    def trace(rays6):
        rays = to_rays(rays6)
        h = o.intersect(rays)
        full = o.intersect_full(rays)
        return dict(t=h["t"], geom=h["geom_id"], prim=h["prim_id"], hit=h["geom_id"] != 0xFFFFFFFF, point=full["point"],
                    shading_normal=full["shading_normal"])

    def six(r):
        return np.ascontiguousarray(np.concatenate([r["origin"], r["direction"]], 1))
    cam = six(o.camera_rays(film_positions(cfg, n)))
    first = trace(cam)
    sec = secondary_rays(cfg, cam, first["hit"], first["point"], first["shading_normal"])
    second = trace(sec)
    ls = o.light_sample(np.ascontiguousarray(first["point"]), uniform_floats(cfg["seed"] + 29, (n, 3)))
    sh, dist = shadow_segments(cam, first["hit"], first["point"], ls["point"])
    out = {}
    for tag, rays, res in (("cam", cam, first), ("sec", sec, second)):
        out[tag + "_rays"] = rays
        for k in ("t", "geom", "prim"):
            out[tag + "_" + k] = res[k]
    out["shadow_rays"] = sh; out["shadow_max_t"] = dist
    out["shadow_occluded"] = o.occluded(to_rays(sh), dist)
    if adversarial is not None and len(adversarial):
        res = trace(adversarial)
        out["adv_rays"] = adversarial
        for k in ("t", "geom", "prim"):
            out["adv_" + k] = res[k]
    return out


def compare_hits(got, want_t, want_geom, want_prim, rel=1e-5):
    """north_star's intersection gate on one batch.  Returns fractions: hit/miss agreement, same primitive, ties (another primitive
    at the same depth: shared edge, vertex or coincident face), t within `rel` among the rays both sides hit."""
    ref_hit = want_geom != 0xFFFFFFFF
    got_hit = got["geom_id"] != 0xFFFFFFFF
    agree = ref_hit == got_hit
    both = agree & ref_hit
    same = both & (got["geom_id"] == want_geom) & (got["prim_id"] == want_prim)
    t_err = np.abs(got["t"].astype(np.float64) - want_t.astype(np.float64)) / np.maximum(np.abs(want_t.astype(np.float64)), 1e-6)
    t_ok = t_err <= rel
    tie = both & ~same & t_ok
    n = float(len(want_t))
    return dict(n=int(n), hit_miss=float(agree.mean()), same_prim=float((~both | same)[agree].mean()) if agree.any() else 1.0,
                prim_or_tie=float(((agree & ~ref_hit) | same | tie).mean()), ties=float(tie.mean()),
                t_within=float(t_ok[both].mean()) if both.any() else 1.0, t_max=float(t_err[both & same].max()) if (both & same).any() else 0.0,
                hit_rate=float(ref_hit.mean()))


if __name__ == "__main__":
    if sys.argv[1] == "scene":
        _scene_worker(sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else None)
