#!/usr/bin/env python3
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref, built by
oracle/ref/build_ref.sh from /root/reference) on seeded inputs.

Runs only in the build container (needs oracle/_ref); the fixtures it writes are committed and are what
the CPU tests pin the oracle against and what the GPU tests compare the CUDA path with.

    python tools/make_golden.py            # everything
    python tools/make_golden.py bsdf rays  # selected groups: bsdf lights rays radiance images

One process per scene: the reference keeps its Embree scene and Job in process globals.
"""
import ctypes
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")

from golden_inputs import (BSDF_CONFIGS, SCENES, bsdf_inputs, light_inputs, make_test_texture, material_params, png_variants,  # noqa: E402
                           ray_inputs, uniform_floats, write_png)

PROBE = os.path.join(ROOT, "oracle", "_ref", "libpathed_ref_probe.so")
HEADLESS = os.path.join(ROOT, "oracle", "_ref", "pathed_ref_headless")


def fptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def probe():
    lib = ctypes.CDLL(PROBE)
    lib.ref_material_new.restype = ctypes.c_void_p
    lib.ref_material_new_textured.restype = ctypes.c_void_p
    lib.ref_env_new.restype = ctypes.c_void_p
    return lib


def gen_bsdf():
    lib = probe()
    texture_png = os.path.join(GOLDEN, "texture_test.png")
    write_png(texture_png, make_test_texture())  # decoded by the reference's stb_image inside the probe
    for name, cfg in BSDF_CONFIGS.items():
        n = 512
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
        frame = np.zeros((n, 9), np.float32)
        lib.ref_tangent_frame(n, fptr(ns), fptr(wo), fptr(frame))
        np.savez_compressed(os.path.join(GOLDEN, "bsdf_%s.npz" % name), f=f, pdf=pdf, sample_wi=swi, sample_pdf=spdf,
                            sample_throughput=sthr, consumed=used, frame=frame)
        print("bsdf", name, "f mean", f.mean(), "nonzero", (f.sum(1) != 0).mean())


def gen_images_decode():
    """What the reference's stb_image returns for every PNG variant (and a binary PPM): the decoder's known answers."""
    import tempfile
    lib = probe()
    out = {}
    files = dict(png_variants())
    files["ppm_p6"] = b"P6\n# comment\n5 3\n255\n" + bytes(range(45))
    files["pgm_p5"] = b"P5 4 2 255\n" + bytes(range(100, 108))
    with tempfile.TemporaryDirectory() as tmp:
        for name, data in files.items():
            path = os.path.join(tmp, name)
            open(path, "wb").write(data)
            w, h = ctypes.c_int(), ctypes.c_int()
            assert lib.ref_load_image(path.encode(), None, 0, ctypes.byref(w), ctypes.byref(h)) == 0, name
            rgb = np.zeros((h.value, w.value, 3), np.uint8)
            lib.ref_load_image(path.encode(), fptr(rgb), rgb.size, ctypes.byref(w), ctypes.byref(h))
            out[name] = rgb
            print("decode", name, rgb.shape, int(rgb.sum()))
    np.savez_compressed(os.path.join(GOLDEN, "image_decode.npz"), **out)


def gen_lights():
    lib = probe()
    n = 512
    tri, sph, ref, xi2 = light_inputs(n)
    out = {}
    p = np.zeros((n, 3), np.float32); nr = np.zeros((n, 3), np.float32); inv = np.zeros(n, np.float32); meas = np.zeros(n, np.int32)
    lib.ref_triangle_sample(fptr(tri), n, fptr(ref), fptr(xi2), fptr(p), fptr(nr), fptr(inv), fptr(meas))
    pdf = np.zeros(n, np.float32)
    lib.ref_triangle_pdf(fptr(tri), n, fptr(p), fptr(ref), fptr(pdf))
    out.update(tri_point=p.copy(), tri_normal=nr.copy(), tri_inv_pdf=inv.copy(), tri_measure=meas.copy(), tri_pdf=pdf.copy())
    lib.ref_sphere_sample(fptr(sph), n, fptr(ref), fptr(xi2), fptr(p), fptr(nr), fptr(inv), fptr(meas))
    lib.ref_sphere_pdf(fptr(sph), n, fptr(p), fptr(ref), fptr(pdf))
    out.update(sph_point=p.copy(), sph_normal=nr.copy(), sph_inv_pdf=inv.copy(), sph_measure=meas.copy(), sph_pdf=pdf.copy())
    np.savez_compressed(os.path.join(GOLDEN, "lights_shapes.npz"), **out)
    print("lights: shapes done; sphere inside fraction", (meas == 1).mean())


SCENE_WORKER = r"""
import ctypes, os, sys
import numpy as np
ROOT = %(root)r
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from golden_inputs import SCENES, ray_inputs, uniform_floats
lib = ctypes.CDLL(%(probe)r)
fptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
name = %(name)r
cfg = SCENES[name]
rc = lib.ref_init(ROOT.encode(), cfg['scene'].encode(), cfg['width'], cfg['height'], 0, cfg['last_bounce'])
assert rc == 0, rc
out = {}
out['num_lights'] = np.array(lib.ref_num_lights())
n = cfg['n_rays']
row_col = ray_inputs(name, n)
cam = np.zeros((n, 6), np.float32)
lib.ref_camera_rays(n, fptr(row_col), fptr(cam))

def trace(rays):
    m = len(rays)
    t = np.zeros(m, np.float32); g = np.zeros(m, np.uint32); p = np.zeros(m, np.uint32)
    uv = np.zeros((m, 2), np.float32); ng = np.zeros((m, 3), np.float32)
    lib.ref_intersect_raw(m, fptr(rays), fptr(t), fptr(g), fptr(p), fptr(uv), fptr(ng))
    hit = np.zeros(m, np.int32); t2 = np.zeros(m, np.float32); pt = np.zeros((m, 3), np.float32)
    nn = np.zeros((m, 3), np.float32); ns = np.zeros((m, 3), np.float32); tuv = np.zeros((m, 2), np.float32)
    em = np.zeros((m, 3), np.float32); dl = np.zeros(m, np.int32)
    lib.ref_intersect(m, fptr(rays), fptr(hit), fptr(t2), fptr(pt), fptr(nn), fptr(ns), fptr(tuv), fptr(em), fptr(dl))
    res = dict(t=t, geom=g, prim=p, bary=uv, ng=ng, hit=hit, point=pt, normal=nn, shading_normal=ns, tex_uv=tuv, emit=em, delta=dl)
    if cfg.get('instanced'):  # RTCHit::instID (SURVEY N4)
        inst = np.zeros((m, 2), np.uint32)
        lib.ref_intersect_inst(m, fptr(rays), fptr(inst))
        res['inst'] = inst
    return res

first = trace(cam)
for k, v in first.items(): out['cam_' + k] = v
out['cam_rays'] = cam
# secondary rays: from the camera hit points, cosine-ish directions about the shading normal
hit = first['hit'] == 1
xi = uniform_floats(cfg['seed'] + 17, (n, 2))
ns = first['shading_normal']; nn = np.where(np.abs(ns[:, :1]) > 0.9, np.array([[0, 1, 0]], np.float32), np.array([[1, 0, 0]], np.float32))
tx = np.cross(ns, nn); tx /= np.maximum(np.linalg.norm(tx, axis=1, keepdims=True), 1e-20); tz = np.cross(ns, tx)
r = np.sqrt(xi[:, :1]); phi = 2 * np.pi * xi[:, 1:]
d = (r * np.cos(phi)) * tx + np.sqrt(1 - xi[:, :1]) * ns + (r * np.sin(phi)) * tz
sec = np.zeros((n, 6), np.float32)
sec[:, :3] = first['point']; sec[:, 3:] = d
sec[~hit] = cam[~hit]
sec = sec.astype(np.float32)
second = trace(sec)
for k, v in second.items(): out['sec_' + k] = v
out['sec_rays'] = sec
# shadow segments between first and second hit points of a permuted partner
perm = np.roll(np.arange(n), 7)
a = first['point']; b = second['point'][perm]
both = hit & (second['hit'][perm] == 1)
seg = b - a; dist = np.linalg.norm(seg, axis=1)
ok = both & (dist > 1e-2)
sh = np.zeros((n, 6), np.float32); sh[:, :3] = a; sh[:, 3:] = seg / np.maximum(dist[:, None], 1e-20)
sh[~ok] = cam[~ok]; dist = np.where(ok, dist, 50.0).astype(np.float32)
occ = np.zeros(n, np.uint8)
lib.ref_occluded(n, fptr(sh), fptr(dist), fptr(occ))
out['shadow_rays'] = sh.astype(np.float32); out['shadow_max_t'] = dist; out['shadow_occluded'] = occ
if 'integrator' in cfg:
    # volumetric queries (participating media): containers are filtered out and leave volume events
    ME = 8
    sh32 = np.ascontiguousarray(sh.astype(np.float32))
    vocc = np.zeros(n, np.uint8); vne = np.zeros(n, np.int32); vet = np.zeros((n, ME), np.float32)
    lib.ref_volumetric_occluded(n, fptr(sh32), fptr(dist), fptr(vocc), fptr(vne), fptr(vet), ME)
    out['vshadow_occluded'] = vocc; out['vshadow_n_events'] = vne; out['vshadow_event_t'] = vet
    for tag, rays in (('vcam', cam), ('vsec', sec)):
        vh = np.zeros(n, np.int32); vt = np.zeros(n, np.float32); vp = np.zeros((n, 3), np.float32); ve = np.zeros((n, 3), np.float32)
        ne = np.zeros(n, np.int32); et = np.zeros((n, ME), np.float32)
        lib.ref_volumetric_intersect(n, fptr(rays), fptr(vh), fptr(vt), fptr(vp), fptr(ve), fptr(ne), fptr(et), ME)
        out[tag + '_hit'] = vh; out[tag + '_t'] = vt; out[tag + '_point'] = vp; out[tag + '_emit'] = ve
        out[tag + '_n_events'] = ne; out[tag + '_event_t'] = et
    lib.ref_set_integrator(cfg['integrator'])
# light sampling from the hit points
nl = int(out['num_lights'])
if nl > 0:
    m = min(n, 1024)
    xi3 = uniform_floats(cfg['seed'] + 29, (m, 3))
    ref = np.ascontiguousarray(first['point'][:m])
    p = np.zeros((m, 3), np.float32); nr = np.zeros((m, 3), np.float32); inv = np.zeros(m, np.float32)
    meas = np.zeros(m, np.int32); sap = np.zeros(m, np.float32); em = np.zeros((m, 3), np.float32)
    lib.ref_scene_sample_direct_lights(m, fptr(ref), fptr(xi3), fptr(p), fptr(nr), fptr(inv), fptr(meas), fptr(sap), fptr(em))
    out.update(ls_ref=ref, ls_point=p, ls_normal=nr, ls_inv_pdf=inv, ls_measure=meas, ls_solid_angle_pdf=sap, ls_emit=em)
    lp = np.zeros(n, np.float32)
    lib.ref_scene_lights_pdf(n, fptr(sec), fptr(lp))
    out['sec_light_pdf'] = lp
env = np.zeros((n, 3), np.float32)
dirs = np.ascontiguousarray(sec[:, 3:])
lib.ref_scene_environment(n, fptr(dirs), fptr(env))
out['sec_env_radiance'] = env
# whole paths with a replayed random stream
m = cfg['n_paths']
stride = 96
xi = uniform_floats(cfg['seed'] + 101, (m, stride))
rgb = np.zeros((m, 3), np.float32); used = np.zeros(m, np.int32)
prim = np.ascontiguousarray(cam[:m])
lib.ref_radiance(m, fptr(prim), fptr(xi), stride, fptr(rgb), fptr(used))
out['path_rgb'] = rgb; out['path_consumed'] = used
np.savez_compressed(%(out)r, **out)
print(name, 'hit rate', hit.mean(), 'occluded', occ.mean(), 'lights', nl, 'path mean', rgb.mean(0), 'max consumed', used.max())
"""


def gen_scenes(names):
    for name in names:
        out = os.path.join(GOLDEN, "scene_%s.npz" % name)
        code = SCENE_WORKER % dict(root=ROOT, probe=PROBE, name=name, out=out)
        subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)


def gen_images(names):
    for name in names:
        cfg = SCENES[name]
        if not cfg.get("image_spp"):
            continue
        w, h, spp = cfg["image_width"], cfg["image_height"], cfg["image_spp"]
        with tempfile.TemporaryDirectory() as tmp:
            from golden_inputs import INTEGRATOR_NAMES
            job = {"spp": spp, "integrator": INTEGRATOR_NAMES[cfg.get("integrator", 0)], "scene": cfg["scene"], "startBounce": 0,
                   "lastBounce": cfg["last_bounce"], "output_directory": os.path.join(tmp, "out"), "showUI": False,
                   "force": True, "width": w, "height": h, "output_name": "golden"}
            job_path = os.path.join(tmp, "job.json")
            with open(job_path, "w") as f:
                json.dump(job, f)
            raw = os.path.join(tmp, "image.f32")
            log = subprocess.run([HEADLESS, "--root", ROOT, job_path, "--raw", raw], check=True, capture_output=True, text=True).stdout
            result = [l for l in log.splitlines() if l.startswith("REF_RESULT")][-1]
            img = np.fromfile(raw, np.float32).reshape(h, w, 3)[::-1].copy()  # back to row 0 = bottom
        np.savez_compressed(os.path.join(GOLDEN, "image_%s.npz" % name), image=img.astype(np.float16), spp=spp)
        print("image", name, result, "mean", img.mean((0, 1)))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    groups = sys.argv[1:] or ["bsdf", "lights", "decode", "scenes", "images"]
    if "bsdf" in groups:
        gen_bsdf()
    if "lights" in groups:
        gen_lights()
    if "decode" in groups:
        gen_images_decode()
    if "scenes" in groups:
        gen_scenes(list(SCENES))
    if "images" in groups:
        gen_images(list(SCENES))
    for g in groups:
        if g.startswith("scene:"):
            gen_scenes([g[6:]])
        if g.startswith("image:"):
            gen_images([g[6:]])


if __name__ == "__main__":
    main()
