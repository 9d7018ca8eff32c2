#!/usr/bin/env python3
"""Generates every scene file and asset the tests and the benchmark use, deterministically.

The reference's five benchmark configurations (BASELINE.json) name scene files whose meshes and
environment maps live in a git-ignored `assets/` directory that is NOT part of the reference
(SURVEY.md F1).  This script writes the scene descriptions in Pathed's own scene-JSON dialect plus
stand-in assets built only from formulas and fixed seeds, so that the reference oracle (oracle/_ref)
and the CUDA path read byte-identical inputs:

  scenes/cornell.json            Cornell box (public-domain McGuire/Cornell measurements), OBJ + MTL,
                                 `f -4 -3 -2 -1` quads, 36 triangles, 2 emissive triangles
  scenes/cornell-glass.json      Cornell shell with per-vertex normals (v/vt/vn faces), mirror box,
                                 1088-triangle glass ball (default ior 1.4), translate (0,0,1)
  scenes/dragon.json             floor quad + ~0.87 M-triangle procedural knot as `assets/dragon.obj`
                                 (stand-in for the Stanford dragon), plastic/Beckmann 0.1, env light only
  scenes/mis-pbrt.json           Veach MIS arrangement: 5 emissive spheres, 4 plastic plates, floor (PLY)
  scenes/teapot.json             glass surface-of-revolution "teapot" body + lid, checkerboard quad, env map
  test_scenes/1_pixel_test.exr   1000x500 environment map, one texel = 10000 at row 239, col 753
  test_scenes/environment_map_sampling.json   white quad lit only by that map
  scenes/instanced.json          SURVEY N4: `instance` / `instanced` models, two instance levels, rotated / scaled / mirrored placements

Outputs are git-ignored (scenes/, assets/, test_scenes/ are listed in .gitignore) and travel with gpurun.
Usage: python tools/make_assets.py [--root DIR] [--dragon-segments N] [--force]
"""
import argparse
import ctypes
import json
import math
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def s(x):
    """Pathed encodes every scene scalar as a JSON string (parsed with stof)."""
    return repr(float(x)) if not isinstance(x, str) else x


def vec(*xs):
    return [s(x) for x in xs]


# ----------------------------------------------------------------------------------------- writers
def write_text(path, text):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(text)


def write_json(path, obj):
    write_text(path, json.dumps(obj, indent=2) + "\n")


def write_obj(path, vertices, faces, normals=None, header=""):
    """faces: (n,3) int 0-based; with normals -> `v//n` syntax sharing the vertex index."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(header)
        np.savetxt(f, vertices, fmt="v %.6f %.6f %.6f")
        if normals is not None:
            np.savetxt(f, normals, fmt="vn %.6f %.6f %.6f")
            idx = faces + 1
            np.savetxt(f, np.stack([idx[:, 0], idx[:, 0], idx[:, 1], idx[:, 1], idx[:, 2], idx[:, 2]], 1),
                       fmt="f %d//%d %d//%d %d//%d")
        else:
            np.savetxt(f, faces + 1, fmt="f %d %d %d")


def write_ply(path, vertices, faces):
    """binary_little_endian PLY exactly as src/ply_parser.cpp expects it."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\n")
        f.write(b"element vertex %d\nproperty float x\nproperty float y\nproperty float z\n" % len(vertices))
        f.write(b"element face %d\nproperty list uchar int vertex_indices\nend_header\n" % len(faces))
        f.write(np.asarray(vertices, dtype="<f4").tobytes())
        for face in faces:
            f.write(struct.pack("<Biii", 3, *[int(i) for i in face]))


_host = None


def write_exr(path, rgb):
    """rgb: (H,W,3) float32, row 0 = top.  Uses the host library's own EXR writer (fp32, uncompressed)."""
    global _host
    if _host is None:
        _host = ctypes.CDLL(os.path.join(ROOT, "pathed_b200", "libpathed_host.so"))
    os.makedirs(os.path.dirname(path), exist_ok=True)
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    h, w, _ = rgb.shape
    rc = _host.pth_exr_write_rgb_f32(path.encode(), w, h, rgb.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError("EXR write failed: " + path)


# ----------------------------------------------------------------------------------------- cornell
CORNELL_MATERIALS = {
    "leftWall": ((0.63, 0.065, 0.05), (0, 0, 0)),
    "rightWall": ((0.14, 0.45, 0.091), (0, 0, 0)),
    "floor": ((0.725, 0.71, 0.68), (0, 0, 0)),
    "ceiling": ((0.725, 0.71, 0.68), (0, 0, 0)),
    "backWall": ((0.725, 0.71, 0.68), (0, 0, 0)),
    "shortBox": ((0.725, 0.71, 0.68), (0, 0, 0)),
    "tallBox": ((0.725, 0.71, 0.68), (0, 0, 0)),
    "light": ((0.78, 0.78, 0.78), (17, 12, 4)),
}


def mtl_text(materials):
    out = ["# generated by tools/make_assets.py"]
    for name, (kd, ke) in materials.items():
        out += ["", "newmtl %s" % name, "  Ns 10.0000", "  illum 2",
                "  Kd %g %g %g" % kd, "  Ks 0 0 0", "  Ke %g %g %g" % ke]
    return "\n".join(out) + "\n"


def box_quads(top, height, sides, duplicate):
    """Quads of a box standing on y = 0: top face, four sides given as (j, i) corner pairs wound b[j], t[j], t[i], b[i]
    (outward normals), then four unused bottom vertices and a face line that re-references an earlier side
    (`duplicate` = that face's relative index, e.g. -12): the original data set has exactly this quirk, which puts two
    coincident quads into the scene, so it is kept."""
    t = [(x, height, z) for x, z in top]
    b = [(x, 0.0, z) for x, z in top]
    quads = [(t, "-4 -3 -2 -1")]
    for j, i in sides:
        quads.append(([b[j], t[j], t[i], b[i]], "-4 -3 -2 -1"))
    quads.append(([b[0], b[1], b[2], b[3]], "%d %d %d %d" % (duplicate, duplicate + 1, duplicate + 2, duplicate + 3)))
    return quads


def make_cornell(root):
    # Cornell box measurements (Cornell Program of Computer Graphics; OBJ form by Cardenas & McGuire, public domain)
    Q = "-4 -3 -2 -1"
    groups = [
        ("floor", [([(-1.01, 0, 0.99), (1.00, 0, 0.99), (1.00, 0, -1.04), (-0.99, 0, -1.04)], Q)]),
        ("ceiling", [([(-1.02, 1.99, 0.99), (-1.02, 1.99, -1.04), (1.00, 1.99, -1.04), (1.00, 1.99, 0.99)], Q)]),
        ("backWall", [([(-0.99, 0, -1.04), (1.00, 0, -1.04), (1.00, 1.99, -1.04), (-1.02, 1.99, -1.04)], Q)]),
        ("rightWall", [([(1.00, 0, -1.04), (1.00, 0, 0.99), (1.00, 1.99, 0.99), (1.00, 1.99, -1.04)], Q)]),
        ("leftWall", [([(-1.01, 0, 0.99), (-0.99, 0, -1.04), (-1.02, 1.99, -1.04), (-1.02, 1.99, 0.99)], Q)]),
        ("shortBox", box_quads([(0.53, 0.75), (0.70, 0.17), (0.13, 0.00), (-0.05, 0.57)], 0.60,
                               [(3, 2), (0, 3), (1, 0), (2, 1)], -12)),
        ("tallBox", box_quads([(-0.53, 0.09), (0.04, -0.09), (-0.14, -0.67), (-0.71, -0.49)], 1.20,
                              [(0, 3), (3, 2), (2, 1), (1, 0)], -8)),
        ("light", [([(-0.24, 1.98, 0.16), (-0.24, 1.98, -0.22), (0.23, 1.98, -0.22), (0.23, 1.98, 0.16)], Q)]),
    ]
    lines = ["# Cornell box, generated by tools/make_assets.py", "mtllib scenes/CornellBox-Original.mtl", ""]
    for name, quads in groups:
        lines += ["g %s" % name, "usemtl %s" % name]
        for quad, face in quads:
            lines += ["v %.2f %.2f %.2f" % p for p in quad]
            lines.append("f " + face)
        lines.append("")
    write_text(os.path.join(root, "scenes/CornellBox-Original.obj"), "\n".join(lines))
    write_text(os.path.join(root, "scenes/CornellBox-Original.mtl"), mtl_text(CORNELL_MATERIALS))
    write_json(os.path.join(root, "scenes/cornell.json"), {
        "sensor": {"lookAt": {"origin": vec(0, 1, 6.8), "target": vec(0, 1, 0), "up": vec(0, 1, 0)}, "fov": s(19.5)},
        "models": [{"type": "obj", "filename": "scenes/CornellBox-Original.obj"}],
    })


def closed_box(lo, hi):
    """Six quads of an axis-aligned box, wound so that (v1 - v0) x (v2 - v0) points outwards."""
    (x0, y0, z0), (x1, y1, z1) = lo, hi
    return [
        [(x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)],  # +z
        [(x1, y0, z0), (x0, y0, z0), (x0, y1, z0), (x1, y1, z0)],  # -z
        [(x1, y0, z1), (x1, y0, z0), (x1, y1, z0), (x1, y1, z1)],  # +x
        [(x0, y0, z0), (x0, y0, z1), (x0, y1, z1), (x0, y1, z0)],  # -x
        [(x0, y1, z1), (x1, y1, z1), (x1, y1, z0), (x0, y1, z0)],  # +y
        [(x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1)],  # -y
    ]


def make_cornell_medium(root):
    """scenes/cornell-medium.json of the reference (participating medium in a Passthrough container, glass sphere, Cornell
    frame).  Its assets/cornell-volume-caustic/*.obj are not in the reference repository: the container is
    re-made as a box just inside the walls (the light stays outside it), the frame as the Cornell box without its blocks.
    A second scene puts a Passthrough SPHERE container with a tinted medium under an environment light."""
    Q = "-4 -3 -2 -1"
    lines = ["# medium container, generated by tools/make_assets.py", ""]
    for quad in closed_box((-0.9, 0.1, -0.9), (0.9, 1.9, 0.9)):
        lines += ["v %.2f %.2f %.2f" % p for p in quad]
        lines.append("f " + Q)
    write_text(os.path.join(root, "assets/cornell-volume-caustic/bounds.obj"), "\n".join(lines) + "\n")
    groups = [
        ("floor", [(-1.01, 0, 0.99), (1.00, 0, 0.99), (1.00, 0, -1.04), (-0.99, 0, -1.04)]),
        ("ceiling", [(-1.02, 1.99, 0.99), (-1.02, 1.99, -1.04), (1.00, 1.99, -1.04), (1.00, 1.99, 0.99)]),
        ("backWall", [(-0.99, 0, -1.04), (1.00, 0, -1.04), (1.00, 1.99, -1.04), (-1.02, 1.99, -1.04)]),
        ("rightWall", [(1.00, 0, -1.04), (1.00, 0, 0.99), (1.00, 1.99, 0.99), (1.00, 1.99, -1.04)]),
        ("leftWall", [(-1.01, 0, 0.99), (-0.99, 0, -1.04), (-1.02, 1.99, -1.04), (-1.02, 1.99, 0.99)]),
        ("light", [(-0.24, 1.98, 0.16), (-0.24, 1.98, -0.22), (0.23, 1.98, -0.22), (0.23, 1.98, 0.16)]),
    ]
    lines = ["# Cornell box without its blocks, generated by tools/make_assets.py", "mtllib scenes/CornellBox-Original.mtl", ""]
    for name, quad in groups:
        lines += ["g %s" % name, "usemtl %s" % name] + ["v %.2f %.2f %.2f" % p for p in quad] + ["f " + Q, ""]
    write_text(os.path.join(root, "assets/cornell-volume-caustic/CornellBox-Frame.obj"), "\n".join(lines))
    write_json(os.path.join(root, "scenes/cornell-medium.json"), {
        "sensor": {"lookAt": {"origin": vec(0, 1, 6.8), "target": vec(0, 1, 0), "up": vec(0, 1, 0)}, "fov": s(19.5)},
        "media": [{"name": "gas", "type": "homogeneous", "sigma_t": vec(1.0, 1.0, 1.0), "sigma_s": vec(1.0, 1.0, 1.0)}],
        "models": [
            {"type": "obj", "filename": "assets/cornell-volume-caustic/bounds.obj", "internal_medium": "gas", "bsdf": {"type": "passthrough"}},
            {"type": "sphere", "radius": s(0.3), "center": vec(0.0, 1.0, 0.0), "bsdf": {"type": "glass"}},
            {"type": "obj", "filename": "assets/cornell-volume-caustic/CornellBox-Frame.obj"},
        ],
    })
    write_json(os.path.join(root, "test_scenes/medium_sphere.json"), {
        "sensor": {"lookAt": {"origin": vec(0, 2.5, 7), "target": vec(0, 0.8, 0), "up": vec(0, 1, 0)}, "fov": s(35)},
        "media": [{"name": "tea", "type": "homogeneous", "sigma_t": vec(0.5, 0.5, 0.5), "sigma_s": vec(0.4, 0.4, 0.4)},
                  {"name": "unused", "type": "homogeneous", "sigma_t": vec(2, 2, 2)}],
        "models": [
            {"type": "quad", "transform": {"scale": vec(4, 1, 4)}, "bsdf": {"type": "lambertian", "diffuseReflectance": vec(0.6, 0.6, 0.6)}},
            {"type": "sphere", "center": vec(0, 1.2, 0), "radius": s(1.1), "internal_medium": "tea", "bsdf": {"type": "passthrough"}},
            {"type": "sphere", "center": vec(0.2, 1.0, 0.1), "radius": s(0.35),
             "bsdf": {"type": "oren-nayar", "diffuseReflectance": vec(0.7, 0.3, 0.2), "sigma": s(0.3)}},
            {"type": "sphere", "center": vec(-2.2, 0.5, 0.8), "radius": s(0.5), "internal_medium": "tea", "bsdf": {"type": "glass", "ior": s(1.3)}},
            {"type": "sphere", "center": vec(2.4, 2.6, 0.5), "radius": s(0.4),
             "bsdf": {"type": "lambertian", "diffuseReflectance": vec(0, 0, 0), "emit": vec(20, 18, 15)}},
        ],
        "environmentLight": {"filename": "assets/teapot/envmap.exr", "scale": s(0.5)},
    })


def uv_sphere(center, radius, segments=32, rings=17):
    """Lat-long sphere: 2 poles + rings*segments vertices, 2*segments + (rings-1)*segments*2 triangles (1088 for 32x17)."""
    verts = [(0.0, 1.0, 0.0)]
    for r in range(1, rings + 1):
        theta = math.pi * r / (rings + 1)
        for k in range(segments):
            phi = 2 * math.pi * k / segments
            verts.append((math.sin(theta) * math.cos(phi), math.cos(theta), math.sin(theta) * math.sin(phi)))
    verts.append((0.0, -1.0, 0.0))
    faces = []
    ring = lambda r, k: 1 + (r - 1) * segments + (k % segments)
    for k in range(segments):
        faces.append((0, ring(1, k + 1), ring(1, k)))
    for r in range(1, rings):
        for k in range(segments):
            a, b, c, d = ring(r, k), ring(r, k + 1), ring(r + 1, k), ring(r + 1, k + 1)
            faces += [(a, b, d), (a, d, c)]
    last = len(verts) - 1
    for k in range(segments):
        faces.append((last, ring(rings, k), ring(rings, k + 1)))
    n = np.array(verts)
    return n * radius + np.array(center), n, np.array(faces)


def make_cornell_glass(root):
    shell = [  # name, quad corners, per-corner normals (the left wall keeps slightly varying vertex normals)
        ("floor", [(1.0, 0.0, -1.04), (-0.99, 0.0, -1.04), (-1.01, 0.0, 0.99), (1.0, 0.0, 0.99)], [(0, 1, 0)] * 4),
        ("ceiling", [(1.0, 1.59, -1.04), (1.0, 1.59, 0.99), (-1.02, 1.59, 0.99), (-1.02, 1.59, -1.04)], [(0, -1, 0)] * 4),
        ("backWall", [(1.0, 1.59, -1.04), (-1.02, 1.59, -1.04), (-0.99, 0.0, -1.04), (1.0, 0.0, -1.04)], [(0, 0, 1)] * 4),
        ("rightWall", [(1.0, 1.59, 0.99), (1.0, 1.59, -1.04), (1.0, 0.0, -1.04), (1.0, 0.0, 0.99)], [(-1, 0, 0)] * 4),
        ("leftWall", [(-1.02, 1.59, -1.04), (-1.02, 1.59, 0.99), (-1.01, 0.0, 0.99), (-0.99, 0.0, -1.04)],
         [(0.9998, 0.0189, 0.0098), (0.9999, 0.0116, 0.0042), (1.0, 0.0063, 0.0), (0.9999, 0.0135, 0.0057)]),
        ("light", [(0.23, 1.58, -0.22), (0.23, 1.58, 0.16), (-0.24, 1.58, 0.16), (-0.24, 1.58, -0.22)], [(0, -1, 0)] * 4),
    ]
    lines = ["# Cornell shell with vertex normals, generated by tools/make_assets.py",
             "mtllib scenes/cornell-glossy/CornellBox-Glossy.mtl", ""]
    for name, quad, normals in shell:
        lines += ["v %.4f %.4f %.4f" % p for p in quad]
        lines += ["vn %.4f %.4f %.4f" % n for n in normals]
        lines += ["vt 0.0000 0.0000 0.0000", "g %s" % name, "usemtl %s" % name, "s 1"]
        # relative indices; normal k belongs to corner k (normals were pushed in corner order)
        lines.append("f -4/-1/-4 -3/-1/-3 -2/-1/-2")
        lines.append("f -2/-1/-2 -1/-1/-1 -4/-1/-4")
        lines.append("")
    write_text(os.path.join(root, "scenes/cornell-glossy/CornellBox-Glossy.obj"), "\n".join(lines))
    mats = dict(CORNELL_MATERIALS)
    mats["sphere"] = ((0.486, 0.631, 0.663), (0, 0, 0))
    write_text(os.path.join(root, "scenes/cornell-glossy/CornellBox-Glossy.mtl"), mtl_text(mats))

    # mirror box: rotated cuboid standing on the floor, faces as v/vt/vn triangles with one normal per face
    cx, cz, half, height, angle = -0.4363, 0.1437, 0.2355, 0.471, math.radians(39.0)
    ca, sa = math.cos(angle), math.sin(angle)
    corner = lambda dx, dz, y: (cx + ca * dx * half - sa * dz * half, y, cz + sa * dx * half + ca * dz * half)
    b = [corner(-1, -1, 0.0), corner(1, -1, 0.0), corner(1, 1, 0.0), corner(-1, 1, 0.0)]
    t = [corner(-1, -1, height), corner(1, -1, height), corner(1, 1, height), corner(-1, 1, height)]
    quads = [([b[3], b[2], b[1], b[0]], (0, -1, 0)), ([t[0], t[1], t[2], t[3]], (0, 1, 0))]
    for i in range(4):
        j = (i + 1) % 4
        e = np.array(b[j]) - np.array(b[i])
        nrm = np.cross(e, (0, 1, 0)); nrm = nrm / np.linalg.norm(nrm)
        quads.append(([b[i], t[i], t[j], b[j]], tuple(-nrm)))
    lines = ["# mirror box, generated by tools/make_assets.py"]
    vt = ["vt 1.0 0.0 0.0", "vt 1.0 1.0 0.0", "vt 0.0 1.0 0.0", "vt 0.0 0.0 0.0"]
    for quad, nrm in quads:
        lines += ["v %.4f %.4f %.4f" % p for p in quad] + ["vn %.4f %.4f %.4f" % nrm] + vt
        lines += ["f -4/-4/-1 -3/-3/-1 -2/-2/-1", "f -2/-2/-1 -1/-1/-1 -4/-4/-1"]
    write_text(os.path.join(root, "scenes/cornell-glossy/box.obj"), "\n".join(lines) + "\n")

    v, n, f = uv_sphere((0.5353, 0.4569, -0.5610), 0.4376)
    lines = ["# glass ball (32 x 17 lat-long sphere, 1088 triangles), generated by tools/make_assets.py"]
    lines += ["v %.4f %.4f %.4f" % tuple(p) for p in v] + ["vn %.4f %.4f %.4f" % tuple(p) for p in n]
    lines += ["vt 0.0 0.0 0.0", "g sphere", "usemtl sphere", "s 1"]
    nv = len(v)
    lines += ["f %d/-1/%d %d/-1/%d %d/-1/%d" % (a - nv, a - nv, b_ - nv, b_ - nv, c - nv, c - nv) for a, b_, c in f]
    write_text(os.path.join(root, "scenes/cornell-glossy/ball.obj"), "\n".join(lines) + "\n")

    write_json(os.path.join(root, "scenes/cornell-glass.json"), {
        "sensor": {"lookAt": {"origin": vec(0, 1, 6.8), "target": vec(0, 1, 0), "up": vec(0, 1, 0)}, "fov": s(19.5)},
        "models": [
            {"type": "obj", "filename": "scenes/cornell-glossy/CornellBox-Glossy.obj"},
            {"type": "obj", "filename": "scenes/cornell-glossy/box.obj", "bsdf": {"type": "mirror"}},
            {"type": "obj", "filename": "scenes/cornell-glossy/ball.obj", "bsdf": {"type": "glass"},
             "transform": {"translate": vec(0, 0, 1)}},
        ],
    })


# ----------------------------------------------------------------------------------------- env maps
def synthetic_sky(width, height, seed, sun_theta, sun_phi, sun_radiance):
    """Seeded sky: horizon gradient + low-frequency clouds + a small very bright sun disc (lat-long, row 0 = +y)."""
    rng = np.random.default_rng(seed)
    theta = (np.arange(height) + 0.5) / height * math.pi
    phi = (np.arange(width) + 0.5) / width * 2 * math.pi
    th, ph = np.meshgrid(theta, phi, indexing="ij")
    up = np.cos(th)
    sky = np.stack([0.35 + 0.25 * (1 - np.abs(up)), 0.55 + 0.2 * (1 - np.abs(up)), 0.95 - 0.1 * (1 - np.abs(up))], -1)
    ground = np.array([0.18, 0.16, 0.13])
    clouds = np.zeros((height, width))
    for k in range(1, 6):
        a, b_, c = rng.uniform(0, 2 * math.pi, 3)
        clouds += np.sin(k * ph + a) * np.cos(k * 1.7 * th + b_) * np.sin(0.5 * k * ph + c) / k
    clouds = np.clip(0.5 + 0.5 * clouds, 0, 1) ** 2
    img = np.where(up[..., None] > 0, sky * (0.6 + 0.8 * clouds[..., None]), ground * (0.8 + 0.2 * clouds[..., None]))
    d = np.stack([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)], -1)
    sd = np.array([math.sin(sun_theta) * math.cos(sun_phi), math.cos(sun_theta), math.sin(sun_theta) * math.sin(sun_phi)])
    cosang = d @ sd
    img = img + (cosang > math.cos(math.radians(1.5)))[..., None] * np.array(sun_radiance)
    img = img + np.clip(cosang, 0, 1)[..., None] ** 64 * 2.0
    return img.astype(np.float32)


# ----------------------------------------------------------------------------------------- dragon stand-in
def knot_mesh(segments_u, segments_v, seed=7):
    """Closed (2,3) torus-knot tube with seeded multi-octave bumps: segments_u*segments_v*2 triangles."""
    rng = np.random.default_rng(seed)
    u = np.arange(segments_u) / segments_u * 2 * math.pi
    p, q, R, r = 2, 3, 44.0, 17.0
    curve = lambda t: np.stack([(R + r * np.cos(q * t)) * np.cos(p * t), (R + r * np.cos(q * t)) * np.sin(p * t),
                                1.35 * r * np.sin(q * t)], -1)
    c = curve(u)
    eps = 1e-4
    tangent = curve(u + eps) - curve(u - eps)
    tangent /= np.linalg.norm(tangent, axis=1, keepdims=True)
    ref = np.array([0.0, 0.0, 1.0])
    n1 = np.cross(tangent, ref); n1 /= np.linalg.norm(n1, axis=1, keepdims=True)
    n2 = np.cross(tangent, n1)
    v = np.arange(segments_v) / segments_v * 2 * math.pi
    uu, vv = np.meshgrid(u, v, indexing="ij")
    radius = 10.0 + 1.5 * np.sin(5 * uu) * np.cos(3 * vv)
    for k in range(6):
        fu, fv = int(rng.integers(8, 90)), int(rng.integers(2, 24))
        radius += rng.uniform(0.15, 0.6) * np.sin(fu * uu + rng.uniform(0, 6.28)) * np.cos(fv * vv + rng.uniform(0, 6.28))
    pts = c[:, None, :] + radius[..., None] * (np.cos(vv)[..., None] * n1[:, None, :] + np.sin(vv)[..., None] * n2[:, None, :])
    idx = np.arange(segments_u * segments_v).reshape(segments_u, segments_v)
    a = idx; b_ = np.roll(idx, -1, 0); cc = np.roll(np.roll(idx, -1, 0), -1, 1); d = np.roll(idx, -1, 1)
    # winding chosen so that (v1-v0)x(v2-v0) points out of the tube: meshes without vn get n_s = n_g from the winding (Q3),
    # and an inside-out mesh would render black (Lambertian / Microfacet return 0 when wo is below n_s, Q2)
    faces = np.concatenate([np.stack([a, cc, b_], -1).reshape(-1, 3), np.stack([a, d, cc], -1).reshape(-1, 3)])
    return pts.reshape(-1, 3) * 1.6, faces  # ~ +-115 units across, the extent of the model dragon.json's camera frames


def make_dragon(root, segments_u):
    verts, faces = knot_mesh(segments_u, 212)
    verts[:, 2] += -40.0 - verts[:, 2].min() + 0.5      # rest on the floor quad (z = -40 in world space)
    verts[:, 1] += 40.0
    # dragon.json rotates the mesh by -53 degrees about y (non-legacy: rotateY(-(-53 deg))); pre-apply the inverse
    ang = math.radians(53.0)
    ca, sa = math.cos(ang), math.sin(ang)
    rot = np.array([[ca, 0, sa], [0, 1, 0], [-sa, 0, ca]])  # = rotateY(+53 deg) of src/matrix.cpp:118-131
    local = verts @ rot                                        # rot^-1 applied to row vectors
    write_obj(os.path.join(root, "assets/dragon.obj"), local, faces,
              header="# procedural stand-in for the Stanford dragon (%d triangles), tools/make_assets.py\n" % len(faces))
    sky = synthetic_sky(2048, 1024, seed=20060807, sun_theta=math.radians(50), sun_phi=math.radians(200), sun_radiance=(900, 800, 650))
    write_exr(os.path.join(root, "assets/20060807_wells6_hd.exr"), sky)
    write_json(os.path.join(root, "scenes/dragon.json"), {
        "sensor": {"lookAt": {"origin": vec(277, -240, 250), "target": vec(0, 60, -30), "up": vec(0, 0, 1)},
                   "fov": s(33), "flipHandedness": True},
        "models": [
            {"type": "quad", "upAxis": "z",
             "transform": {"scale": vec(1000, 1000, 1), "rotate": vec(0, 0, 0), "translate": vec(0, 0, -40)},
             "bsdf": {"type": "lambertian", "diffuseReflectance": vec(0.1, 0.1, 0.1)}},
            {"type": "obj", "filename": "assets/dragon.obj",
             "transform": {"scale": vec(1, 1, 1), "rotate": vec(0, -53, 0), "translate": vec(0, 0, 0)},
             "bsdf": {"type": "plastic", "diffuseReflectance": vec(0.1, 0.1, 0.4),
                      "distribution": {"type": "beckmann", "alpha": s(0.1)}}},
        ],
        "environmentLight": {"filename": "assets/20060807_wells6_hd.exr", "scale": s(2.5),
                             "transform": {"legacy": True, "scale": vec(1, -1, 1), "rotate": vec(180, 0, -90)}},
    })


# ----------------------------------------------------------------------------------------- mis-pbrt stand-in
def make_mis(root):
    plates = {  # Veach MIS arrangement: four tilted 8-wide plates, roughest closest to the camera
        "plate1": [(4, -2.70651, 0.25609), (4, -2.08375, -0.526323), (-4, -2.08375, -0.526323), (-4, -2.70651, 0.25609)],
        "plate2": [(4, -3.28825, 1.36972), (4, -2.83856, 0.476536), (-4, -2.83856, 0.476536), (-4, -3.28825, 1.36972)],
        "plate3": [(4, -3.73096, 2.70046), (4, -3.43378, 1.74564), (-4, -3.43378, 1.74564), (-4, -3.73096, 2.70046)],
        "plate4": [(4, -3.99615, 4.0667), (4, -3.82069, 3.08221), (-4, -3.82069, 3.08221), (-4, -3.99615, 4.0667)],
    }
    for name, quad in plates.items():
        write_ply(os.path.join(root, "assets/mis-pbrt/geometry/%s.ply" % name), quad, [(0, 1, 2), (2, 3, 0)])
    floor = [(-10, -4.14615, -10), (-10, -4.14615, 10), (10, -4.14615, 10), (10, -4.14615, -10),
             (-10, -10, -2), (10, -10, -2), (10, 10, -2), (-10, 10, -2)]
    write_ply(os.path.join(root, "assets/mis-pbrt/geometry/floor.ply"), floor, [(0, 1, 2), (2, 3, 0), (4, 5, 6), (6, 7, 4)])
    spheres = [((10, 10, 4), 0.5, 800), ((-1.25, 0, 0), 0.1, 100), ((-3.75, 0, 0), 0.03333, 901.803),
               ((1.25, 0, 0), 0.3, 11.1111), ((3.75, 0, 0), 0.9, 1.23457)]
    models = [{"type": "sphere", "center": vec(*c), "radius": s(r),
               "bsdf": {"type": "lambertian", "diffuseReflectance": vec(0, 0, 0), "emit": vec(e, e, e)}} for c, r, e in spheres]
    for name, alpha in (("plate1", 0.005), ("plate2", 0.02), ("plate3", 0.05), ("plate4", 0.1)):
        models.append({"type": "ply", "filename": "assets/mis-pbrt/geometry/%s.ply" % name,
                       "bsdf": {"type": "plastic", "distribution": {"type": "beckmann", "alpha": s(alpha)},
                                "diffuseReflectance": vec(0.07, 0.09, 0.13)}})
    models.append({"type": "ply", "filename": "assets/mis-pbrt/geometry/floor.ply",
                   "bsdf": {"type": "lambertian", "diffuseReflectance": vec(0.4, 0.4, 0.4)}})
    write_json(os.path.join(root, "scenes/mis-pbrt.json"), {
        "sensor": {"lookAt": {"origin": vec(0, 2, 15), "target": vec(0, 1.69521, 14.0476), "up": vec(0, 0.952421, -0.304787)},
                   "fov": "28.0000262073138", "flipHandedness": True},
        "models": models,
    })


# ----------------------------------------------------------------------------------------- teapot stand-in
def revolve(profile, segments):
    """Surface of revolution about y of a closed profile [(radius, y)]; poles where radius == 0."""
    prof = np.array(profile)
    ang = np.arange(segments) / segments * 2 * math.pi
    verts = np.stack([prof[:, None, 0] * np.cos(ang)[None, :], np.repeat(prof[:, 1:2], segments, 1),
                      prof[:, None, 0] * np.sin(ang)[None, :]], -1).reshape(-1, 3)
    n = len(prof)
    idx = np.arange(n * segments).reshape(n, segments)
    a = idx[:-1]; b_ = idx[1:]; c = np.roll(idx[1:], -1, 1); d = np.roll(idx[:-1], -1, 1)
    faces = np.concatenate([np.stack([a, c, b_], -1).reshape(-1, 3), np.stack([a, d, c], -1).reshape(-1, 3)])
    # drop degenerate triangles at the poles
    p = verts[faces]
    area = np.linalg.norm(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), axis=1)
    return verts, faces[area > 1e-9]


def make_teapot(root):
    t = np.linspace(0, 1, 188)
    body_r = 4.2 * np.sin(np.clip(t, 0, 1) * math.pi) ** 0.55 * (1 - 0.25 * t) + 0.0
    body_y = 0.05 + 5.2 * t
    body_r[0] = 0.0; body_r[-1] = 0.0
    v0, f0 = revolve(list(zip(body_r, body_y)), 128)          # ~47.9 k triangles
    t = np.linspace(0, 1, 154)
    lid_r = np.where(t < 0.7, 2.6 * np.clip(np.cos(t / 0.7 * math.pi / 2), 0, 1) ** 0.7 + 0.35, 0.35 + 0.45 * np.sin((t - 0.7) / 0.3 * math.pi))
    lid_r = lid_r * (t < 1.0)
    lid_y = 5.6 + 1.9 * t
    lid_r = np.concatenate([[0.0], lid_r]); lid_y = np.concatenate([[5.6], lid_y]); lid_r[-1] = 0.0
    v1, f1 = revolve(list(zip(lid_r, lid_y)), 256)             # ~78.8 k triangles
    write_obj(os.path.join(root, "assets/teapot/Mesh000.obj"), v0, f0, header="# procedural teapot body, tools/make_assets.py\n")
    write_obj(os.path.join(root, "assets/teapot/Mesh001.obj"), v1, f1, header="# procedural teapot lid, tools/make_assets.py\n")
    sky = synthetic_sky(1024, 512, seed=1234, sun_theta=math.radians(40), sun_phi=math.radians(30), sun_radiance=(400, 380, 330))
    write_exr(os.path.join(root, "assets/teapot/envmap.exr"), sky)
    glass = {"type": "glass", "alpha": s(0.01), "diffuseReflectance": vec(1, 1, 1)}
    write_json(os.path.join(root, "scenes/teapot.json"), {
        "sensor": {"lookAt": {"origin": vec(23.895, 11.2207, 0.0400773), "target": vec(-0.953633, 2.17253, -0.0972613),
                              "up": vec(0, 1, 0)}, "fov": s(35)},
        "models": [
            {"type": "obj", "filename": "assets/teapot/Mesh000.obj", "bsdf": glass},
            {"type": "obj", "filename": "assets/teapot/Mesh001.obj", "bsdf": glass},
            {"type": "quad", "upAxis": "z",
             "transform": {"legacy": True, "scale": vec(113.071, 113.071, 113.071), "rotate": vec(90, 45, 0), "translate": vec(0, 0, 0)},
             "bsdf": {"type": "lambertian",
                      "albedo": {"type": "checkerboard", "onColor": vec(0.725, 0.71, 0.68), "offColor": vec(0.325, 0.31, 0.25),
                                 "resolution": {"u": s(20), "v": s(20)}},
                      "diffuseReflectance": vec(1, 1, 1)}},
        ],
        "environmentLight": {"filename": "assets/teapot/envmap.exr", "scale": s(1)},
    })


# ----------------------------------------------------------------------------------------- textured room
def procedural_texture(width, height, seed, kind):
    """8-bit RGB images with structure at several scales (so a wrong texel, flip or wrap shows up in a render)."""
    y, x = np.mgrid[0:height, 0:width].astype(np.float64)
    rng = np.random.default_rng(seed)
    if kind == "wood":
        rings = np.sin((x / width * 9 + 0.6 * np.sin(y / height * 7)) * 2 * math.pi) * 0.5 + 0.5
        grain = rng.random((height, width)) * 0.15
        rgb = np.stack([0.45 + 0.35 * rings + grain, 0.25 + 0.25 * rings + grain * 0.7, 0.10 + 0.12 * rings + grain * 0.4], -1)
    elif kind == "tiles":
        tile = ((x // (width / 8)).astype(int) + (y // (height / 6)).astype(int)) % 3
        base = np.array([[0.8, 0.75, 0.7], [0.3, 0.45, 0.6], [0.65, 0.3, 0.25]])[tile]
        grout = ((x % (width / 8)) < 2) | ((y % (height / 6)) < 2)
        rgb = np.where(grout[..., None], 0.15, base + (rng.random((height, width, 1)) - 0.5) * 0.1)
    else:  # "picture": gradients with a bright disc, asymmetric so flips are visible
        rgb = np.stack([x / width, y / height, 0.5 + 0.5 * np.sin((x + 2 * y) / 23.0)], -1)
        disc = ((x - 0.7 * width) ** 2 + (y - 0.3 * height) ** 2) < (0.12 * height) ** 2
        rgb = np.where(disc[..., None], np.array([1.0, 0.9, 0.2]), rgb)
    return (np.clip(rgb, 0, 1) * 255 + 0.5).astype(np.uint8)


def make_textured(root):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_inputs import write_png
    os.makedirs(os.path.join(root, "assets/textured"), exist_ok=True)
    write_png(os.path.join(root, "assets/textured/wood.png"), procedural_texture(256, 192, 11, "wood"))
    write_png(os.path.join(root, "assets/textured/tiles.png"), procedural_texture(200, 150, 12, "tiles"))
    write_png(os.path.join(root, "assets/textured/picture.png"), procedural_texture(97, 61, 13, "picture"))
    # a cuboid with per-face uv coordinates (v/vt/vn), textured plastic
    cx, cz, half, height, angle = 0.35, 0.1, 0.3, 0.6, math.radians(25.0)
    ca, sa = math.cos(angle), math.sin(angle)
    corner = lambda dx, dz, y: (cx + ca * dx * half - sa * dz * half, y, cz + sa * dx * half + ca * dz * half)
    b = [corner(-1, -1, 0.0), corner(1, -1, 0.0), corner(1, 1, 0.0), corner(-1, 1, 0.0)]
    t = [corner(-1, -1, height), corner(1, -1, height), corner(1, 1, height), corner(-1, 1, height)]
    quads = [([b[3], b[2], b[1], b[0]], (0, -1, 0)), ([t[0], t[1], t[2], t[3]], (0, 1, 0))]
    for i in range(4):
        j = (i + 1) % 4
        e = np.array(b[j]) - np.array(b[i])
        nrm = np.cross(e, (0, 1, 0)); nrm = nrm / np.linalg.norm(nrm)
        quads.append(([b[i], t[i], t[j], b[j]], tuple(-nrm)))
    lines = ["# textured box, generated by tools/make_assets.py"]
    vt = ["vt 2.0 -0.5 0.0", "vt 2.0 1.5 0.0", "vt 0.0 1.5 0.0", "vt 0.0 -0.5 0.0"]  # outside [0, 1]: exercises the wrap
    for quad, nrm in quads:
        lines += ["v %.4f %.4f %.4f" % p for p in quad] + ["vn %.4f %.4f %.4f" % nrm] + vt
        lines += ["f -4/-4/-1 -3/-3/-1 -2/-2/-1", "f -2/-2/-1 -1/-1/-1 -4/-4/-1"]
    write_text(os.path.join(root, "assets/textured/box.obj"), "\n".join(lines) + "\n")
    write_json(os.path.join(root, "scenes/textured.json"), {
        "sensor": {"lookAt": {"origin": vec(0, 1.1, 3.6), "target": vec(0, 0.6, 0), "up": vec(0, 1, 0)}, "fov": s(38)},
        "models": [
            {"type": "quad", "transform": {"scale": vec(1.6, 1, 1.6)},  # floor
             "bsdf": {"type": "lambertian", "texture": "assets/textured/wood.png", "diffuseReflectance": vec(1, 1, 1)}},
            {"type": "quad", "upAxis": "z", "transform": {"scale": vec(1.6, 1.0, 1), "translate": vec(0, 1.0, -1.6)},  # back wall
             "bsdf": {"type": "lambertian", "texture": "assets/textured/tiles.png", "diffuseReflectance": vec(1, 1, 1)}},
            {"type": "quad", "upAxis": "z", "transform": {"scale": vec(0.5, 0.32, 1), "translate": vec(-0.7, 1.1, -1.59)},  # framed picture
             "bsdf": {"type": "lambertian", "texture": "assets/textured/picture.png", "diffuseReflectance": vec(1, 1, 1)}},
            {"type": "obj", "filename": "assets/textured/box.obj",
             "bsdf": {"type": "plastic", "texture": "assets/textured/wood.png", "diffuseReflectance": vec(1, 1, 1),
                      "distribution": {"type": "beckmann", "alpha": s(0.15)}}},
            {"type": "sphere", "center": vec(-0.55, 0.35, 0.45), "radius": s(0.35),
             "bsdf": {"type": "lambertian", "diffuseReflectance": vec(0.7, 0.7, 0.75)}},
            {"type": "quad", "transform": {"legacy": True, "scale": vec(0.5, 1, 0.5), "rotate": vec(180, 0, 0), "translate": vec(0, 2.4, 0.3)},
             "bsdf": {"type": "lambertian", "diffuseReflectance": vec(0, 0, 0), "emit": vec(18, 16, 12)}},
        ],
    })


# ----------------------------------------------------------------------------------------- instancing
def column_major(rows):
    """4x4 given as rows -> the 16 strings of an `instanced` model's transform (RTC_FORMAT_FLOAT4X4_COLUMN_MAJOR)"""
    m = np.asarray(rows, np.float64).reshape(4, 4)
    return [s(float(np.float32(v))) for v in m.T.reshape(-1)]


def trs(translate, rotate_y_deg=0.0, rotate_x_deg=0.0, scale=(1.0, 1.0, 1.0)):
    ry, rx = math.radians(rotate_y_deg), math.radians(rotate_x_deg)
    S = np.diag([scale[0], scale[1], scale[2], 1.0])
    Ry = np.array([[math.cos(ry), 0, math.sin(ry), 0], [0, 1, 0, 0], [-math.sin(ry), 0, math.cos(ry), 0], [0, 0, 0, 1]])
    Rx = np.array([[1, 0, 0, 0], [0, math.cos(rx), -math.sin(rx), 0], [0, math.sin(rx), math.cos(rx), 0], [0, 0, 0, 1]])
    T = np.eye(4); T[:3, 3] = translate
    return T @ Ry @ Rx @ S


def make_instanced(root):
    """scenes/instanced.json: no scene of the reference uses `instance` / `instanced`, so this one exercises the parser paths
    (src/scene_parser.cpp:231-249, :449-492) and both instance levels: a smooth-shaded blob (v//vn faces) and a flat-shaded box
    (no normals: n_s = n_g) defined once, placed under rotations, non-uniform scales and a mirroring, and a `cluster` instance
    scene that itself places the blob twice (instID[0] = cluster placement, instID[1] = blob placement)."""
    d = os.path.join(root, "assets/instanced")
    os.makedirs(d, exist_ok=True)
    verts, normals, faces = uv_sphere((0.0, 0.0, 0.0), 1.0, segments=24, rings=11)
    bump = 1.0 + 0.25 * np.sin(3.0 * verts[:, 0]) * np.cos(2.0 * verts[:, 2])  # not a sphere: hits depend on the rotation
    write_obj(os.path.join(d, "blob.obj"), verts * bump[:, None] * 0.3, faces, normals, header="# instanced blob, tools/make_assets.py\n")
    lines = ["# instanced box, tools/make_assets.py"]
    for quad in closed_box((-0.25, 0.0, -0.25), (0.25, 0.5, 0.25)):
        lines += ["v %.4f %.4f %.4f" % p for p in quad] + ["f -4 -3 -2", "f -2 -1 -4"]
    write_text(os.path.join(d, "box.obj"), "\n".join(lines) + "\n")
    blob = {"type": "obj", "filename": "assets/instanced/blob.obj",
            "bsdf": {"type": "plastic", "diffuseReflectance": vec(0.6, 0.25, 0.2), "distribution": {"type": "beckmann", "alpha": s(0.2)}}}
    box = {"type": "obj", "filename": "assets/instanced/box.obj", "transform": {"translate": vec(0.0, 0.0, 0.0)},
           "bsdf": {"type": "lambertian", "diffuseReflectance": vec(0.3, 0.5, 0.7)}}
    mirror = np.diag([-1.0, 1.0, 1.0, 1.0])
    write_json(os.path.join(root, "scenes/instanced.json"), {
        "sensor": {"lookAt": {"origin": vec(0, 1.6, 4.2), "target": vec(0, 0.45, 0), "up": vec(0, 1, 0)}, "fov": s(36)},
        "models": [
            {"type": "quad", "transform": {"scale": vec(2.5, 1, 2.5)},  # geometry 0 of the root scene
             "bsdf": {"type": "lambertian", "diffuseReflectance": vec(0.7, 0.7, 0.7)}},
            {"type": "instance", "name": "blob", "models": [blob]},      # definitions take no geometry id
            {"type": "instance", "name": "box", "models": [box]},
            {"type": "instance", "name": "cluster", "models": [
                {"type": "instanced", "instance_name": "blob", "transform": column_major(trs((0.35, 0.0, 0.0), 40.0, 0.0, (0.6, 0.6, 0.6)))},
                {"type": "instanced", "instance_name": "blob", "transform": column_major(trs((-0.3, 0.15, 0.1), -70.0, 25.0, (0.5, 0.8, 0.5)))},
                {"type": "obj", "filename": "assets/instanced/box.obj", "transform": {"scale": vec(0.4, 0.4, 0.4), "translate": vec(0.0, -0.35, 0.0)},
                 "bsdf": {"type": "lambertian", "diffuseReflectance": vec(0.8, 0.7, 0.2)}},
            ]},
            {"type": "instanced", "instance_name": "blob", "transform": column_major(trs((-1.1, 0.45, 0.2), 30.0, 0.0, (1.0, 1.3, 1.0)))},   # geometry 1
            {"type": "instanced", "instance_name": "box", "transform": column_major(trs((0.0, 0.0, -0.6), 35.0, 0.0, (1.2, 1.6, 0.8)))},     # geometry 2
            {"type": "instanced", "instance_name": "box", "transform": column_major(trs((1.2, 0.0, 0.5), -20.0, 0.0) @ mirror)},              # 3: mirrored
            {"type": "instanced", "instance_name": "cluster", "transform": column_major(trs((0.15, 0.75, 0.9), 15.0, 10.0, (0.9, 0.9, 0.9)))},  # 4
            {"type": "instanced", "instance_name": "cluster", "transform": column_major(trs((-0.2, 1.35, -0.9), 200.0, -15.0, (1.1, 1.1, 1.1)))},  # 5
            {"type": "quad", "transform": {"legacy": True, "scale": vec(0.7, 1, 0.7), "rotate": vec(180, 0, 0), "translate": vec(0, 3.0, 0.5)},  # 6
             "bsdf": {"type": "lambertian", "diffuseReflectance": vec(0, 0, 0), "emit": vec(20, 19, 17)}},
        ],
    })


# ----------------------------------------------------------------------------------------- test scenes
def make_test_scenes(root):
    img = np.zeros((500, 1000, 3), np.float32)
    img[239, 753] = 10000.0
    write_exr(os.path.join(root, "test_scenes/1_pixel_test.exr"), img)
    write_json(os.path.join(root, "test_scenes/environment_map_sampling.json"), {
        "sensor": {"lookAt": {"origin": vec(0, 4, 8), "target": vec(0, 0, 0), "up": vec(0, 1, 0)}, "fov": s(40)},
        "models": [
            {"type": "quad", "transform": {"scale": vec(3, 1, 3)},
             "bsdf": {"type": "lambertian", "diffuseReflectance": vec(1, 1, 1)}},
            {"type": "sphere", "center": vec(0, 1, 0), "radius": s(0.75),
             "bsdf": {"type": "oren-nayar", "diffuseReflectance": vec(0.7, 0.6, 0.5), "sigma": s(0.3)}},
            {"type": "sphere", "center": vec(-2, 0.6, 0.5), "radius": s(0.6),
             "bsdf": {"type": "microfacet", "distribution": {"type": "ggx", "alpha": s(0.2)}}},
        ],
        "environmentLight": {"filename": "test_scenes/1_pixel_test.exr", "scale": s(1)},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--root", default=ROOT)
    ap.add_argument("--dragon-segments", type=int, default=2048, help="knot segments along the curve (x212x2 triangles)")
    ap.add_argument("--force", action="store_true")
    args = ap.parse_args()
    stamp = os.path.join(args.root, "assets/.generated-v5-%d" % args.dragon_segments)
    expected = ["scenes/cornell.json", "scenes/cornell-glass.json", "scenes/dragon.json", "scenes/mis-pbrt.json", "scenes/teapot.json",
                "assets/dragon.obj", "assets/20060807_wells6_hd.exr", "assets/teapot/envmap.exr", "test_scenes/1_pixel_test.exr",
                "scenes/textured.json", "assets/textured/wood.png", "scenes/cornell-medium.json", "test_scenes/medium_sphere.json", "scenes/instanced.json", "assets/instanced/blob.obj"]
    if os.path.exists(stamp) and all(os.path.exists(os.path.join(args.root, f)) for f in expected) and not args.force:
        print("assets up to date:", stamp)
        return 0
    make_cornell(args.root)
    make_cornell_glass(args.root)
    make_mis(args.root)
    make_teapot(args.root)
    make_test_scenes(args.root)
    make_textured(args.root)
    make_cornell_medium(args.root)
    make_instanced(args.root)
    make_dragon(args.root, args.dragon_segments)
    write_text(stamp, "ok\n")
    print("generated scenes/, assets/, test_scenes/ under", args.root)
    return 0


if __name__ == "__main__":
    sys.exit(main())
