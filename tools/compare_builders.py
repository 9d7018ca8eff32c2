"""Host-side comparison of the two BVH builders on the dragon stand-in mesh (no GPU needed): SAH cost of the wide BVH and
counted traversal work of the scalar traversal on primary-like and incoherent rays (ptc_bvh_selfcheck_builder)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_bvh_host import selfcheck  # noqa: E402
from pathed_b200._binding import rays_array  # noqa: E402


def load_obj(path):
    vs, fs = [], []
    with open(path) as f:
        for line in f:
            if line.startswith("v "):
                vs.append(line.split()[1:4])
            elif line.startswith("f "):
                fs.append([t.split("/")[0] for t in line.split()[1:4]])
    return np.array(vs, np.float32), np.array(fs, np.int64).astype(np.uint32) - 1


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    pts, faces = load_obj(os.path.join(ROOT, "assets", "dragon.obj"))
    print("mesh", pts.shape, faces.shape)
    rng = np.random.default_rng(3)
    lo, hi = pts.min(0), pts.max(0)
    centre, radius = (lo + hi) / 2, np.linalg.norm(hi - lo) / 2
    # primary-like: from a point outside towards the mesh; incoherent: from surface points into random directions
    eye = centre + np.array([1.5, -1.2, 1.0], np.float32) * radius
    targets = centre + rng.uniform(-0.5, 0.5, (n, 3)).astype(np.float32) * (hi - lo)
    d0 = targets - eye; d0 /= np.linalg.norm(d0, axis=1, keepdims=True)
    tri = faces[rng.integers(0, len(faces), n)]
    surf = pts[tri].mean(1)
    d1 = rng.normal(size=(n, 3)).astype(np.float32); d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    for name, rays in (("primary", rays_array(np.tile(eye, (n, 1)).astype(np.float32), d0.astype(np.float32))),
                       ("incoherent", rays_array((surf + 1e-2 * d1).astype(np.float32), d1))):
        for builder in [int(b) for b in os.environ.get('BUILDERS', '0,1').split(',')]:
            t0 = time.time()
            brute = os.environ.get('BRUTE', '0') == '1'
            t_bvh, p_bvh, t_bf, p_bf, st = selfcheck(pts, faces, rays, builder, brute)
            ok = (np.array_equal(p_bvh, p_bf) and np.array_equal(t_bvh, t_bf)) if brute else None
            print("%-10s builder %d: exact %s nodes %d slots/node %.2f depth %d sah %.2f inner/ray %.2f tris/ray %.2f bytes/ray %.0f (%.1f s)" % (
                name, builder, ok, st["nodes"], st["slots"] / st["nodes"], st["max_depth"], st["sah_cost"], st["inner_visits"] / n,
                st["triangle_tests"] / n, (80 * st["inner_visits"] + 48 * st["triangle_tests"]) / n, time.time() - t0))


if __name__ == "__main__":
    main()
