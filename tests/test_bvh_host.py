"""CPU tests of the host BVH builder and of the traversal code it shares with the kernels (ptc_bvh_selfcheck needs no GPU):
the compressed wide BVH must return exactly what a brute-force scan with the same triangle test returns."""
import ctypes

import numpy as np
import pytest

from golden_inputs import uniform_floats, unit_vectors
from pathed_b200._binding import RAY_DTYPE, cuda_lib, rays_array


def selfcheck(positions, indices, rays, builder=0, brute_force=True):
    lib = cuda_lib()
    positions = np.ascontiguousarray(positions, np.float32); indices = np.ascontiguousarray(indices, np.uint32)
    n = len(rays)
    t_bvh = np.zeros(n, np.float32); p_bvh = np.zeros(n, np.uint32); t_bf = np.zeros(n, np.float32); p_bf = np.zeros(n, np.uint32)
    stats = (ctypes.c_uint64 * 6)()
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    cost = ctypes.c_double(0)
    rc = lib.ptc_bvh_selfcheck_builder(ctypes.c_int(builder), ptr(positions), ctypes.c_uint32(len(positions)), ptr(indices),
                                       ctypes.c_uint32(len(indices)), ptr(rays), ctypes.c_uint32(n), ptr(t_bvh), ptr(p_bvh),
                                       ptr(t_bf) if brute_force else None, ptr(p_bf) if brute_force else None, stats, ctypes.byref(cost))
    assert rc == 0
    keys = ("nodes", "triangles", "slots", "max_depth", "inner_visits", "triangle_tests")
    st = dict(zip(keys, [int(x) for x in stats]))
    st["sah_cost"] = cost.value
    return t_bvh, p_bvh, t_bf, p_bf, st


def bumpy_sphere(n_u, n_v, seed):
    u = np.arange(n_u) / n_u * 2 * np.pi
    v = (np.arange(n_v) + 0.5) / n_v * np.pi
    uu, vv = np.meshgrid(u, v, indexing="ij")
    r = 1.0 + 0.15 * np.sin(7 * uu + seed) * np.cos(5 * vv)
    pts = np.stack([r * np.sin(vv) * np.cos(uu), r * np.cos(vv), r * np.sin(vv) * np.sin(uu)], -1).reshape(-1, 3)
    idx = np.arange(n_u * n_v).reshape(n_u, n_v)
    a, b, c, d = idx[:, :-1], np.roll(idx, -1, 0)[:, :-1], np.roll(idx, -1, 0)[:, 1:], idx[:, 1:]
    faces = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)])
    return pts.astype(np.float32), faces.astype(np.uint32)


BUILDERS = [pytest.param(0, id="host-sah"), pytest.param(1, id="device-algorithm")]


@pytest.mark.parametrize("builder", BUILDERS)
def test_wide_bvh_matches_brute_force_on_a_mesh(builder):
    pts, faces = bumpy_sphere(96, 64, 3)
    n = 3000
    origins = unit_vectors(5, n) * np.float32(3.0)
    targets = unit_vectors(6, n) * (uniform_floats(7, (n, 1)) * np.float32(1.2))
    d = targets - origins
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    inside = np.zeros((200, 3), np.float32)  # rays from inside the mesh always hit
    rays = rays_array(np.concatenate([origins, inside]), np.concatenate([d, unit_vectors(8, 200)]))
    t_bvh, p_bvh, t_bf, p_bf, st = selfcheck(pts, faces, rays, builder)
    assert np.array_equal(p_bvh, p_bf) and np.array_equal(t_bvh, t_bf)  # bit-exact: same triangle arithmetic, same tie rule
    assert (p_bvh[-200:] != 0xFFFFFFFF).all() and (p_bvh != 0xFFFFFFFF).mean() > 0.5
    assert st["triangles"] == len(faces)
    # the SAH-optimal collapse should fill the 8-wide nodes well and keep traversal work far below a linear scan
    assert st["slots"] / st["nodes"] > 5.0, st
    assert st["inner_visits"] / len(rays) < 40 and st["triangle_tests"] / len(rays) < 30, st


@pytest.mark.parametrize("builder", BUILDERS)
@pytest.mark.parametrize("case", ["single", "degenerate", "coincident", "soup"])
def test_wide_bvh_edge_cases(case, builder):
    if case == "single":
        pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32); faces = np.array([[0, 1, 2]], np.uint32)
    elif case == "degenerate":  # zero-area and repeated triangles, all centroids identical
        pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0.3, 0.3, 0]], np.float32)
        faces = np.array([[0, 1, 2]] * 9 + [[3, 3, 3]] * 4, np.uint32)
    elif case == "coincident":  # two stacked copies of a grid: equal-depth ties must resolve to the larger primitive index
        g = np.stack(np.meshgrid(np.arange(6), np.arange(6), indexing="ij"), -1).reshape(-1, 2)
        pts = np.concatenate([g, np.zeros((36, 1))], 1).astype(np.float32)
        idx = np.arange(36).reshape(6, 6)
        quads = np.stack([idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]], -1).reshape(-1, 4)
        tris = np.concatenate([quads[:, [0, 1, 2]], quads[:, [0, 2, 3]]])
        faces = np.concatenate([tris, tris]).astype(np.uint32)
    else:
        rng = np.random.default_rng(11)
        pts = rng.uniform(-1, 1, (900, 3)).astype(np.float32); faces = np.arange(900, dtype=np.uint32).reshape(-1, 3)
    n = 1500
    o = np.stack([uniform_floats(21, (n,)) * 6 - 0.5, uniform_floats(22, (n,)) * 6 - 0.5, np.full(n, 4.0, np.float32)], 1).astype(np.float32)
    d = np.tile(np.array([[0, 0, -1]], np.float32), (n, 1))
    if case == "soup":
        o = unit_vectors(23, n) * np.float32(3); d = -o / np.linalg.norm(o, axis=1, keepdims=True) + 0.2 * unit_vectors(24, n)
    rays = rays_array(o, d)
    t_bvh, p_bvh, t_bf, p_bf, st = selfcheck(pts, faces, rays, builder)
    assert np.array_equal(p_bvh, p_bf) and np.array_equal(t_bvh, t_bf)
    assert st["triangles"] == len(faces)
    if case == "coincident":
        hit = p_bvh != 0xFFFFFFFF
        assert hit.any() and (p_bvh[hit] >= 50).all()


def test_device_builder_quality_is_close_to_the_host_sah_builder():
    """PLOC clustering (the device builder's algorithm, run here through its host emulation) against the binned-SAH host
    builder on the same mesh and rays: SAH cost and counted traversal work within 15 %."""
    pts, faces = bumpy_sphere(160, 96, 5)
    n = 4000
    origins = unit_vectors(31, n) * np.float32(3.0)
    d = unit_vectors(32, n) * np.float32(0.9) - origins
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = rays_array(origins, d)
    host = selfcheck(pts, faces, rays, 0)[4]
    dev = selfcheck(pts, faces, rays, 1)[4]
    print("host", host, "device", dev)
    assert dev["sah_cost"] < 1.15 * host["sah_cost"], (host, dev)
    work = lambda st: 80 * st["inner_visits"] + 48 * st["triangle_tests"]
    assert work(dev) < 1.15 * work(host), (host, dev)
    assert dev["slots"] / dev["nodes"] > 5.0


@pytest.mark.parametrize("top", [None, 4, 64, 100000])
def test_top_level_sah_pass_builds_the_recorded_trees(top, monkeypatch):
    """The SAH pass over the clusters PLOC leaves (bvh_build_gpu.cu: TopAxisOp / TopSplitOp / TopEmitOp / TopFinishOp, level by level)
    builds the tree the one-thread depth-first pass it replaced built: node, slot and depth counts of the wide BVH recorded with that
    pass, for the default top size, a tiny one, one between, and one that hands the whole mesh to the SAH pass (no PLOC at all); a mesh with
    duplicated triangles and a floor 50x its size exercises the median fallback for identical centroids."""
    if top is None:
        monkeypatch.delenv("PTC_PLOC_TOP", raising=False)
    else:
        monkeypatch.setenv("PTC_PLOC_TOP", str(top))
    pts, faces = bumpy_sphere(60, 40, 3)
    faces = np.concatenate([faces, faces[:500]])
    floor = np.array([[-50, -2, -50], [50, -2, -50], [50, -2, 50], [-50, -2, 50]], np.float32)
    n = len(pts)
    pts = np.concatenate([pts, floor])
    faces = np.concatenate([faces, np.array([[n, n + 1, n + 2], [n, n + 2, n + 3]], np.uint32)]).astype(np.uint32)
    rays = rays_array(unit_vectors(5, 400) * np.float32(3.0), -unit_vectors(5, 400))
    t_bvh, p_bvh, t_bf, p_bf, st = selfcheck(pts, faces, rays, 1)
    # same depths; the same triangle or its duplicate (which of two coincident copies survives the inclusive depth test depends on the
    # order they are met in: |den| * (T / |den|) may round below T)
    assert np.array_equal(t_bvh, t_bf)
    hit = p_bf != 0xFFFFFFFF
    assert np.array_equal(p_bvh != 0xFFFFFFFF, hit) and np.array_equal(faces[p_bvh[hit]], faces[p_bf[hit]]) and hit.mean() > 0.9
    recorded = {None: (506, 3811, 6), 4: (496, 3779, 6), 64: (508, 3825, 6), 100000: (496, 3775, 5)}[top]
    assert (st["nodes"], st["slots"], st["max_depth"]) == recorded, st
