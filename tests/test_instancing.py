"""hierarchical instancing through the C ABI (ptc_begin_instance / ptc_end_instance / ptc_add_instance /
ptc_intersect_instanced).  The oracle restates Embree's two-level ray transform (kernels/geometry/instance_intersector.cpp:52-109);
the CUDA library flattens every placement into its one wide BVH at ptc_commit.  The scene-file route (`instance` / `instanced`
models, src/scene_parser.cpp:231-249, :449-492) is pinned against the compiled reference by the `instanced` fixtures of
tests/test_oracle_vs_reference.py and tests/test_gpu_parity.py; this file covers the ABI itself."""
import numpy as np
import pytest

from golden_inputs import material_desc, uniform_floats, unit_vectors
from oracle_binding import oracle_context
from pathed_b200 import PathedError
from pathed_b200._binding import rays_array

INVALID = 0xFFFFFFFF


def _translate_scale(t, s):
    m = np.eye(4, dtype=np.float32)
    m[0, 0] = m[1, 1] = m[2, 2] = s
    m[:3, 3] = t
    return m


def _rotation_y(deg):
    a = np.radians(deg)
    m = np.eye(4, dtype=np.float32)
    m[0, 0] = np.cos(a); m[0, 2] = np.sin(a); m[2, 0] = -np.sin(a); m[2, 2] = np.cos(a)
    return m


def _two_level_scene(api):
    """root: floor triangle (geom 0), placement of `pair` (geom 1), placement of `tri` (geom 2), area light (geom 3);
    `tri`: one triangle with vertex normals; `pair`: two placements of `tri` (geoms 0, 1) and a mesh of its own (geom 2)"""
    grey = api.add_material(material_desc(dict(type=0, diffuse=(0.6, 0.6, 0.6))))
    red = api.add_material(material_desc(dict(type=5, diffuse=(0.7, 0.2, 0.2), distribution=0, alpha=0.2)))
    light = api.add_material(material_desc(dict(type=0, diffuse=(0, 0, 0), emit=(9, 9, 9))))
    assert api.add_triangle_mesh([[-4, 0, -4], [4, 0, -4], [0, 0, 6]], None, None, [[0, 2, 1]], grey) == 0
    tri = api.begin_instance()
    n = np.array([[0.1, 0.2, 1.0], [-0.2, 0.1, 1.0], [0.0, -0.1, 1.0]], np.float32)
    assert api.add_triangle_mesh([[-0.5, 0, 0], [0.5, 0, 0], [0, 1, 0]], n, [[0, 0], [1, 0], [0.5, 1]], [[0, 1, 2]], red) == 0  # ids count per scene
    api.end_instance()
    pair = api.begin_instance()
    assert api.add_instance(tri, _translate_scale((-0.7, 0, 0), 0.8) @ _rotation_y(30)) == 0
    assert api.add_instance(tri, _translate_scale((0.7, 0.2, 0), 1.2) @ _rotation_y(-40)) == 1
    assert api.add_triangle_mesh([[-0.3, 1.2, 0], [0.3, 1.2, 0], [0, 1.6, 0.2]], None, None, [[0, 1, 2]], grey) == 2
    api.end_instance()
    assert api.add_instance(pair, _translate_scale((0, 0.1, -1.0), 1.0) @ _rotation_y(10)) == 1
    assert api.add_instance(tri, _translate_scale((0, 0.2, 1.0), 1.5)) == 2
    assert api.add_triangle_mesh([[-1, 4, -1], [1, 4, -1], [0, 4, 1]], None, None, [[0, 1, 2]], light) == 3
    api.set_camera((0, 1.2, 5), (0, 0.7, 0), (0, 1, 0), 0.7, 64, 48)
    api.commit()
    return api


def _rays(n, seed):
    o = np.array([0, 1.2, 5], np.float32) + (uniform_floats(seed, (n, 3)) - 0.5).astype(np.float32) * np.float32(0.5)
    target = (uniform_floats(seed + 1, (n, 3)) - 0.5).astype(np.float32) * np.array([3, 2.5, 3], np.float32) + np.array([0, 0.8, 0], np.float32)
    d = target - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return rays_array(o, d.astype(np.float32))


def _check_known_answer(api):
    """one triangle facing +z, placed at z = -5 with a uniform scale of 2: the ray down -z from the origin hits it at t = 5 (world
    units), on geometry 0 / primitive 0 of the instance scene, instID[0] = the placement's geometry id, and Ng stays LOCAL
    (unscaled: (0, 0, 1) x |e2 x e1| = 1), as Embree returns it and src/scene.cpp:181-189 uses it"""
    m = api.add_material(material_desc(dict(type=0, diffuse=(0.5, 0.5, 0.5))))
    api.add_triangle_mesh([[-9, -9, -9], [-8, -9, -9], [-9, -8, -9]], None, None, [[0, 1, 2]], m)  # root geometry 0, out of the way
    scene = api.begin_instance()
    api.add_triangle_mesh([[-0.5, -0.5, 0], [0.5, -0.5, 0], [0, 0.5, 0]], None, None, [[0, 1, 2]], m)
    api.end_instance()
    placement = api.add_instance(scene, _translate_scale((0, 0, -5), 2.0))
    assert placement == 1
    api.set_camera((0, 0, 5), (0, 0, 0), (0, 1, 0), 0.5, 8, 8)
    api.commit()
    hits, inst = api.intersect_instanced(rays_array([[0, 0, 0], [3, 0, 0]], [[0, 0, -1], [0, 0, -1]]))
    assert abs(float(hits["t"][0]) - 5.0) < 1e-5 and hits["geom_id"][0] == 0 and hits["prim_id"][0] == 0
    assert inst[0].tolist() == [placement, INVALID]
    assert np.allclose(hits["ng"][0], [0, 0, 1.0], atol=1e-6)
    assert hits["geom_id"][1] == INVALID and inst[1].tolist() == [INVALID, INVALID]
    full = api.intersect_full(rays_array([[0, 0, 0]], [[0, 0, -1]]))
    assert np.allclose(full["point"][0], [0, 0, -5], atol=1e-5) and np.allclose(full["normal"][0], [0, 0, 1], atol=1e-6)


def _check_error_paths(api):
    m = api.add_material(material_desc(dict(type=0, diffuse=(0.5, 0.5, 0.5))))
    with pytest.raises(PathedError):
        api.end_instance()                                   # nothing open
    with pytest.raises(PathedError):
        api.add_instance(7, np.eye(4))                       # unknown instance scene
    scene = api.begin_instance()
    with pytest.raises(PathedError):
        api.add_sphere((0, 0, 0), 1.0, m)                    # the reference attaches spheres to the global scene (src/sphere.cpp:46)
    with pytest.raises(PathedError):
        api.add_instance(scene, np.eye(4))                   # a scene cannot contain itself
    api.set_camera((0, 0, 5), (0, 0, 0), (0, 1, 0), 0.5, 8, 8)
    with pytest.raises(PathedError):
        api.commit()                                         # definition still open


def test_oracle_instancing_known_answer_and_errors():
    _check_known_answer(oracle_context())
    _check_error_paths(oracle_context())


def test_oracle_two_level_scene_brute_force_equals_bvh():
    o = _two_level_scene(oracle_context())
    rays = _rays(20000, 5)
    o.set_option("brute_force", 1)
    a, ia = o.intersect_instanced(rays)
    o.set_option("brute_force", 0)
    b, ib = o.intersect_instanced(rays)
    assert np.array_equal(a["prim_id"], b["prim_id"]) and np.array_equal(ia, ib) and np.array_equal(a["t"], b["t"])
    hit = a["geom_id"] != INVALID
    assert 0.3 < hit.mean() < 1.0 and (ia[:, 1] != INVALID).sum() > 200 and ((ia[:, 0] == 2) & (ia[:, 1] == INVALID)).sum() > 200
    assert o.num_lights() == 1


@pytest.mark.gpu
def test_cuda_instancing_known_answer_and_errors():
    from pathed_b200 import create_context
    _check_known_answer(create_context(0))
    _check_error_paths(create_context(0))


@pytest.mark.gpu
@pytest.mark.parametrize("builder", [1, 0])
def test_cuda_flattened_instances_match_the_two_level_oracle(builder):
    """hit / miss, (instID[0], instID[1], geomID, primID) and t of the flattened BVH (device and host builder) against the oracle's
    two-level traversal; whole renders agree per pixel on identical Philox streams"""
    from pathed_b200 import create_context
    ctx = create_context(0)
    ctx.set_option("bvh_builder", builder)
    _two_level_scene(ctx)
    o = _two_level_scene(oracle_context())
    rays = _rays(1 << 17, 9)
    got, gi = ctx.intersect_instanced(rays)
    want, wi = o.intersect_instanced(rays)
    agree = (got["geom_id"] != INVALID) == (want["geom_id"] != INVALID)
    hit = agree & (want["geom_id"] != INVALID)
    same = hit & (got["geom_id"] == want["geom_id"]) & (got["prim_id"] == want["prim_id"]) & (gi == wi).all(1)
    # both sides move the ray into the placement's space with Embree's arithmetic before the triangle test: t, u, v are the same floats
    t_ok = np.abs(got["t"].astype(np.float64) - want["t"]) <= 1e-5 * np.abs(want["t"])
    exact = (got["t"] == want["t"]) & (got["u"] == want["u"]) & (got["v"] == want["v"])
    print("builder", builder, "hit/miss", agree.mean(), "ids", same[hit].mean(), "t", t_ok[hit].mean(), "bit-exact t,u,v", exact[same].mean())
    assert agree.mean() >= 0.9999 and same[hit].mean() >= 0.9999 and t_ok[hit].mean() >= 0.9999 and exact[same].mean() >= 0.999
    assert ctx.num_lights() == o.num_lights() == 1
    img = ctx.render(11, 0, 4, 0, 6)
    ref = o.render(11, 0, 4, 0, 6)
    err = np.abs(img - ref) / (np.abs(ref) + 1e-3 * max(ref.mean(), 1e-3))
    assert (err.max(-1) < 1e-3).mean() >= 0.99 and abs(img.mean() - ref.mean()) <= 0.02 * ref.mean()
