"""Known-answer vectors the reference's own tests / fixtures hold for the hot path, as checks that run against
any library exporting the ABI (the CPU oracle in the CPU suite, the CUDA library in the GPU suite).

  test/transform_test.cpp:6-14   normalToWorldSpace(n, dir) maps (0,1,0) to n exactly          -> tangent_frame_maps_y_to_normal
  test/vector_test.cpp:6-14      Vector3::reflect                                              -> reflect_follows_the_source
  test_scenes/1_pixel_test.exr   one non-zero texel (row 239, col 753) in a 1000x500 map       -> one_pixel_environment_map
test/camera_test.cpp pins Camera::calculatePixel, the light tracer's world->pixel mapping, which is not on the path (and its
"Cornell light" case, pixel x 57 / y 92, disagrees with src/camera.cpp:57-93 at this commit, which mirrors the film: x 42, y 7);
Camera::generateRay is pinned by the compiled reference instead (tests/golden/scene_*.npz: cam_rays).
"""
import numpy as np

from golden_inputs import LAMBERTIAN, MIRROR, material_desc
from parity import make_isects


def _commit_dummy(api, materials):
    ids = [api.add_material(material_desc(m)) for m in materials]
    api.add_triangle_mesh([[0, 0, 0], [1, 0, 0], [0, 1, 0]], None, None, [[0, 1, 2]], ids[0])
    api.set_camera((0, 0, 5), (0, 0, 0), (0, 1, 0), 0.5, 8, 8)
    api.commit()
    return ids


def tangent_frame_maps_y_to_normal(api):
    """a cosine-hemisphere sample with xi1 = 0 is the local direction (0, 1, 0); in world space it must be the normal, bit for bit"""
    n = (np.array([1, 2, 3], np.float32) / np.sqrt(np.float32(14))).astype(np.float32)
    mat = _commit_dummy(api, [dict(type=LAMBERTIAN, diffuse=(1, 1, 1))])[0]
    isects = make_isects(np.array([[1, 0, 0]], np.float32), n[None], n[None], np.zeros((1, 2), np.float32), mat)
    wi, pdf, thr = api.bsdf_sample(mat, isects, np.array([[0.0, 0.37, 0.0]], np.float32))
    assert np.array_equal(wi[0], isects["shading_normal"][0]), (wi, n)


def reflect_follows_the_source(api):
    """src/vector.cpp:64-67 computes 2 (n.w) n - w.  (test/vector_test.cpp expects the opposite sign for an un-normalised input;
    that test binary cannot be built at this commit and disagrees with the source the renderer runs.)  A mirror with the
    normal along +y must send wo = (-a, b, 0) to (a, b, 0)."""
    mat = _commit_dummy(api, [dict(type=MIRROR)])[0]
    wo = np.array([[-0.6, 0.8, 0.0]], np.float32)
    up = np.array([[0, 1, 0]], np.float32)
    wi, pdf, thr = api.bsdf_sample(mat, make_isects(wo, up, up, np.zeros((1, 2), np.float32), mat), np.zeros((1, 3), np.float32))
    assert np.allclose(wi[0], [0.6, 0.8, 0.0], atol=1e-6) and pdf[0] == 1.0
    assert np.allclose(thr[0], 1.0 / 0.8, rtol=1e-6)  # Mirror::sample: throughput 1 / cos(theta), src/mirror.cpp:21-37


def one_pixel_environment_map(api_with_scene):
    """test_scenes/environment_map_sampling.json: every environment sample lands on texel (row 239, col 753) of the 1000x500 map,
    at its centre, with pdf = W H / (sin(theta) 2 pi^2) (src/environment_light.cpp:82-105) and radiance 10000"""
    api = api_with_scene
    n = 256
    from golden_inputs import uniform_floats
    xi = uniform_floats(4242, (n, 3))
    xi[:, 0] = 0.0  # the environment light is the only light
    ref = np.zeros((n, 3), np.float32); ref[:, 1] = 0.25
    ls = api.light_sample(ref, xi)
    w, h = 1000, 500
    theta = np.float32((239 + 0.5) / h) * np.float32(np.pi)
    phi = np.float32((753 + 0.5) / w) * np.float32(2 * np.pi)
    want_dir = np.array([np.sin(theta) * np.cos(phi), np.cos(theta), np.sin(theta) * np.sin(phi)], np.float64)
    got_dir = (ls["point"] - ref) / 10000.0
    assert np.abs(got_dir - want_dir).max() < 2e-4  # point = ref + 10000 * dir in fp32
    assert (ls["measure"] == 0).all()
    want_pdf = w * h / (np.sin(np.float64(theta)) * 2 * np.pi ** 2)
    assert np.allclose(ls["solid_angle_pdf"], want_pdf, rtol=2e-5)
    assert np.allclose(ls["emit"], 10000.0)
