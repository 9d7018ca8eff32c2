"""GPU tests (pytest -m gpu): the CUDA path, called through the C ABI, against (a) golden fixtures produced by the
unmodified reference and (b) the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from golden_inputs import BSDF_CONFIGS, SCENES, bsdf_inputs, light_inputs, material_desc, ray_inputs, uniform_floats
from oracle_binding import oracle_scene
from parity import REL, check_container_known_answer, check_volumetric_queries, frac_within, golden, make_isects, rel_err, rel_mse, to_rays

pytestmark = pytest.mark.gpu


def gpu_context():
    from pathed_b200 import create_context
    return create_context(0)


def gpu_scene(name, width=None, height=None):
    from pathed_b200 import load_scene
    cfg = SCENES[name]
    return load_scene(cfg["scene"], width or cfg["width"], height or cfg["height"], integrator=cfg.get("integrator", 0))


def _tiny_scene(ctx, materials):
    ids = [ctx.add_material(material_desc(m, ctx)) for m in materials]
    ctx.add_triangle_mesh([[0, 0, 0], [1, 0, 0], [0, 1, 0]], None, None, [[0, 1, 2]], ids[0])
    ctx.set_camera((0, 0, 5), (0, 0, 0), (0, 1, 0), 0.5, 8, 8)
    ctx.commit()
    return ids


@pytest.mark.parametrize("name", sorted(BSDF_CONFIGS))
def test_bsdf_matches_reference(name):
    """north_star: every BSDF eval/pdf/sample on fixed inputs within 1e-5 relative of the reference C++"""
    g = golden("bsdf_" + name)
    wo, ng, ns, uv, wi, xi = bsdf_inputs(name, len(g["pdf"]))
    ctx = gpu_context()
    mat = _tiny_scene(ctx, [BSDF_CONFIGS[name]])[0]
    isects = make_isects(wo, ng, ns, uv, mat)
    f, pdf = ctx.bsdf_eval(mat, isects, wi)
    ok_f, e_f = frac_within(f, g["f"])
    ok_p, e_p = frac_within(pdf, g["pdf"])
    # eval and pdf: every tuple of every configuration within 1e-5 (measured on 2^16 tuples: <= 1.8e-6, tests/golden/bsdf_error_table.json)
    assert ok_f == 1.0 and ok_p == 1.0, (ok_f, e_f.max(), ok_p, e_p.max())
    swi, spdf, sthr = ctx.bsdf_sample(mat, isects, xi)
    ok_wi, e_wi = frac_within(swi, g["sample_wi"])
    ok_pdf, e_pdf = frac_within(spdf, g["sample_pdf"])
    ok_thr, e_thr = frac_within(sthr, g["sample_throughput"])
    if BSDF_CONFIGS[name]["type"] in (4, 5):
        # what sample() of the microfacet family returns is ill-conditioned in the sampled direction, which goes through the host's
        # libm (see tests/test_large_batches.py): at most one of the 512 tuples may leave 1e-5, and never by more than 5e-4
        assert min(ok_wi, ok_pdf, ok_thr) >= 1.0 - 1.5 / len(spdf) and max(e_wi.max(), e_pdf.max(), e_thr.max()) <= 5e-4, (ok_wi, ok_pdf, ok_thr)
    else:
        assert ok_wi == 1.0 and ok_pdf == 1.0 and ok_thr == 1.0, (ok_wi, e_wi.max(), ok_pdf, e_pdf.max(), ok_thr, e_thr.max())


def test_shape_lights_match_reference():
    g = golden("lights_shapes")
    n = len(g["tri_pdf"])
    tri, sph, ref, xi2 = light_inputs(n)
    xi3 = np.concatenate([np.zeros((n, 1), np.float32), xi2], 1)
    emissive = dict(type=0, diffuse=(0, 0, 0), emit=(1, 2, 3))
    ctx = gpu_context()
    m = ctx.add_material(material_desc(emissive))
    ctx.add_triangle_mesh(tri.reshape(3, 3), None, None, [[0, 1, 2]], m)
    ctx.set_camera((0, 0, 5), (0, 0, 0), (0, 1, 0), 0.5, 8, 8)
    ctx.commit()
    ls = ctx.light_sample(ref, xi3)
    assert frac_within(ls["point"], g["tri_point"])[0] == 1.0
    assert frac_within(ls["inv_pdf"], g["tri_inv_pdf"])[0] == 1.0
    ctx2 = gpu_context()
    m = ctx2.add_material(material_desc(emissive))
    ctx2.add_sphere(sph[:3], sph[3], m)
    ctx2.set_camera((0, 0, 5), (0, 0, 0), (0, 1, 0), 0.5, 8, 8)
    ctx2.commit()
    ls = ctx2.light_sample(ref, xi3)
    assert (ls["measure"] == g["sph_measure"]).all()
    assert frac_within(ls["point"], g["sph_point"], tol=5e-5)[0] >= 0.99
    assert frac_within(ls["inv_pdf"], g["sph_inv_pdf"])[0] == 1.0


@pytest.mark.parametrize("name", sorted(SCENES))
def test_intersection_matches_embree(name):
    """north_star: hit/miss and primitive id agree with Embree on >= 99.99 % of a fixed ray batch, t within 1e-5 relative"""
    cfg = SCENES[name]
    g = golden("scene_" + name)
    ctx = gpu_scene(name)
    assert ctx.num_lights() == int(g["num_lights"])
    cam = ctx.camera_rays(ray_inputs(name, cfg["n_rays"]))
    assert frac_within(cam["direction"], g["cam_rays"][:, 3:])[0] == 1.0
    for prefix in ("cam_", "sec_"):
        rays = to_rays(g[prefix + "rays"])
        hits = ctx.intersect(rays)
        ref_hit = g[prefix + "geom"] != 0xFFFFFFFF
        got_hit = hits["geom_id"] != 0xFFFFFFFF
        agree = ref_hit == got_hit
        same_prim = agree & (~ref_hit | ((hits["geom_id"] == g[prefix + "geom"]) & (hits["prim_id"] == g[prefix + "prim"])))
        t_ok = rel_err(hits["t"], g[prefix + "t"]) <= REL
        tie = agree & ref_hit & ~same_prim & t_ok  # another primitive at the same depth (shared edge / coincident face)
        print(name, prefix, "hit/miss", agree.mean(), "prim", same_prim.mean(), "ties", tie.mean())
        assert agree.mean() >= 0.9999
        assert (same_prim | tie).mean() >= 0.9999
        assert t_ok[agree & ref_hit].mean() >= 0.9999
        if prefix + "inst" in g:  # RTCHit::instID of both instance levels, exact wherever the same primitive was hit
            _, inst = ctx.intersect_instanced(rays)
            exact = agree & ref_hit & same_prim
            assert np.array_equal(inst[exact], g[prefix + "inst"][exact])
            assert (g[prefix + "inst"][:, 0] != 0xFFFFFFFF).sum() > 500 and (g[prefix + "inst"][:, 1] != 0xFFFFFFFF).sum() > 50  # both levels are exercised
        ok = agree & ref_hit & same_prim
        assert frac_within(hits["u"][ok], g[prefix + "bary"][ok, 0], floor=1e-3, tol=1e-4)[0] >= 0.999
        # sphere Ng = td*D - perp cancels, and Embree's rd2 is an rcp + Newton step, so allow a few ulp more there
        assert frac_within(hits["ng"][ok], g[prefix + "ng"][ok], tol=1e-4)[0] >= 0.999
        full = ctx.intersect_full(rays)
        assert frac_within(full["point"][ok], g[prefix + "point"][ok], tol=REL)[0] >= 0.9999
        assert frac_within(full["shading_normal"][ok], g[prefix + "shading_normal"][ok], tol=2e-5)[0] >= 0.999
    occ = ctx.occluded(to_rays(g["shadow_rays"]), g["shadow_max_t"])
    # this 2-8 k-ray fixture joins pairs of surface points, many of them in one plane with a box resting on it (knife-edge
    # hits along the box's base, where Embree's exact node test culls what the inclusive triangle test would accept): at most a few
    # rays differ.  The north-star gate (>= 99.99 % on 2^20 NEE shadow rays per scene) is tests/test_large_batches.py.
    assert (occ == g["shadow_occluded"]).mean() >= 0.999
    if "ls_ref" in g:
        m = len(g["ls_ref"])
        ls = ctx.light_sample(g["ls_ref"], uniform_floats(cfg["seed"] + 29, (m, 3)))
        assert (ls["measure"] == g["ls_measure"]).all()
        assert frac_within(ls["point"], g["ls_point"], tol=5e-5)[0] >= 0.99
        assert frac_within(ls["inv_pdf"], g["ls_inv_pdf"], tol=5e-5)[0] >= 0.99
        assert frac_within(ls["emit"], g["ls_emit"])[0] == 1.0
        lp = ctx.light_pdf(to_rays(g["sec_rays"]))
        same_kind = (np.sign(lp + 1.5) == np.sign(g["sec_light_pdf"] + 1.5)) & ((lp == -1) == (g["sec_light_pdf"] == -1))
        assert same_kind.mean() >= 0.999
        assert frac_within(lp[same_kind], g["sec_light_pdf"][same_kind], tol=5e-5)[0] >= 0.99
    env = ctx.environment_radiance(g["sec_rays"][:, 3:])
    assert frac_within(env, g["sec_env_radiance"])[0] >= 0.999


@pytest.mark.parametrize("name", sorted(SCENES))
def test_paths_match_reference_with_replayed_stream(name):
    """PathTracer::L with the reference's random numbers in the reference's order"""
    cfg = SCENES[name]
    g = golden("scene_" + name)
    ctx = gpu_scene(name)
    m = cfg["n_paths"]
    xi = uniform_floats(cfg["seed"] + 101, (m, 96))
    rgb = ctx.radiance_replay(to_rays(g["cam_rays"][:m]), xi, 0, cfg["last_bounce"])
    ok, e = frac_within(rgb, g["path_rgb"], tol=1e-4, floor=1e-4)
    print(name, "paths within 1e-4:", ok)
    assert ok >= 0.998, (ok, np.sort(e)[-10:])  # measured: 0.9990 - 1.0 (a path whose branch flips on the last bit of a libm result)
    assert abs(rgb.mean() - g["path_rgb"].mean()) <= 0.02 * abs(g["path_rgb"].mean()) + 1e-6


@pytest.mark.parametrize("name", sorted(n for n in SCENES if "integrator" in SCENES[n]))
def test_volumetric_queries_match_reference(name):
    """Scene::testVolumetricOcclusion / testVolumetricIntersect with their volume events against the reference"""
    check_volumetric_queries(gpu_scene(name), golden("scene_" + name))


@pytest.mark.parametrize("name", ["cornell", "cornell_glass", "mis", "env_sampling", "teapot", "textured", "instanced", "cornell_medium",
                                  "medium_sphere", "cornell_medium_pt"])
def test_wavefront_render_matches_oracle_per_pixel(name):
    """same Philox streams (pixel, sample, bounce) on both sides: the wavefront stages must reproduce the oracle's
    per-pixel sums, up to the rare path whose branch flips on a last-bit difference in libm"""
    cfg = SCENES[name]
    w, h = cfg["width"] // 2, cfg["height"] // 2
    spp = 4
    ctx = gpu_scene(name, w, h)
    img = ctx.render(1234, 0, spp, 0, cfg["last_bounce"])
    ref = oracle_scene(cfg["scene"], w, h, integrator=cfg.get("integrator", 0)).render(1234, 0, spp, 0, cfg["last_bounce"])
    assert np.isfinite(img).all()
    err = np.abs(img - ref) / (np.abs(ref) + 1e-3 * max(ref.mean(), 1e-3))
    frac = float((err.max(-1) < 1e-3).mean())
    print(name, "pixels within 1e-3:", frac, "mean", img.mean(), ref.mean())
    assert frac >= 0.995  # measured: 1.0 on every scene
    assert abs(img.mean() - ref.mean()) <= 0.02 * ref.mean() + 1e-7


@pytest.mark.parametrize("name", ["cornell_medium", "medium_sphere", "cornell_glass", "mis"])
def test_volume_wavefront_equals_the_one_thread_per_path_kernel(name):
    """VolumePathTracer: the wavefront stages (one merged probe / continuation traversal, transmittance computed by the shadow
    warps) against the kernel that follows whole paths with the oracle-pinned building blocks -- same samples, same control flow;
    only the place of the transmittance factor in the light-sampling products differs (<= 2 ulp per term).  Scenes without media
    take the plain extend kernel over the queue (no container surface: the two acceptance rules coincide)."""
    from pathed_b200 import load_scene
    from pathed_b200._binding import VOLUME_PATH_TRACER
    cfg = SCENES[name]
    w, h = cfg["width"] // 2, cfg["height"] // 2
    ctx = load_scene(cfg["scene"], w, h, integrator=VOLUME_PATH_TRACER)
    for (spp, start, last) in [(8, 0, cfg["last_bounce"]), (3, 0, 0), (4, 1, 1), (4, 2, 3), (2, 0, 1), (2, 0, -1)]:
        ctx.set_option("volume_megakernel", 0)
        before = ctx.stats()
        a = ctx.render(77, 5, spp, start, last)
        mid = ctx.stats()
        ctx.set_option("volume_megakernel", 1)
        b = ctx.render(77, 5, spp, start, last)
        after = ctx.stats()
        assert np.isfinite(a).all()
        err = np.abs(a - b) / (np.abs(b) + 1e-6 * max(float(b.mean()), 1e-6))
        print(name, spp, start, last, "max rel", float(err.max()), "rays", mid.closest_rays - before.closest_rays, after.closest_rays - mid.closest_rays)
        assert err.max() <= 2e-5, (name, spp, start, last, float(err.max()))
        # shadow rays: the wavefront skips those whose contribution is exactly black; closest-hit rays: never more than the path kernel
        assert mid.closest_rays - before.closest_rays <= after.closest_rays - mid.closest_rays
    ctx.set_option("volume_megakernel", 0)
    one = ctx.render(3, 0, 6, 0, cfg["last_bounce"])
    ctx.set_option("paths_per_wave", w * h * 2)  # three waves instead of one: same image, bit for bit
    assert np.array_equal(one, ctx.render(3, 0, 6, 0, cfg["last_bounce"]))


def test_render_is_deterministic_and_splits_over_samples():
    """bit-exact: same seed twice; and 8 spp in one call == samples 0-3 then 4-7 into the same buffer (what the
    spp split across GPUs relies on)"""
    ctx = gpu_scene("cornell_glass", 48, 48)
    a = ctx.render(99, 0, 8, 0, 10)
    b = ctx.render(99, 0, 8, 0, 10)
    assert np.array_equal(a, b)
    c = ctx.render(99, 0, 4, 0, 10)
    c = ctx.render(99, 4, 4, 0, 10, accum=c)
    assert np.array_equal(a, c)
    ctx.set_option("paths_per_wave", 48 * 48 * 2)
    d = ctx.render(99, 0, 8, 0, 10)
    assert np.array_equal(a, d)


@pytest.mark.parametrize("scene,size,last", [("cornell_glass", 48, 10), ("mis", 40, 6), ("cornell_medium_pt", 40, 6), ("instanced", 40, 5),
                                             ("cornell_medium", 40, 6), ("medium_sphere", 40, 6)])
def test_interleaved_lanes_render_the_same_image(scene, size, last):
    """ptc_set_option("lanes", n): the wave's samples are traced as n part-waves on n streams and added in sample order -- the image is
    the one-lane image bit for bit, for sample counts the lanes divide and for ragged ones, over several waves (PathTracer and the
    VolumePathTracer's wavefront stages; 0 = the library's own choice, the default)"""
    ctx = gpu_scene(scene, size, size)
    ctx.set_option("lanes", 1)
    one = {spp: ctx.render(21, 3, spp, 0, last) for spp in (1, 5, 8)}
    for lanes in (0, 2, 3, 4, 8):
        ctx.set_option("lanes", lanes)
        for spp, image in one.items():
            assert np.array_equal(ctx.render(21, 3, spp, 0, last), image, equal_nan=True), (lanes, spp)
    ctx.set_option("paths_per_wave", size * size * 3)  # 8 spp = three waves of 3, 3, 2 samples, each split over the lanes
    assert np.array_equal(ctx.render(21, 3, 8, 0, last), one[8], equal_nan=True)


def test_render_into_page_locked_memory_equals_the_staged_copy():
    """ptc_render copies from and to a page-locked radianceLookup directly, a pageable one through its staging buffer: same sums,
    accumulated on top of what the buffer held (src/sample_integrator.cpp:61-63), also over several waves and for zero samples"""
    import torch
    ctx = gpu_scene("cornell", 64, 64)
    pageable = ctx.render(7, 0, 6, 0, 5, accum=np.full((64, 64, 3), 0.5, np.float32))
    pinned = torch.full((64, 64, 3), 0.5, dtype=torch.float32).pin_memory().numpy()
    assert ctx.render(7, 0, 6, 0, 5, accum=pinned) is pinned
    assert np.array_equal(pinned, pageable)
    ctx.set_option("paths_per_wave", 64 * 64 * 2)  # three waves: the upload is waited for by the first accumulation only
    again = torch.full((64, 64, 3), 0.5, dtype=torch.float32).pin_memory().numpy()
    ctx.render(7, 0, 6, 0, 5, accum=again)
    assert np.array_equal(again, pageable)
    ctx.render(7, 0, 0, 0, 5, accum=again)         # no wave at all
    assert np.array_equal(again, pageable)


def test_path_state_reserved_before_the_scene_exists():
    """ptc_reserve_paths: the per-path arrays allocated on a bare context (what the CLI does while it parses the scene) serve the
    renders that follow, smaller and larger than the reservation; the images do not depend on it"""
    images = []
    for reserve in (0, 8 * 8 * 2, 1 << 20):
        ctx = gpu_context()
        if reserve:
            ctx.reserve_paths(reserve)
        _tiny_scene(ctx, [dict(type=0, diffuse=(0.7, 0.6, 0.5), emit=(3, 2, 1))])
        images.append((ctx.render(11, 0, 2, 0, 4), ctx.render(11, 2, 6, 0, 4)))
        ctx.close()
    for a, b in images[1:]:
        assert np.array_equal(a, images[0][0]) and np.array_equal(b, images[0][1])
    import ctypes
    from pathed_b200 import cuda_lib
    assert cuda_lib().ptc_reserve_paths(None, ctypes.c_uint64(1)) != 0  # no context: refused with a status


def test_bounce_window_matches_oracle():
    cfg = SCENES["cornell"]
    ctx = gpu_scene("cornell", 32, 32)
    orc = oracle_scene(cfg["scene"], 32, 32)
    for start, last in [(0, 0), (0, 1), (1, 1), (2, 3), (0, 2)]:
        img = ctx.render(5, 0, 4, start, last)
        ref = orc.render(5, 0, 4, start, last)
        err = np.abs(img - ref) / (np.abs(ref) + 1e-4)
        assert (err.max(-1) < 1e-3).mean() >= 0.98, (start, last)


@pytest.mark.parametrize("name", sorted(n for n, c in SCENES.items() if c.get("image_spp")))
def test_converged_image_matches_reference_render(name):
    """north_star: converged image vs the reference CPU render within relMSE 1e-3 (different RNG streams)"""
    import os
    from parity import GOLDEN
    path = os.path.join(GOLDEN, "image_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip("no golden image for " + name)
    g = np.load(path)
    ref = g["image"].astype(np.float32)
    cfg = SCENES[name]
    ctx = gpu_scene(name, cfg["image_width"], cfg["image_height"])
    spp = 4096
    img = ctx.render(2024, 0, spp, 0, cfg["last_bounce"]) / spp
    e = rel_mse(img, ref)
    print(name, "relMSE vs reference render (%d spp vs %d spp):" % (spp, int(g["spp"])), e)
    # the reference image itself carries noise of order var/spp_ref; both contribute to the measured relMSE
    # image_noise: the scene's Monte-Carlo noise relative to the configs', from its oracle-vs-oracle noise floor
    assert e <= 1e-3 * cfg.get("image_noise", 1.0) * (1.0 + 4096.0 / float(g["spp"])), e
    assert abs(img.mean() - ref.mean()) <= 0.02 * ref.mean()


def test_error_statuses():
    from pathed_b200 import PathedError, create_context
    ctx = create_context(0)
    ctx.width = ctx.height = 4
    with pytest.raises(PathedError):
        ctx.render(1, 0, 1, 0, 10, accum=np.zeros((4, 4, 3), np.float32))  # not committed
    bad = material_desc(dict(type=0))
    bad.type = 17
    with pytest.raises(PathedError):
        ctx.add_material(bad)
    m = ctx.add_material(material_desc(dict(type=0, diffuse=(1, 1, 1))))
    with pytest.raises(PathedError):
        ctx.add_triangle_mesh([[0, 0, 0], [1, 0, 0], [0, 1, 0]], None, None, [[0, 1, 5]], m)  # index out of range
    with pytest.raises(PathedError):
        ctx.add_sphere((0, 0, 0), 1.0, 9)  # unknown material
    textured = material_desc(dict(type=0, diffuse=(1, 1, 1)))
    textured.albedo_kind, textured.texture = 2, 3
    with pytest.raises(PathedError):
        ctx.add_material(textured)  # texture id out of range
    with pytest.raises(PathedError):
        ctx.add_texture(np.zeros((0, 4, 3), np.uint8))  # Texture::load: "Error loading texture"
    textured.type, textured.texture = 3, ctx.add_texture(np.zeros((2, 2, 3), np.uint8))
    with pytest.raises(PathedError):
        ctx.add_material(textured)  # only Lambertian / Plastic take a texture
    with pytest.raises(PathedError):
        ctx.commit()  # no camera


def test_container_known_answer():
    from pathed_b200 import load_scene
    check_container_known_answer(load_scene("scenes/cornell-medium.json", 32, 32))


def test_media_edge_cases():
    """SURVEY N3 edge cases: status codes of the medium calls, empty query batches, a Passthrough surface WITHOUT a medium (an
    ordinary occluder: the filter only skips containers that enclose a medium, src/scene.cpp:59-63), the VolumePathTracer on a
    scene without media (= no events anywhere), unbounded event lists."""
    from pathed_b200 import PathedError
    from pathed_b200._binding import NO_MEDIUM, PASSTHROUGH, VOLUME_PATH_TRACER, rays_array
    ctx = gpu_context()
    with pytest.raises(PathedError):
        ctx.set_internal_medium(0, 0)  # no such geometry
    with pytest.raises(PathedError):
        ctx.set_integrator(7)
    gas = ctx.add_medium((1.0, 1.0, 1.0), (0.5, 0.5, 0.5))
    wall = ctx.add_material(material_desc(dict(type=0, diffuse=(0.5, 0.5, 0.5))))
    glass = ctx.add_material(material_desc(dict(type=PASSTHROUGH)))
    quad = lambda z: [[-1, -1, z], [1, -1, z], [1, 1, z], [-1, 1, z]]
    far = ctx.add_triangle_mesh(quad(-5.0), None, None, [[0, 1, 2], [0, 2, 3]], wall)
    layers = [ctx.add_triangle_mesh(quad(-1.0 - 0.25 * i), None, None, [[0, 1, 2], [0, 2, 3]], glass) for i in range(12)]
    bare = ctx.add_triangle_mesh(quad(-4.5), None, None, [[0, 1, 2], [0, 2, 3]], glass)  # container material, no medium
    for g in layers:
        ctx.set_internal_medium(g, gas)
    with pytest.raises(PathedError):
        ctx.set_internal_medium(far, 5)  # no such medium
    ctx.set_internal_medium(far, NO_MEDIUM)
    ctx.set_camera((0, 0, 5), (0, 0, 0), (0, 1, 0), 0.5, 8, 8)
    ctx.commit()
    rays = rays_array([[0.1, 0.2, 0.0]], [[0, 0, -1]])
    isects, ne, et, em = ctx.intersect_volumetric(rays)
    assert isects["hit"][0] == 1 and abs(isects["t"][0] - 4.5) < 1e-5  # the bare Passthrough quad is a hit
    # 12 events counted; 8 of them stored (whichever the traversal met first), sorted by t
    assert ne[0] == 12 and (em[0] == gas).all() and (np.diff(et[0]) > 0).all()
    assert all(np.isclose(1.0 + 0.25 * np.arange(12), t).any() for t in et[0])
    occ, ne, et, em = ctx.occluded_volumetric(rays, np.array([4.0], np.float32))
    assert occ[0] == 0 and ne[0] == 12
    occ, ne, et, em = ctx.occluded_volumetric(rays, np.array([4.75], np.float32))
    assert occ[0] == 1 and ne[0] == 0  # the bare quad occludes; events of occluded rays are not reported
    assert ctx.occluded(rays, np.array([4.0], np.float32))[0] == 0 and ctx.occluded(rays, np.array([4.75], np.float32))[0] == 1
    assert ctx.intersect_full(rays)["t"][0] == np.float32(1.0)  # Scene::testIntersect hits containers
    empty = rays_array(np.zeros((0, 3)), np.zeros((0, 3)))
    assert len(ctx.intersect_volumetric(empty)[0]) == 0 and len(ctx.occluded_volumetric(empty, np.zeros(0, np.float32))[0]) == 0
    # no media at all: the volume integrator must agree with the oracle running the same integrator
    cfg = SCENES["cornell_glass"]
    from pathed_b200 import load_scene
    img = load_scene(cfg["scene"], 32, 32, integrator=VOLUME_PATH_TRACER).render(3, 0, 4, 0, 6)
    ref = oracle_scene(cfg["scene"], 32, 32, integrator=VOLUME_PATH_TRACER).render(3, 0, 4, 0, 6)
    err = np.abs(img - ref) / (np.abs(ref) + 1e-3 * ref.mean())
    assert (err.max(-1) < 1e-3).mean() >= 0.97


def test_ragged_inputs_and_limits_match_oracle():
    """Edge cases of the render call: a resolution that is not a multiple of the 8 x 4 path tiles (row-major slot order), one
    sample, a window that starts after the first bounces, lastBounce = -1 (unbounded in the reference, src/bounce_controller.cpp:
    20-25; capped at PTC_MAX_BOUNCES on both sides), zero samples, empty query batches, a mesh with no triangles."""
    from pathed_b200 import PathedError, load_scene
    from pathed_b200._binding import rays_array
    cfg = SCENES["cornell_glass"]
    for (w, h, spp, start, last) in [(37, 23, 3, 0, 10), (40, 22, 1, 2, 5), (16, 12, 2, 0, -1)]:
        img = load_scene(cfg["scene"], w, h).render(9, 0, spp, start, last)
        ref = oracle_scene(cfg["scene"], w, h).render(9, 0, spp, start, last)
        err = np.abs(img - ref) / (np.abs(ref) + 1e-3 * max(ref.mean(), 1e-3))
        assert (err.max(-1) < 1e-3).mean() >= 0.97, (w, h, spp, start, last)
    ctx = load_scene(cfg["scene"], 16, 12)
    before = np.full((12, 16, 3), 0.25, np.float32)
    assert np.array_equal(ctx.render(1, 0, 0, 0, 10, accum=before.copy()), before)  # zero samples: radianceLookup untouched
    with pytest.raises(PathedError):
        ctx.render(1, 0, 1, 5, 2)  # startBounce > lastBounce
    empty = rays_array(np.zeros((0, 3)), np.zeros((0, 3)))
    assert len(ctx.intersect(empty)) == 0 and len(ctx.occluded(empty, np.zeros(0, np.float32))) == 0 and len(ctx.intersect_full(empty)) == 0
    bare = gpu_context()
    m = bare.add_material(material_desc(dict(type=0, diffuse=(1, 1, 1))))
    bare.add_triangle_mesh(np.zeros((0, 3), np.float32), None, None, np.zeros((0, 3), np.uint32), m)  # a geometry with no faces
    bare.add_sphere((0, 0, 0), 1.0, m)
    bare.set_camera((0, 0, 5), (0, 0, 0), (0, 1, 0), 0.5, 8, 8)
    bare.commit()
    hit = bare.intersect(rays_array([[0, 0, 5]], [[0, 0, -1]]))
    assert hit["geom_id"][0] == 1 and abs(hit["t"][0] - 4.0) < 1e-5  # geometry ids follow the attach order, empty mesh included


def test_empty_scene_renders_environment_only():
    ctx = gpu_context()
    env = np.zeros((8, 16, 4), np.float32)
    env[..., :3] = 0.5
    ctx.set_environment(env, 2.0)
    ctx.set_camera((0, 0, 5), (0, 0, 0), (0, 1, 0), 0.5, 8, 8)
    ctx.commit()
    img = ctx.render(1, 0, 2, 0, 10)
    assert np.allclose(img, 2.0)  # 2 spp x (0.5 * scale 2)


def test_reference_known_answers():
    """the reference's own known-answer vectors (tests/known_answers.py) on the CUDA path"""
    import known_answers as ka
    ka.tangent_frame_maps_y_to_normal(gpu_context())
    ka.reflect_follows_the_source(gpu_context())
    from pathed_b200 import load_scene
    ka.one_pixel_environment_map(load_scene("test_scenes/environment_map_sampling.json", 32, 24))


def test_full_size_dragon_properties():
    """BASELINE's full configuration (dragon.json, 1024 x 1024) through size-independent properties: (1) the box pixel filter
    makes a 16 x 16 block average of the full-size render the same estimator as the 64 x 64 render, so 16 spp at 1024^2
    (= 4096 samples per block) must match the reference's converged 64 x 64 image; (2) splitting the samples over two calls
    (what the multi-GPU spp split does) is bit-exact; (3) every sample is finite and the ray counters are consistent."""
    import os
    from parity import GOLDEN
    cfg = SCENES["dragon"]
    g = np.load(os.path.join(GOLDEN, "image_dragon.npz"))
    ref = g["image"].astype(np.float32)
    ctx = gpu_scene("dragon", 1024, 1024)
    ctx.reset_stats()
    img = ctx.render(31, 0, 16, 0, cfg["last_bounce"])
    st = ctx.stats()
    assert np.isfinite(img).all()
    assert st.samples == 1024 * 1024 * 16 and st.closest_rays >= st.samples and st.shadow_rays <= st.closest_rays
    blocks = img.reshape(64, 16, 64, 16, 3).sum((1, 3)) / np.float32(16 * 16 * 16)
    e = rel_mse(blocks, ref)
    print("full-size dragon, block-averaged vs reference 64x64 render: relMSE", e)
    assert e <= 1e-3 * (1.0 + 4096.0 / float(g["spp"])), e
    assert abs(blocks.mean() - ref.mean()) <= 0.02 * ref.mean()
    half = ctx.render(31, 0, 8, 0, cfg["last_bounce"])
    half = ctx.render(31, 8, 8, 0, cfg["last_bounce"], accum=half)
    assert np.array_equal(half, img)


@pytest.mark.parametrize("name,width,height,spp", [("cornell", 512, 512, 64), ("cornell_glass", 512, 512, 64), ("mis", 768, 512, 64),
                                                   ("teapot", 1920, 1080, 16), ("cornell_medium", 512, 512, 64)])
def test_full_size_configs_block_average_to_the_reference_images(name, width, height, spp):
    """The other BASELINE configurations at their full resolutions, through the same size-independent property as the dragon
    test: with the box pixel filter a b x b block average of the full-size render estimates the same integral as one pixel of
    the reference's small converged image (b^2 * spp samples per block)."""
    import os
    from parity import GOLDEN
    cfg = SCENES[name]
    g = np.load(os.path.join(GOLDEN, "image_%s.npz" % name))
    ref = g["image"].astype(np.float32)
    rh, rw = ref.shape[:2]
    b = width // rw
    assert b * rw == width and b * rh == height
    ctx = gpu_scene(name, width, height)
    ctx.reset_stats()
    img = ctx.render(77, 0, spp, 0, cfg["last_bounce"])
    st = ctx.stats()
    assert np.isfinite(img).all() and st.samples == width * height * spp
    blocks = img.reshape(rh, b, rw, b, 3).sum((1, 3)) / np.float32(b * b * spp)
    e = rel_mse(blocks, ref)
    per_block = b * b * spp
    print(name, "%dx%d, %d spp: block-averaged (%d samples per block) vs reference render (%d spp): relMSE" % (width, height, spp, per_block, int(g["spp"])), e)
    assert e <= 1e-3 * cfg.get("image_noise", 1.0) * (4096.0 / per_block + 4096.0 / float(g["spp"])), e
    assert abs(blocks.mean() - ref.mean()) <= 0.02 * ref.mean()


@pytest.mark.parametrize("name", ["cornell_glass", "dragon", "mis"])
def test_device_bvh_builder_matches_host_builder(name):
    """the BVH built on the device (Morton sort + PLOC + wide collapse kernels, the default) and the host
    binned-SAH build must give bit-identical hit records, occlusion flags and images -- a BVH only prunes, and equal-depth ties
    are resolved by primitive index, not by visiting order."""
    from pathed_b200 import load_scene
    cfg = SCENES[name]
    g = golden("scene_" + name)
    dev = load_scene(cfg["scene"], cfg["width"], cfg["height"], options={"bvh_builder": 1})
    host = load_scene(cfg["scene"], cfg["width"], cfg["height"], options={"bvh_builder": 0})
    sd, sh = dev.stats(), host.stats()
    print(name, "device build", sd.bvh_build_ms, "ms", sd.bvh_nodes, "nodes depth", sd.bvh_depth, "ploc iterations", sd.bvh_ploc_iterations,
          "| host build", sh.bvh_build_ms, "ms", sh.bvh_nodes, "nodes depth", sh.bvh_depth)
    assert sd.bvh_builder == 1 and sh.bvh_builder == 0
    assert sd.bvh_triangles == sh.bvh_triangles and sd.bvh_nodes > 0
    for prefix in ("cam_", "sec_"):
        rays = to_rays(g[prefix + "rays"])
        a, b = dev.intersect(rays), host.intersect(rays)
        for field in ("t", "u", "v", "geom_id", "prim_id"):
            assert np.array_equal(a[field], b[field]), (prefix, field)
    shadow = to_rays(g["shadow_rays"])
    assert np.array_equal(dev.occluded(shadow, g["shadow_max_t"]), host.occluded(shadow, g["shadow_max_t"]))
    assert np.array_equal(dev.render(11, 0, 4, 0, cfg["last_bounce"]), host.render(11, 0, 4, 0, cfg["last_bounce"]))


def test_scene_without_lights_renders_black():
    """no emitter and no environment map: Scene::sampleDirectLights is undefined in the reference (m_lights[0] of an empty vector,
    SURVEY Q18); the device path is defined -- no direct lighting, a finite black image -- and agrees with the oracle"""
    from oracle_binding import oracle_context
    imgs = []
    for ctx in (gpu_context(), oracle_context()):
        m = ctx.add_material(material_desc(dict(type=0, diffuse=(0.5, 0.5, 0.5))))
        p = ctx.add_material(material_desc(dict(type=5, diffuse=(0.5, 0.5, 0.5), distribution=0, alpha=0.1)))
        ctx.add_triangle_mesh([[-2, -2, 0], [2, -2, 0], [0, 2, 0]], None, None, [[0, 1, 2]], m)
        ctx.add_triangle_mesh([[-2, -2, -1], [2, -2, -1], [0, 2, -1]], None, None, [[0, 1, 2]], p)
        ctx.set_camera((0, 0, 5), (0, 0, 0), (0, 1, 0), 0.6, 16, 16)
        ctx.commit()
        assert ctx.num_lights() == 0
        imgs.append(ctx.render(3, 0, 4, 0, 5))
    assert np.isfinite(imgs[0]).all() and (imgs[0] == 0).all() and np.array_equal(imgs[0], imgs[1])


def test_philox_known_answers_on_the_device():
    """Random123's kat_vectors for philox4x32-10 through the DEVICE generator (ptc_philox4x32_10), and the uniform streams of path
    vertices (ptc_uniforms: counter = (pixel, sample, bounce, draw / 4), lane draw % 4, top 24 bits) bit for bit against the oracle's"""
    import ctypes
    from oracle_binding import oracle_lib
    ctx = gpu_context()
    counters = [(0, 0, 0, 0), (0xffffffff,) * 4, (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)]
    keys = [(0, 0), (0xffffffff,) * 2, (0xa4093822, 0x299f31d0)]
    want = [(0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)]
    assert np.array_equal(ctx.philox(counters, keys), np.array(want, np.uint32))
    rng = np.random.default_rng(5)
    big_c = rng.integers(0, 2 ** 32, (4096, 4), dtype=np.uint64).astype(np.uint32); big_k = rng.integers(0, 2 ** 32, (4096, 2), dtype=np.uint64).astype(np.uint32)
    got = ctx.philox(big_c, big_k)
    lib = oracle_lib()
    out = (ctypes.c_uint32 * 4)()
    for i in range(0, 4096, 37):
        lib.orc_philox4x32_10((ctypes.c_uint32 * 4)(*big_c[i].tolist()), (ctypes.c_uint32 * 2)(*big_k[i].tolist()), out)
        assert tuple(out) == tuple(got[i].tolist())
    streams = np.array([(0, 0, 0), (3, 5, 1), (1048575, 4095, 64), (0xFFFFFFFF, 0xFFFFFFFF, 7)], np.uint32)
    seed = 0x5EED0123456789AB
    u = ctx.uniforms(seed, streams, 64)
    assert (u >= 0).all() and (u < 1).all()
    for s, (pixel, sample, bounce) in enumerate(streams.tolist()):
        ref = np.array([lib.orc_uniform(ctypes.c_uint64(seed), pixel, sample, bounce, d) for d in range(64)], np.float32)
        assert np.array_equal(u[s], ref)
