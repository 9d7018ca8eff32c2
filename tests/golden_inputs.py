"""Seeded inputs shared by tools/make_golden.py (which runs the reference on them) and the parity tests
(which run the oracle / the CUDA path on the very same arrays).  Inputs are regenerated, outputs are stored.

The generator is a splitmix64 hash of the element index, so it does not depend on numpy's RNG streams.
"""
import numpy as np

LAMBERTIAN, OREN_NAYAR, MIRROR, GLASS, MICROFACET, PLASTIC = range(6)
BECKMANN, GGX = 0, 1


def uniform_floats(seed, shape):
    """uniform [0, 1) fp32, exactly representable (24 bits), reproducible everywhere"""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        z = (np.arange(1, n + 1, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15))
        z = z * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)).reshape(shape)


def unit_vectors(seed, n):
    u = uniform_floats(seed, (n, 2)).astype(np.float64)
    z = 1 - 2 * u[:, 0]
    r = np.sqrt(np.maximum(0, 1 - z * z))
    phi = 2 * np.pi * u[:, 1]
    return np.stack([r * np.cos(phi), z, r * np.sin(phi)], 1).astype(np.float32)


def _normalize32(v):
    v = v.astype(np.float32)
    n = np.sqrt((v * v).sum(1, dtype=np.float32)).astype(np.float32)
    return (v / n[:, None]).astype(np.float32)


# every material family the hot path has, with the parameters of the five configs plus the sweep of SURVEY §8(d)
BSDF_CONFIGS = {
    "lambertian": dict(type=LAMBERTIAN, diffuse=(0.725, 0.71, 0.68)),
    "lambertian_emissive": dict(type=LAMBERTIAN, diffuse=(0.78, 0.78, 0.78), emit=(17, 12, 4)),
    "lambertian_checker": dict(type=LAMBERTIAN, diffuse=(1, 1, 1), checker=((0.725, 0.71, 0.68), (0.325, 0.31, 0.25), (20, 20))),
    "oren_nayar_0": dict(type=OREN_NAYAR, diffuse=(0.7, 0.6, 0.5), sigma=0.0),
    "oren_nayar_03": dict(type=OREN_NAYAR, diffuse=(0.7, 0.6, 0.5), sigma=0.3),
    "oren_nayar_1": dict(type=OREN_NAYAR, diffuse=(1.0, 1.0, 1.0), sigma=1.0),
    "mirror": dict(type=MIRROR),
    "glass_11": dict(type=GLASS, ior=1.1),
    "glass_14": dict(type=GLASS, ior=1.4),
    "glass_15": dict(type=GLASS, ior=1.5),
    "glass_20": dict(type=GLASS, ior=2.0),
    "beckmann_0005": dict(type=MICROFACET, distribution=BECKMANN, alpha=0.005),
    "beckmann_002": dict(type=MICROFACET, distribution=BECKMANN, alpha=0.02),
    "beckmann_005": dict(type=MICROFACET, distribution=BECKMANN, alpha=0.05),
    "beckmann_01": dict(type=MICROFACET, distribution=BECKMANN, alpha=0.1),
    "beckmann_05": dict(type=MICROFACET, distribution=BECKMANN, alpha=0.5),
    "ggx_01": dict(type=MICROFACET, distribution=GGX, alpha=0.1),
    "ggx_05": dict(type=MICROFACET, distribution=GGX, alpha=0.5),
    "plastic_dragon": dict(type=PLASTIC, diffuse=(0.1, 0.1, 0.4), distribution=BECKMANN, alpha=0.1),
    "plastic_plate1": dict(type=PLASTIC, diffuse=(0.07, 0.09, 0.13), distribution=BECKMANN, alpha=0.005),
    "plastic_ggx": dict(type=PLASTIC, diffuse=(0.5, 0.4, 0.3), distribution=GGX, alpha=0.3),
}


def material_params(cfg):
    """Parameter block of oracle/ref/probe_main.cpp:ref_material_new."""
    p = np.zeros(18, np.float32)
    p[0:3] = cfg.get("diffuse", (0, 0, 0))
    p[3:6] = cfg.get("emit", (0, 0, 0))
    p[6] = cfg.get("sigma", cfg.get("ior", 0.0))
    p[7] = cfg.get("distribution", 0)
    p[8] = cfg.get("alpha", 0.0)
    if "checker" in cfg:
        on, off, res = cfg["checker"]
        p[9] = 1
        p[10:13] = on; p[13:16] = off; p[16:18] = res
    return p


def material_desc(cfg):
    """The same material as a ptc_material_desc."""
    from pathed_b200._binding import MaterialDesc
    d = MaterialDesc()
    d.type = cfg["type"]
    d.diffuse[:] = cfg.get("diffuse", (0, 0, 0))
    d.emit[:] = cfg.get("emit", (0, 0, 0))
    d.sigma = cfg.get("sigma", 0.0)
    d.ior = cfg.get("ior", 1.4)
    d.distribution = cfg.get("distribution", 0)
    d.alpha = cfg.get("alpha", 0.0)
    if "checker" in cfg:
        on, off, res = cfg["checker"]
        d.albedo_kind = 1
        d.checker_on[:] = on; d.checker_off[:] = off; d.checker_resolution[:] = res
    return d


def bsdf_inputs(name, n):
    """(wo, ng, ns, uv, wi, xi): normals uniform on the sphere; wo/wi uniform on the sphere for the first half
    (exercises every back-side branch), forced into the +n hemisphere for the second half."""
    seed = 1000 + sorted(BSDF_CONFIGS).index(name) * 10
    ns = unit_vectors(seed + 1, n)
    wo = unit_vectors(seed + 2, n)
    wi = unit_vectors(seed + 3, n)
    half = n // 2
    for w in (wo, wi):
        d = (w[half:] * ns[half:]).sum(1, keepdims=True)
        w[half:] = np.where(d < 0, w[half:] - 2 * d * ns[half:], w[half:])
    # third quarter: wi close to the mirror direction of wo, so that narrow glossy lobes are evaluated where they are non-zero
    a, b = half, half + n // 4
    d = (wo[a:b] * ns[a:b]).sum(1, keepdims=True)
    wi[a:b] = 2 * d * ns[a:b] - wo[a:b] + np.float32(0.01) * unit_vectors(seed + 7, b - a)
    wo, wi = _normalize32(wo), _normalize32(wi)
    # a quarter of the tuples get a geometric normal that differs from the shading normal
    ng = ns.copy()
    ng[::4] = _normalize32(ns[::4] + 0.2 * unit_vectors(seed + 4, len(ns[::4])))
    uv = (uniform_floats(seed + 5, (n, 2)) * np.float32(1.5) - np.float32(0.25)).astype(np.float32)
    xi = uniform_floats(seed + 6, (n, 3))
    c = np.ascontiguousarray
    return c(wo), c(ng), c(ns), c(uv), c(wi), c(xi)


def light_inputs(n):
    tri = np.array([-0.24, 1.98, 0.16, -0.24, 1.98, -0.22, 0.23, 1.98, -0.22], np.float32)
    sph = np.array([1.25, 0.0, 0.0, 0.3], np.float32)
    ref = (unit_vectors(77, n) * (uniform_floats(78, (n, 1)) * np.float32(3.0) + np.float32(0.05))).astype(np.float32)
    ref[:, 0] += np.float32(1.0)
    xi2 = uniform_floats(79, (n, 2))
    return tri, sph, np.ascontiguousarray(ref), xi2


# scenes with golden ray / light / path / image fixtures
SCENES = {
    "cornell": dict(scene="scenes/cornell.json", width=64, height=64, last_bounce=10, seed=11, n_rays=4096, n_paths=2048,
                    image_width=64, image_height=64, image_spp=4096),
    "cornell_glass": dict(scene="scenes/cornell-glass.json", width=64, height=64, last_bounce=10, seed=12, n_rays=4096,
                          n_paths=2048, image_width=64, image_height=64, image_spp=4096),
    "mis": dict(scene="scenes/mis-pbrt.json", width=96, height=64, last_bounce=10, seed=13, n_rays=4096, n_paths=2048,
                image_width=96, image_height=64, image_spp=4096),
    "teapot": dict(scene="scenes/teapot.json", width=96, height=54, last_bounce=10, seed=14, n_rays=4096, n_paths=1024,
                   image_width=96, image_height=54, image_spp=2048),
    "dragon": dict(scene="scenes/dragon.json", width=64, height=64, last_bounce=10, seed=15, n_rays=8192, n_paths=1024,
                   image_width=64, image_height=64, image_spp=2048),
    "env_sampling": dict(scene="test_scenes/environment_map_sampling.json", width=64, height=48, last_bounce=4, seed=16,
                         n_rays=2048, n_paths=1024, image_width=64, image_height=48, image_spp=1024),
}


def ray_inputs(name, n):
    """(row, col) film positions, jittered, covering the whole image"""
    cfg = SCENES[name]
    u = uniform_floats(cfg["seed"], (n, 2))
    rc = np.stack([u[:, 0] * np.float32(cfg["height"]) - np.float32(0.5), u[:, 1] * np.float32(cfg["width"]) - np.float32(0.5)], 1)
    return np.ascontiguousarray(rc.astype(np.float32))
