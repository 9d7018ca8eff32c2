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
    # N1 image textures (src/texture.cpp): the albedo comes from tests/golden/texture_test.png, diffuse is ignored
    "lambertian_textured": dict(type=LAMBERTIAN, diffuse=(1, 1, 1), textured=True),
    "plastic_textured": dict(type=PLASTIC, diffuse=(1, 1, 1), distribution=BECKMANN, alpha=0.1, textured=True),
}


# input seeds are tied to this order: the round-1 configs sorted by name, later additions appended (keeps old fixtures stable)
_LATER_CONFIGS = ["lambertian_textured", "plastic_textured"]
_SEED_ORDER = sorted(n for n in BSDF_CONFIGS if n not in _LATER_CONFIGS) + _LATER_CONFIGS


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


def material_desc(cfg, api=None):
    """The same material as a ptc_material_desc; a textured config registers test_texture() with `api` first."""
    if cfg.get("textured"):
        cfg = dict(cfg, texture_id=api.add_texture(make_test_texture()))
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
    if "texture_id" in cfg:  # the caller registered test_texture() with add_texture first
        d.albedo_kind = 2
        d.texture = cfg["texture_id"]
    return d


def write_png(path, rgb):
    """8-bit RGB, non-interlaced PNG with per-row filter type 0..4 cycling (exercises every unfilter of a decoder)."""
    import struct
    import zlib
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, _ = rgb.shape
    raw = bytearray()
    prev = np.zeros(w * 3, np.int32)
    for y in range(h):
        cur = rgb[y].reshape(-1).astype(np.int32)
        left = np.concatenate([np.zeros(3, np.int32), cur[:-3]])
        upleft = np.concatenate([np.zeros(3, np.int32), prev[:-3]])
        ft = y % 5
        if ft == 0: out = cur
        elif ft == 1: out = cur - left
        elif ft == 2: out = cur - prev
        elif ft == 3: out = cur - ((left + prev) >> 1)
        else:
            p = left + prev - upleft
            pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - upleft)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, upleft))
            out = cur - pred
        raw.append(ft)
        raw += (out & 255).astype(np.uint8).tobytes()
        prev = cur

    def chunk(kind, data):
        return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(bytes(raw), 9)) + chunk(b"IEND", b""))


def png_variants():
    """name -> PNG file bytes covering what a decoder must handle: every colour type, bit depths 1-16, Adam7 interlacing,
    a palette with fewer than 256 entries, several IDAT chunks.  Filter type 0 only (write_png covers the filters)."""
    import struct
    import zlib

    def chunk(kind, data):
        return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xFFFFFFFF)

    def pack_rows(samples, depth):  # samples: (h, w * channels) integers
        rows = []
        for row in samples:
            if depth == 8:
                rows.append(bytes(int(v) for v in row))
            elif depth == 16:
                rows.append(b"".join(struct.pack(">H", int(v)) for v in row))
            else:
                bits = "".join(format(int(v), "0%db" % depth) for v in row)
                bits += "0" * (-len(bits) % 8)
                rows.append(bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8)))
        return rows

    def encode(samples, w, h, channels, depth, color_type, interlace=False, palette=None, split=1):
        samples = np.asarray(samples).reshape(h, w, channels)
        raw = b""
        if not interlace:
            raw = b"".join(b"\x00" + r for r in pack_rows(samples.reshape(h, w * channels), depth))
        else:
            for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
                sub = samples[y0::dy, x0::dx]
                if sub.shape[0] and sub.shape[1]:
                    raw += b"".join(b"\x00" + r for r in pack_rows(sub.reshape(sub.shape[0], -1), depth))
        z = zlib.compress(raw, 6)
        parts = [z[i * len(z) // split:(i + 1) * len(z) // split] for i in range(split)]
        out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color_type, 0, 0, 1 if interlace else 0))
        if palette is not None:
            out += chunk(b"PLTE", bytes(int(v) for v in np.asarray(palette).reshape(-1)))
        return out + b"".join(chunk(b"IDAT", part) for part in parts) + chunk(b"IEND", b"")

    w, h = 13, 9
    u = (uniform_floats(4242, (h, w, 4)) * 65536).astype(np.int64)
    out = {
        "rgb8": encode(u[..., :3] >> 8, w, h, 3, 8, 2),
        "rgb8_interlaced": encode(u[..., :3] >> 8, w, h, 3, 8, 2, interlace=True, split=3),
        "rgba8": encode(u >> 8, w, h, 4, 8, 6),
        "rgb16": encode(u[..., :3], w, h, 3, 16, 2),
        "rgba16_interlaced": encode(u, w, h, 4, 16, 6, interlace=True),
        "gray8": encode(u[..., :1] >> 8, w, h, 1, 8, 0),
        "gray16": encode(u[..., :1], w, h, 1, 16, 0),
        "gray_alpha8": encode(u[..., :2] >> 8, w, h, 2, 8, 4),
        "gray4": encode(u[..., :1] >> 12, w, h, 1, 4, 0),
        "gray2_interlaced": encode(u[..., :1] >> 14, w, h, 1, 2, 0, interlace=True),
        "gray1": encode(u[..., :1] >> 15, w, h, 1, 1, 0),
        "palette8": encode(u[..., :1] % 200, w, h, 1, 8, 3, palette=(uniform_floats(4343, (200, 3)) * 256).astype(np.int64)),
        "palette4_interlaced": encode(u[..., :1] >> 12, w, h, 1, 4, 3, interlace=True, palette=(uniform_floats(4344, (16, 3)) * 256).astype(np.int64)),
        "palette1": encode(u[..., :1] >> 15, w, h, 1, 1, 3, palette=[[10, 20, 30], [200, 210, 220]]),
    }
    return out


def make_test_texture(width=37, height=23):
    """Deterministic 8-bit RGB test image (row 0 = top): smooth ramps plus a hash pattern, every byte value occurs."""
    y, x = np.mgrid[0:height, 0:width].astype(np.uint32)
    h = (x * np.uint32(2654435761) + y * np.uint32(40503) + np.uint32(12345)) >> np.uint32(7)
    rgb = np.stack([(x * 255 // (width - 1)) ^ (h & 31), (y * 255 // (height - 1)) ^ ((h >> 5) & 63), (h >> 11) & 255], -1)
    return np.ascontiguousarray(rgb & 255, dtype=np.uint8)


def bsdf_inputs(name, n):
    """(wo, ng, ns, uv, wi, xi): normals uniform on the sphere; wo/wi uniform on the sphere for the first half
    (exercises every back-side branch), forced into the +n hemisphere for the second half."""
    seed = 1000 + _SEED_ORDER.index(name) * 10
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
    # image textures on Lambertian and Plastic (PNG assets, uv outside [0, 1] on the box), area light
    "textured": dict(scene="scenes/textured.json", width=64, height=48, last_bounce=10, seed=17, n_rays=4096, n_paths=2048,
                     image_width=64, image_height=48, image_spp=2048),
    "env_sampling": dict(scene="test_scenes/environment_map_sampling.json", width=64, height=48, last_bounce=4, seed=16,
                         n_rays=2048, n_paths=1024, image_width=64, image_height=48, image_spp=1024),
    # participating media.  integrator = 1: VolumePathTracer (paths and images); the ray fixtures add the
    # volumetric queries (testVolumetricIntersect / testVolumetricOcclusion with their volume events)
    "cornell_medium": dict(scene="scenes/cornell-medium.json", width=64, height=64, last_bounce=10, seed=18, n_rays=4096,
                           n_paths=2048, image_width=64, image_height=64, image_spp=2048, integrator=1),
    "medium_sphere": dict(scene="test_scenes/medium_sphere.json", width=64, height=48, last_bounce=8, seed=19, n_rays=4096,
                          n_paths=2048, image_width=64, image_height=48, image_spp=2048, integrator=1),
    # the same container scene under the plain PathTracer: Passthrough vertices + the occlusion filter in testOcclusion (A8f)
    # image_noise: two independent 1024-spp oracle renders of this scene differ by relMSE 2.1e-2 (light is only found through
    # BSDF hits behind the delta Passthrough vertices), i.e. 2.6e-3 at 4096 spp against the 1e-3 the north-star gate assumes
    "cornell_medium_pt": dict(scene="scenes/cornell-medium.json", width=64, height=64, last_bounce=10, seed=20, n_rays=2048,
                              n_paths=2048, image_width=64, image_height=64, image_spp=2048, integrator=0, image_noise=3.0),
}
# hierarchical instancing -- `instance` / `instanced` models, two levels, rotated / scaled / mirrored placements; the ray
# fixtures carry RTCHit::instID
SCENES["instanced"] = dict(scene="scenes/instanced.json", width=96, height=72, last_bounce=10, seed=21, n_rays=8192, n_paths=2048,
                           image_width=96, image_height=72, image_spp=2048, instanced=True)
INTEGRATOR_NAMES = {0: "PathTracer", 1: "VolumePathTracer"}


def ray_inputs(name, n):
    """(row, col) film positions, jittered, covering the whole image"""
    cfg = SCENES[name]
    u = uniform_floats(cfg["seed"], (n, 2))
    rc = np.stack([u[:, 0] * np.float32(cfg["height"]) - np.float32(0.5), u[:, 1] * np.float32(cfg["width"]) - np.float32(0.5)], 1)
    return np.ascontiguousarray(rc.astype(np.float32))
