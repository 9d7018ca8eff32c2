"""SURVEY N1 — image textures (reference: src/texture.cpp:12-49, src/scene_parser.cpp:625-651).

CPU part: the host layer's image decoder against what the reference's vendored stb_image returned for the same files
(tests/golden/image_decode.npz, written by tools/make_golden.py through oracle/_ref), the scene parser, and the oracle's
Texture::lookup against the reference (the textured BSDF fixtures run in test_oracle_vs_reference.py).  The GPU part of
N1 runs with the other parity tests (test_gpu_parity.py is parametrised over the "textured" scene and BSDF configs)."""
import json
import os

import numpy as np
import pytest

from golden_inputs import make_test_texture, png_variants, write_png
from pathed_b200 import PathedError
from pathed_b200._binding import REPO_ROOT, SceneFile, load_image_rgb8
from oracle_binding import oracle_context
from parity import golden, make_isects

GOLDEN = os.path.join(REPO_ROOT, "tests", "golden")


def test_png_decoder_reads_the_committed_texture():
    rgb = load_image_rgb8(os.path.join(GOLDEN, "texture_test.png"))
    assert np.array_equal(rgb, make_test_texture())  # all five scanline filters occur in this file


def test_decoder_matches_the_reference_stb_image(tmp_path):
    g = golden("image_decode")
    files = dict(png_variants())
    files["ppm_p6"] = b"P6\n# comment\n5 3\n255\n" + bytes(range(45))
    files["pgm_p5"] = b"P5 4 2 255\n" + bytes(range(100, 108))
    assert sorted(files) == sorted(g.keys())
    for name, data in files.items():
        path = str(tmp_path / name)
        open(path, "wb").write(data)
        assert np.array_equal(load_image_rgb8(path), g[name]), name


def test_decoder_errors_like_texture_load(tmp_path):
    """Texture::load throws "Error loading texture" when stbi_load fails (src/texture.cpp:28-31)"""
    with pytest.raises(PathedError, match="Error loading texture"):
        load_image_rgb8(str(tmp_path / "missing.png"))
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not an image at all")
    with pytest.raises(PathedError, match="Error loading texture"):
        load_image_rgb8(str(bad))
    truncated = tmp_path / "truncated.png"
    truncated.write_bytes(png_variants()["rgb8"][:60])
    with pytest.raises(PathedError, match="Error loading texture"):
        load_image_rgb8(str(truncated))
    jpeg = tmp_path / "photo.jpg"
    jpeg.write_bytes(b"\xff\xd8\xff\xe0" + b"\0" * 32)
    with pytest.raises(PathedError, match="JPEG"):
        load_image_rgb8(str(jpeg))


def _scene_with_texture(tmp_path, bsdf):
    write_png(str(tmp_path / "tex.png"), make_test_texture())
    scene = {"sensor": {"lookAt": {"origin": ["0", "2", "5"], "target": ["0", "0", "0"], "up": ["0", "1", "0"]}, "fov": "40"},
             "models": [{"type": "quad", "bsdf": bsdf},
                        {"type": "quad", "transform": {"translate": ["0", "1", "0"]},
                         "bsdf": {"type": "lambertian", "texture": "tex.png", "diffuseReflectance": ["1", "1", "1"]}}]}
    json.dump(scene, open(tmp_path / "scene.json", "w"))
    return SceneFile("scene.json", 16, 16, str(tmp_path))


def test_parser_registers_textures_for_lambertian_and_plastic(tmp_path):
    scene = _scene_with_texture(tmp_path, {"type": "plastic", "texture": "tex.png", "diffuseReflectance": ["0.5", "0.5", "0.5"],
                                           "distribution": {"type": "ggx", "alpha": "0.2"}})
    plastic, lambertian = scene.material(0), scene.material(1)
    assert (plastic.type, plastic.albedo_kind, plastic.texture) == (5, 2, 0)
    assert (lambertian.type, lambertian.albedo_kind, lambertian.texture) == (0, 2, 0)  # same file: decoded once, shared
    # the texture wins over a checkerboard "albedo" (src/scene_parser.cpp:645-651)
    scene = _scene_with_texture(tmp_path, {"type": "lambertian", "texture": "tex.png", "diffuseReflectance": ["1", "1", "1"],
                                           "albedo": {"type": "checkerboard", "onColor": ["1", "1", "1"], "offColor": ["0", "0", "0"],
                                                      "resolution": {"u": "2", "v": "2"}}})
    assert scene.material(0).albedo_kind == 2
    # feeding the parsed scene registers the texture before the materials that name it
    api = scene.feed(oracle_context())
    assert api.num_lights() == 0


def test_parser_reports_a_missing_texture(tmp_path):
    scene = {"sensor": {"lookAt": {"origin": ["0", "2", "5"], "target": ["0", "0", "0"], "up": ["0", "1", "0"]}, "fov": "40"},
             "models": [{"type": "quad", "bsdf": {"type": "lambertian", "texture": "nope.png", "diffuseReflectance": ["1", "1", "1"]}}]}
    json.dump(scene, open(tmp_path / "scene.json", "w"))
    with pytest.raises(PathedError, match="Error loading texture"):
        SceneFile("scene.json", 16, 16, str(tmp_path))


def test_texture_lookup_wrap_flip_and_nearest_texel():
    """Texture::lookup (src/texture.cpp:34-49) on hand-computed cases: u wraps by floor, v is flipped after wrapping,
    the texel is roundf(u * (w - 1)), roundf(v * (h - 1)) and the colour pow(c / 255, 2.2)"""
    from golden_inputs import LAMBERTIAN, material_desc
    tex = make_test_texture()
    h, w, _ = tex.shape
    o = oracle_context()
    mat = o.add_material(material_desc(dict(type=LAMBERTIAN, diffuse=(1, 1, 1), textured=True), o))
    uv = np.array([[0.0, 0.0], [0.999, 0.999], [0.5, 0.25], [1.25, -0.25], [-0.3, 2.6], [0.0138, 0.9773]], np.float32)
    n = len(uv)
    up = np.tile(np.array([[0, 1, 0]], np.float32), (n, 1))
    isects = make_isects(up, up, up, uv, mat)
    f, pdf = o.bsdf_eval(mat, isects, up)
    for i, (u, v) in enumerate(uv):
        uu = np.float32(u) - np.float32(int(np.floor(u)))
        vv = np.float32(1) - (np.float32(v) - np.float32(int(np.floor(v))))
        x = int(np.floor(np.float32(uu * np.float32(w - 1)) + np.float32(0.5)))
        y = int(np.floor(np.float32(vv * np.float32(h - 1)) + np.float32(0.5)))
        want = (tex[y, x].astype(np.float32) / np.float32(255)) ** np.float32(2.2) / np.float32(np.pi)
        assert np.allclose(f[i], want, rtol=2e-6, atol=0), (i, f[i], want)
    assert np.allclose(pdf, 1 / np.pi, rtol=1e-6)


def test_texture_misuse_is_rejected():
    from golden_inputs import GLASS, LAMBERTIAN, material_desc
    o = oracle_context()
    d = material_desc(dict(type=LAMBERTIAN, diffuse=(1, 1, 1)))
    d.albedo_kind, d.texture = 2, 5  # no such texture
    with pytest.raises(PathedError):
        o.add_material(d)
    tex = o.add_texture(make_test_texture())
    d = material_desc(dict(type=GLASS))
    d.albedo_kind, d.texture = 2, tex  # only Lambertian and Plastic take a texture
    with pytest.raises(PathedError):
        o.add_material(d)
