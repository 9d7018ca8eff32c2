"""CPU tests of the C++ host API (pathed_b200/host/pathed.hpp): Job, BounceController, Image behave like the reference's
(/root/reference/include/job.h, src/job.cpp, src/bounce_controller.cpp, src/image.cpp)."""
import json
import os
import subprocess

import numpy as np
import pytest

from pathed_b200 import PathedError, bounce_controller, image_save, job_describe, read_exr
from pathed_b200._binding import PKG_DIR, REPO_ROOT


def _job(tmp_path, **over):
    job = {"spp": 4, "integrator": "PathTracer", "scene": "scenes/cornell.json", "startBounce": 0, "lastBounce": 10,
           "output_directory": str(tmp_path / "out"), "output_name": "final", "showUI": False, "force": True, "width": 32, "height": 24}
    job.update(over)
    for k in [k for k, v in job.items() if v is None]:
        del job[k]
    path = str(tmp_path / "job.json")
    json.dump(job, open(path, "w"))
    return path


def test_job_accessors_and_defaults(tmp_path):
    d = job_describe(_job(tmp_path))
    assert (d["width"], d["height"], d["spp"], d["startBounce"], d["lastBounce"]) == (32, 24, 4, 0, 10)
    assert d["output_directory"].endswith("/out/") and d["output_name"] == "final" and d["scene"] == "scenes/cornell.json"
    assert d["showUI"] is False and d["force"] is True and d["integrator_status"] == "ok"
    assert (d["gpus"], d["seed"], d["wave_spp"]) == (1, 0x5EED, 64)
    # spp <= 0 means "until stopped" (include/job.h:27-33); force defaults to false when absent
    d = job_describe(_job(tmp_path, spp=0, force=None, gpus=4, seed=99, lastBounce=-1))
    assert d["spp"] == 9999999 and d["force"] is False and d["gpus"] == 4 and d["seed"] == 99 and d["lastBounce"] == -1


def test_job_unknown_integrator_is_unimplemented(tmp_path):
    # src/job.cpp:96: throw "Unimplemented"; the research integrators are outside the accelerated path
    for name in ("BDPT", "LightTracer", "NoSuchThing"):
        assert job_describe(_job(tmp_path, integrator=name))["integrator_status"] == "Unimplemented"
    # "VolumePathTracer" (src/job.cpp:71-72) is built
    assert job_describe(_job(tmp_path, integrator="VolumePathTracer"))["integrator_status"] == "ok"


def test_job_missing_keys_raise(tmp_path):
    with pytest.raises(PathedError):
        job_describe(_job(tmp_path, startBounce=None))
    with pytest.raises(PathedError):
        job_describe(str(tmp_path / "nope.json"))


def test_bounce_controller_windows():
    # src/bounce_controller.cpp:14-25
    assert bounce_controller(0, 10, 0)[:2] == (True, False)
    assert bounce_controller(0, 10, 10)[:2] == (True, False)
    assert bounce_controller(0, 10, 11)[:2] == (False, True)
    assert bounce_controller(2, 3, 1)[:2] == (False, False)
    assert bounce_controller(2, -1, 1000)[:2] == (True, False)
    assert bounce_controller(2, 3, 0)[2] == (1, 2)
    assert bounce_controller(0, 0, 0)[2] == (0, 0)
    assert bounce_controller(0, -1, 0)[2] == (0, -1)


def test_image_layout_exr_and_preview(tmp_path):
    """src/image.cpp: raw stored flipped (row 0 = bottom scanline), EXR = HALF B,G,R, checkpoint name <stem>-%05dspp.exr,
    preview = min(v^(1/2.2), 1) * 255 truncated"""
    rng = np.random.default_rng(3)
    rgb = (rng.random((6, 5, 3)) * 2).astype(np.float32)
    rgb[0, 0] = (0.25, 0.5, 4.0)
    out = str(tmp_path)
    preview = image_save(out, "auto", rgb, 8, "preview.bmp")
    for name in ("auto.exr", "auto-00008spp.exr"):
        exr = read_exr(os.path.join(out, name))
        assert exr.shape == (6, 5, 4)
        want = rgb[::-1].astype(np.float16).astype(np.float32)  # top scanline first, HALF
        assert np.array_equal(exr[..., :3], want)
    want8 = (np.minimum(np.power(rgb, np.float32(1 / 2.2), dtype=np.float32), 1.0) * 255).astype(np.uint8)
    assert np.abs(preview.astype(int) - want8.astype(int)).max() <= 1
    assert tuple(preview[0, 0]) == (int(0.25 ** (1 / 2.2) * 255), int(0.5 ** (1 / 2.2) * 255), 255)
    bmp = open(os.path.join(out, "preview.bmp"), "rb").read()
    assert bmp[:2] == b"BM" and len(bmp) == 54 + 6 * 16  # 5 px * 3 B padded to 16 B per row
    # stb's BMP puts data row 0 at the top of the picture = the LAST stored row; stored order is B,G,R
    last_row = bmp[54 + 5 * 16:54 + 5 * 16 + 3]
    assert tuple(last_row) == (preview[0, 0, 2], preview[0, 0, 1], preview[0, 0, 0])


def test_cli_fails_loudly_without_gpu(tmp_path):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    r = subprocess.run([os.path.join(PKG_DIR, "pathed"), _job(tmp_path), "--root", REPO_ROOT], capture_output=True, text=True)
    assert r.returncode == 1 and "Failed to create device" in r.stdout
    assert os.path.exists(str(tmp_path / "out" / "report.json"))  # Job::init ran first, as in app/main.cpp:66-73
