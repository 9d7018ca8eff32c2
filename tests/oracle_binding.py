"""Test-side loaders for the CPU checker (oracle/liboracle.so) and the compiled reference probe (oracle/_ref)."""
import ctypes
import os
import subprocess

import numpy as np

from pathed_b200._binding import Api, REPO_ROOT, SceneFile

ORACLE_DIR = os.path.join(REPO_ROOT, "oracle")
_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"])
        _oracle = ctypes.CDLL(path)
        _oracle.orc_uniform.restype = ctypes.c_float
    return _oracle


def oracle_context():
    return Api(oracle_lib(), "orc_")


def oracle_scene(scene_json, width, height, root=REPO_ROOT, integrator=0):
    api = oracle_context()
    SceneFile(scene_json, width, height, root).feed(api)
    api.set_integrator(integrator)
    return api


REF_PROBE = os.path.join(ORACLE_DIR, "_ref", "libpathed_ref_probe.so")
REF_HEADLESS = os.path.join(ORACLE_DIR, "_ref", "pathed_ref_headless")


def have_reference():
    return os.path.exists(REF_PROBE)
