#!/usr/bin/env python3
"""Builds tuning variants of libpathed_cuda.so HERE (nvcc cross-compiles without a GPU), in parallel, so that the GPU box only
copies and benches them (tools/sweep_prebuilt.sh) instead of spending box time in nvcc.

    python tools/build_variants.py base="" g8="-DPTC_SAMPLE_GROUP=8" ...      ->  variants/<name>/libpathed_cuda.so (+ defines.txt)
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathed_b200 import build as b  # noqa: E402


def one(item):
    name, defs = item
    out = os.path.join(ROOT, "variants", name)
    os.makedirs(out, exist_ok=True)
    target = os.path.join(out, "libpathed_cuda.so")
    cmd = [b._nvcc()] + b.NVCC_FLAGS + defs.split() + ["-o", target] + b._sources(b.CSRC, (".cu",))
    r = subprocess.run(cmd, capture_output=True, text=True)
    open(os.path.join(out, "defines.txt"), "w").write(defs + "\n")
    return name, r.returncode, r.stderr[-2000:]


def main():
    items = [a.split("=", 1) for a in sys.argv[1:]]
    if "--clean" in sys.argv:
        shutil.rmtree(os.path.join(ROOT, "variants"), ignore_errors=True)
        items = [i for i in items if len(i) == 2]
    with ThreadPoolExecutor(max_workers=min(7, max(1, len(items)))) as ex:
        for name, rc, err in ex.map(one, items):
            print(name, "ok" if rc == 0 else "FAILED\n" + err)


if __name__ == "__main__":
    main()
