"""In-tree builds of the two product libraries (explicit compiler command lines, no JIT cache):

  libpathed_cuda.so   hand-written CUDA for sm_100a (csrc/*.cu) behind the C ABI of include/pathed_cuda.h
  libpathed_host.so   C++ host layer mirroring Pathed's Job / parseScene / Image / Integrator (host/*.cpp)
  pathed              the command-line renderer (host/main.cpp), the drop-in for the reference's `./pathed job.json`
"""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
              "-ccbin", GXX, "-Xcompiler", "-fPIC,-O2", "-shared"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(directory, exts):
    return sorted(os.path.join(directory, f) for f in os.listdir(directory) if f.endswith(exts))


def build_cuda(force=False, verbose=False):
    target = os.path.join(PKG, "libpathed_cuda.so")
    deps = _sources(CSRC, (".cu", ".cuh", ".h")) + [os.path.join(os.path.dirname(PKG), "include", "pathed_cuda.h")]
    if force or _stale(target, deps):
        extra = os.environ.get("PTC_NVCC_DEFINES", "").split()  # tuning sweeps: e.g. -DPTC_POSTPONE_DIV=4
        cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", target] + _sources(CSRC, (".cu",))
        subprocess.check_call(cmd)
    return target


def build_host(force=False):
    target = os.path.join(PKG, "libpathed_host.so")
    deps = _sources(HOST, (".cpp", ".hpp"))
    lib_sources = [s for s in _sources(HOST, (".cpp",)) if not s.endswith("main.cpp")]
    cuda = build_cuda()
    if force or _stale(target, deps + [cuda]):
        subprocess.check_call([GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", target] + lib_sources +
                              ["-L" + PKG, "-lpathed_cuda", "-Wl,-rpath,$ORIGIN", "-lz", "-ldl", "-lpthread"])
    return target


def build_cli(force=False):
    main = os.path.join(HOST, "main.cpp")
    if not os.path.exists(main):
        return None
    target = os.path.join(PKG, "pathed")
    if force or _stale(target, _sources(HOST, (".cpp", ".hpp"))):
        subprocess.check_call([GXX, "-O2", "-std=c++17", "-Wall", "-o", target, main, "-L" + PKG, "-lpathed_host", "-lpathed_cuda",
                               "-Wl,-rpath,$ORIGIN", "-lz", "-ldl", "-lpthread"])
    return target


def build_all(force=False):
    build_cuda(force)
    build_host(force)
    build_cli(force)
