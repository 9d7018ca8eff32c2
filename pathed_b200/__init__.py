"""pathed_b200 — B200-native replacement for Pathed's surface path-tracing hot path.

The product is the CUDA library `libpathed_cuda.so` (C ABI: include/pathed_cuda.h) plus the C++ host layer
`libpathed_host.so` mirroring Pathed's Job / parseScene / Image / Integrator API.  This package only holds
thin ctypes bindings for scripts (tests, bench.py); there is no Python or CPU compute path.
"""
from ._binding import (Api, MaterialDesc, PathedError, SceneFile, Stats, create_context, cuda_lib, host_lib,  # noqa: F401
                       load_scene, rays_array, job_describe, bounce_controller, image_save, read_exr, scene_query, RAY_DTYPE, HIT_DTYPE, ISECT_DTYPE, LIGHT_SAMPLE_DTYPE)
