#!/usr/bin/env bash
# compute-sanitizer over small renders of every kernel family (run under gpurun, one GPU): memcheck, then racecheck -- the queue
# kernels append with warp-aggregated atomics and hand work between stages through device-side counters (race
# detection).  Summaries land in gpurun_out/sanitizer_<tool>.log; copy them to profiles/ for the record.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/sanitize_me.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from pathed_b200 import load_scene
from pathed_b200._binding import VOLUME_PATH_TRACER
for scene, integ in (("scenes/cornell-glass.json", 0), ("scenes/mis-pbrt.json", 0), ("scenes/instanced.json", 0), ("scenes/cornell-medium.json", 0), ("scenes/cornell-medium.json", VOLUME_PATH_TRACER)):
    ctx = load_scene(scene, 32, 24, integrator=integ)
    img = ctx.render(3, 0, 2, 0, 6)
    ctx.framebuffer_clear(); ctx.framebuffer_render_checkpoints(3, 0, 3, 0, 6, [1, 2])
    out = ctx.framebuffer_gather([], divisor=3)
    print(scene, integ, float(img.mean()), float(out.mean()))
    ctx.close()
PY
for tool in ${SANITIZER_TOOLS:-memcheck racecheck}; do
  timeout ${SANITIZER_TIMEOUT:-1500} compute-sanitizer --tool $tool --print-limit 20 python /tmp/sanitize_me.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1)"
done
