#!/usr/bin/env bash
# One GPU session of the round's second half (run under gpurun): GPU tests, occupancy sweep of the volume material kernels over
# variants built beforehand (tools/build_variants.py), and the size of the builder's SAH top against render speed and build time.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
if [ -d variants ]; then SWEEP_BENCH_ARGS="--workload cornell-medium" bash tools/sweep_prebuilt.sh; cp gpurun_out/sweep.txt gpurun_out/sweep_vmat.txt; fi
for t in 512 64; do
  PTC_BUILD_TIMING=1 PTC_PLOC_TOP=$t timeout 600 python bench.py --no-cpu-baseline --steps 8 2>gpurun_out/top$t.err | tail -1 > gpurun_out/top$t.json
  grep buildWideBVH gpurun_out/top$t.err | head -2
  python -c "
import json; d=json.load(open('gpurun_out/top$t.json')); print($t, d['value'], d['e2e']['value'], d['setup']['bvh_build'], d['roofline']['frac'], d['stages']['extend_ms'], d['stages']['shade_ms'])"
done
