#!/usr/bin/env bash
# Tuning sweep on the GPU box: rebuilds libpathed_cuda.so with each set of -D defines and prints the bench stage times.
# usage: tools/sweep.sh "<defines A>" "<defines B>" ...     (run under gpurun; results in gpurun_out/sweep.txt)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/sweep.txt
for defs in "$@"; do
  PTC_NVCC_DEFINES="$defs" python -c "from pathed_b200 import build as b; b.build_cuda(force=True); b.build_host(force=True)" >/dev/null 2>&1
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline ${SWEEP_BENCH_ARGS:-} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); s=d['stages'] or {}; r=d['roofline'] or {}
stage=' '.join('%s %8.2f' % (k[:-3], s[k]) for k in ('extend_ms', 'shadow_ms', 'shade_ms', 'volume_kernel_ms') if k in s)
print('%-60s value %7.1f  %s  inner/ray %.2f tri/ray %.2f' % ('$defs', d['value'], stage, r.get('inner_visits_per_ray', 0), r.get('triangle_tests_per_ray', 0)))" | tee -a gpurun_out/sweep.txt
done
python -c "from pathed_b200 import build as b; b.build_cuda(force=True); b.build_host(force=True)" >/dev/null 2>&1
