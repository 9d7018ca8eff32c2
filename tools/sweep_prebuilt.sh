#!/usr/bin/env bash
# Tuning sweep on the GPU box over variants built beforehand by tools/build_variants.py: copies each variants/<name>/libpathed_cuda.so
# over the product library, runs the bench, prints the stage times; restores the default build at the end.
# usage: tools/sweep_prebuilt.sh [name ...]   (default: every directory under variants/; results in gpurun_out/sweep.txt)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/sweep.txt
cp pathed_b200/libpathed_cuda.so /tmp/libpathed_cuda.default.so
names="$*"; [ -z "$names" ] && names=$(ls variants)
for name in $names; do
  cp variants/$name/libpathed_cuda.so pathed_b200/libpathed_cuda.so
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline ${SWEEP_BENCH_ARGS:-} 2>gpurun_out/sweep_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); s=d['stages'] or {}; r=d['roofline'] or {}
stage=' '.join('%s %8.2f' % (k[:-3], s[k]) for k in ('extend_ms', 'shadow_ms', 'shade_ms', 'other_ms', 'volume_kernel_ms') if k in s)
print('%-28s value %7.1f e2e %7.1f  %s  inner/ray %.2f tri/ray %.2f' % ('$name', d['value'], d['e2e']['value'], stage, r.get('inner_visits_per_ray', 0), r.get('triangle_tests_per_ray', 0)))" | tee -a gpurun_out/sweep.txt
done
cp /tmp/libpathed_cuda.default.so pathed_b200/libpathed_cuda.so
