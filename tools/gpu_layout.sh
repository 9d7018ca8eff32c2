#!/usr/bin/env bash
# Path-state layout experiment (run under gpurun): separate allocations vs one slab with different gaps between the arrays
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/layout.txt
run() { # label workload steps env...
  label=$1; w=$2; steps=$3; shift 3
  env "$@" timeout 300 python bench.py --workload $w --steps $steps --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stages']; print('%-34s %-8s value %7.1f e2e %7.1f extend %7.2f shadow %7.2f shade %7.2f' % ('$label', '$w', d['value'], d['e2e']['value'], s['extend_ms'], s['shadow_ms'], s['shade_ms']))" | tee -a gpurun_out/layout.txt
}
for w in ${LAYOUT_WORKLOADS:-cornell dragon}; do
  steps=8; [ $w = dragon ] && steps=6
  run "separate mallocs" $w $steps PTC_PATH_SLAB=0
  run "separate, no second build" $w $steps PTC_PATH_SLAB=0 PTC_BENCH_NO_SECOND_BUILD=1
  for g in ${LAYOUT_STAGGERS:-0 256 9472 66816 2106624}; do
    run "slab gap $g" $w $steps PTC_PATH_STAGGER=$g
  done
  run "slab gap 9472, no second build" $w $steps PTC_PATH_STAGGER=9472 PTC_BENCH_NO_SECOND_BUILD=1
done
