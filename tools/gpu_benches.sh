#!/usr/bin/env bash
# Bench lines of every BASELINE configuration on one GPU (run under gpurun): the default line (dragon, with cpu_baseline), the reference
# arm, and the other workloads without the CPU leg.  Results: gpurun_out/r02_bench_<workload>.json
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/b_build.log 2>&1
timeout 900 python bench.py > gpurun_out/r02_bench_dragon.log 2> gpurun_out/r02_bench_dragon.err; tail -1 gpurun_out/r02_bench_dragon.log > gpurun_out/r02_bench_dragon.json
timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r02_bench_reference_arm.log 2> gpurun_out/r02_bench_reference_arm.err; tail -1 gpurun_out/r02_bench_reference_arm.log > gpurun_out/r02_bench_reference_arm.json
for w in teapot cornell cornell-glass mis-pbrt cornell-medium; do
  timeout 600 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_$w.log 2> gpurun_out/r02_bench_$w.err; tail -1 gpurun_out/r02_bench_$w.log > gpurun_out/r02_bench_$w.json
done
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02_bench_*.json")):
    try:
        d = json.load(open(p))
        r = d.get("roofline") or {}
        print("%-44s value %8.1f e2e %8.1f %s  roofline %.2f traffic %s cpu %s" % (p.split("/")[-1], d["value"], d["e2e"]["value"], d["unit"], r.get("frac", 0), r.get("traffic"), (d.get("cpu_baseline") or {}).get("value")))
    except Exception as e:
        print(p, "unreadable:", e)
PY
