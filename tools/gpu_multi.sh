#!/usr/bin/env bash
# Multi-GPU session (run under `gpurun --gpus N`): the multi-GPU tests, the strong-scaling bench of the BASELINE jobs (spp split
# over N ranks, NCCL reduce) and the same jobs through the in-process path of the CLI (`pathed` with "gpus": N: ptc_replicate,
# per-GPU framebuffers, one peer-memory gather + resolve per image).  Results: gpurun_out/r02_multi_*.
set -u
N=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/m_build.log 2>&1
( timeout 900 python -m pytest tests -m gpu -q -k "multi_gpu or replicated or framebuffer or checkpoint" 2>&1 | tail -5 ) > gpurun_out/r02_multi_pytest.log 2>&1
tail -2 gpurun_out/r02_multi_pytest.log
run_bench() { # name, args...
  local name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" \
    > gpurun_out/r02_multi_$name.log 2> gpurun_out/r02_multi_$name.err
  tail -1 gpurun_out/r02_multi_$name.log > gpurun_out/r02_multi_$name.json
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_multi_$name.json"))
print("$name", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 2), d["config"]["parallelism"])
PY
}
run_bench teapot_strong_${N}gpu --workload teapot --scaling strong --steps 3 --warmup 3 --no-cpu-baseline
run_bench dragon_strong_${N}gpu --workload dragon --scaling strong --steps 6 --warmup 3 --no-cpu-baseline
# the CLI on the BASELINE jobs: one process, N devices
job() { # name scene width height spp gpus
  mkdir -p /tmp/job_$1
  cat > /tmp/job_$1/job.json <<JSON
{"spp": $5, "integrator": "PathTracer", "scene": "$2", "startBounce": 0, "lastBounce": 10, "output_directory": "/tmp/job_$1/out",
 "output_name": "final", "showUI": false, "force": true, "width": $3, "height": $4, "gpus": $6}
JSON
  ( time pathed_b200/pathed /tmp/job_$1/job.json --root "$PWD" ) > gpurun_out/r02_multi_cli_$1.log 2>&1
  grep -E "Scene ready|PATHED_RESULT|^real" gpurun_out/r02_multi_cli_$1.log
}
job teapot_${N}gpu scenes/teapot.json 1920 1080 1024 $N
job teapot_1gpu scenes/teapot.json 1920 1080 1024 1
job dragon_${N}gpu scenes/dragon.json 1024 1024 256 $N
job dragon_1gpu scenes/dragon.json 1024 1024 256 1
python - <<'PY'
# the N-GPU image against the 1-GPU image of the same job (same Philox streams; sums differ in fp32 order only)
import numpy as np
from pathed_b200 import read_exr
import glob
for scene in ("teapot", "dragon"):
    many = [p for p in glob.glob("/tmp/job_%s_*gpu/out/final.exr" % scene) if "_1gpu" not in p]
    if many:
        a = read_exr(many[0]); b = read_exr("/tmp/job_%s_1gpu/out/final.exr" % scene)
        ok = np.isfinite(a) & np.isfinite(b)
        print(scene, "N-GPU vs 1-GPU final.exr: max abs diff", float(np.abs(a - b)[ok].max()), "mean", float(b[ok].mean()), "non-finite", int((~ok).sum()))
PY
