#!/usr/bin/env bash
# the BASELINE jobs through the command-line renderer on N GPUs of the box (default 1): time to image and its breakdown
set -u
N=${1:-1}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c_build.log 2>&1
job() { # name scene width height spp gpus
  mkdir -p /tmp/job_$1
  cat > /tmp/job_$1/job.json <<JSON
{"spp": $5, "integrator": "PathTracer", "scene": "$2", "startBounce": 0, "lastBounce": 10, "output_directory": "/tmp/job_$1/out",
 "output_name": "final", "showUI": false, "force": true, "width": $3, "height": $4, "gpus": $6}
JSON
  ( time pathed_b200/pathed /tmp/job_$1/job.json --root "$PWD" ) > gpurun_out/r02_cli_$1.log 2>&1
  echo "== $1"; grep -E "^Scene|PATHED_RESULT|^real" gpurun_out/r02_cli_$1.log
}
job dragon_${N}gpu scenes/dragon.json 1024 1024 256 $N
job dragon_${N}gpu_again scenes/dragon.json 1024 1024 256 $N
job teapot_${N}gpu scenes/teapot.json 1920 1080 1024 $N
