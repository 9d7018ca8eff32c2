#!/usr/bin/env bash
# GPU-box session: GPU test suite, then the interleaved-lane sweep (tools/sweep_lanes.py).  usage: tools/gpu_lanes.sh [workload] [steps] [mode]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s_build.log 2>&1
if [ "${LANES_TESTS:-1}" = "1" ]; then ( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/s_pytest.log 2>&1; tail -6 gpurun_out/s_pytest.log; fi
timeout 900 python tools/sweep_lanes.py "${1:-dragon}" "${2:-4}" "${3:-full}" 2> gpurun_out/sweep_lanes.err | grep -v "^Parsing\|^Done\|^Reserving\|EnvironmentLight" | tail -120
tail -5 gpurun_out/sweep_lanes.err
