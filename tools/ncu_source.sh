#!/usr/bin/env bash
# Source-level ncu captures (run under gpurun, one GPU): the first launches of every wavefront stage of one small wave, full
# section set with the CUDA source imported, so that `ncu -i <rep> --page source --csv` here attributes instructions, lane
# utilisation and stalls to lines of pathed_b200/csrc.  usage: tools/ncu_source.sh <tag> [spp] [launch count]
set -u
cd "$(dirname "$0")/.."
tag=${1:-r02}; spp=${2:-16}; count=${3:-9}
mkdir -p gpurun_out
timeout 1500 ncu --set full --import-source on --clock-control none -k 'regex:materialKernel|logicKernel|traverseKernel' -c $count \
  -f -o gpurun_out/${tag}_stages python tools/profile_wave.py dragon $spp gpurun_out/${tag}_wave_counts.json > gpurun_out/${tag}_ncu_source.log 2>&1
ls -la gpurun_out/${tag}_stages.ncu-rep
