#!/usr/bin/env bash
# ncu --set full of the extend launch of bounce 1 (16.8 M incoherent rays) and of the shadow launch of bounce 1, with the
# source page, for the library built with the given defines.  usage (under gpurun): tools/profile_traverse.sh <tag> "<defines>"
set -u
cd "$(dirname "$0")/.."
tag=$1; defs=${2:-}
mkdir -p gpurun_out
PTC_NVCC_DEFINES="$defs" python -c "from pathed_b200 import build as b; b.build_cuda(force=True); b.build_host(force=True)" >/dev/null 2>&1
for k in extend shadow; do
  # traverseKernel launches of a wave: extend(0), extend(1), shadow(1), extend(2), shadow(2), ...
  if [ $k = extend ]; then skip=1; else skip=2; fi
  ncu --set full --import-source on --clock-control none -k regex:traverseKernel -s $skip -c 1 -f -o gpurun_out/${tag}_${k} \
      python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${tag}_${k}.log 2>&1
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page raw --csv > gpurun_out/${tag}_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page source --csv > gpurun_out/${tag}_${k}_source.csv 2>/dev/null
done
python -c "from pathed_b200 import build as b; b.build_cuda(force=True); b.build_host(force=True)" >/dev/null 2>&1
