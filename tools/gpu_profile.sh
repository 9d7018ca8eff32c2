#!/usr/bin/env bash
# Profiles of one wave of the dragon workload (run under gpurun, one GPU).  usage: tools/gpu_profile.sh <tag>
#  1. per-launch metrics of every kernel of a 64-spp wave (gpu time, DRAM / L2 bytes, issue utilisation, lanes, hit rates, occupancy)
#  2. --set full with sources of the first stage launches of a 16-spp wave (bounces 0 and 1 of every stage)
set -u
cd "$(dirname "$0")/.."
tag=${1:-r02}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/p_build.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__t_bytes.sum
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${tag}_wave_metrics.csv python tools/profile_wave.py dragon 64 gpurun_out/${tag}_wave_counts.json > gpurun_out/${tag}_wave.log 2>&1
tail -1 gpurun_out/${tag}_wave.log
timeout 1200 ncu --set full --import-source on --clock-control none -k 'regex:materialKernel|logicKernel|traverseKernel' -c 9 \
  -f -o gpurun_out/${tag}_stages python tools/profile_wave.py dragon 16 gpurun_out/${tag}_wave16_counts.json > gpurun_out/${tag}_stages.log 2>&1
ls -la gpurun_out/${tag}_stages.ncu-rep
# 2b. per-launch metrics of one VolumePathTracer wave (cornell-medium, 64 spp): the wavefront stages of volume_wavefront.cuh
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${tag}_volume_wave_metrics.csv python tools/profile_wave.py cornell-medium 64 gpurun_out/${tag}_volume_wave_counts.json > gpurun_out/${tag}_volume_wave.log 2>&1
tail -1 gpurun_out/${tag}_volume_wave.log
# 3. L2 -> SM read bandwidth ceiling (tools/l2_bandwidth.cu), for roofline.l2_frac
if [ -f tools/l2_bandwidth.cu ]; then
  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/l2_bandwidth tools/l2_bandwidth.cu > gpurun_out/${tag}_l2_build.log 2>&1 && /tmp/l2_bandwidth > gpurun_out/${tag}_l2_peak.json 2> gpurun_out/${tag}_l2.err
  cat gpurun_out/${tag}_l2_peak.json | cut -c1-300
fi
