#!/usr/bin/env bash
# Last GPU session of a round (run under gpurun, one GPU): GPU tests, the per-launch ncu metrics of one wave (-> profiles/traffic.json
# through tools/ncu_traffic.py, stage counters through tools/stage_summary.py), the default bench line, builder timing.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s_build.log 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/s_pytest.log 2>&1; tail -6 gpurun_out/s_pytest.log
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__t_bytes.sum
timeout 400 ncu --metrics $M --clock-control none -k 'regex:traverseKernel|logicKernel|materialKernel|generateKernel|accumulateKernel' --csv --log-file gpurun_out/final_wave_metrics.csv python tools/profile_wave.py dragon 64 gpurun_out/final_wave_counts.json > gpurun_out/final_wave.log 2>&1
tail -1 gpurun_out/final_wave.log
PTC_BUILD_TIMING=1 timeout 300 python bench.py > gpurun_out/final_bench_dragon.log 2> gpurun_out/final_bench_dragon.err; tail -1 gpurun_out/final_bench_dragon.log > gpurun_out/final_bench_dragon.json
grep buildWideBVH gpurun_out/final_bench_dragon.err | head -3
python -c "
import json; d=json.load(open('gpurun_out/final_bench_dragon.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], d['setup']['bvh_build'])"
