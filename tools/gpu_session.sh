set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s1_build.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s1_smi.txt
( time timeout 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^Parsing\|^Done\|^Reserving\|EnvironmentLight" | tail -150 ) > gpurun_out/s1_pytest.log 2>&1
timeout 600 python tools/measure_parity.py gpurun_out/bsdf_error_table.json > gpurun_out/s1_parity.log 2>&1
timeout 120 tools/l2_bandwidth > gpurun_out/l2_peak.json 2> gpurun_out/s1_l2.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/s1_bench_ref.log 2>&1
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/s1_bench.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none -k regex:traverseKernel --csv --log-file gpurun_out/wave_metrics.csv python tools/profile_wave.py dragon 64 gpurun_out/wave_counts.json > gpurun_out/s1_ncu.log 2>&1
tail -3 gpurun_out/s1_pytest.log; tail -2 gpurun_out/s1_bench.log | cut -c1-600
