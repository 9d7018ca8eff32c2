# GPU-box session (run under gpurun): tests, parity table, bench, ncu launch list + per-launch metrics of one wave
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s_build.log 2>&1
( time timeout 1800 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^Parsing\|^Done\|^Reserving\|EnvironmentLight" | tail -250 ) > gpurun_out/s_pytest.log 2>&1
timeout 600 python tools/measure_parity.py gpurun_out/bsdf_error_table.json > gpurun_out/s_parity.log 2>&1
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/s_bench.log 2>&1
if [ "${SESSION_SWEEP:-}" != "" ]; then SWEEP_BENCH_ARGS="--steps 4" bash tools/sweep.sh "" "$SESSION_SWEEP" > gpurun_out/s_sweep.log 2>&1; cp gpurun_out/sweep.txt gpurun_out/s_sweep.txt; fi
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/wave_metrics_all.csv python tools/profile_wave.py dragon 64 gpurun_out/wave_counts.json > gpurun_out/s_ncu.log 2>&1
tail -3 gpurun_out/s_pytest.log; tail -1 gpurun_out/s_bench.log | cut -c1-300
