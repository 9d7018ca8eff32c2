set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s_build.log 2>&1
( time timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/s_pytest.log 2>&1
SWEEP_BENCH_ARGS="--steps 4" bash tools/sweep.sh "" "-DPTC_STREAM_STATE=0" > gpurun_out/s_sweep.log 2>&1; cp gpurun_out/sweep.txt gpurun_out/s_sweep.txt
timeout 600 python tools/measure_parity.py gpurun_out/bsdf_error_table.json > gpurun_out/s_parity.log 2>&1
tail -3 gpurun_out/s_pytest.log; cat gpurun_out/s_sweep.txt
