set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s_build.log 2>&1
( time timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^Parsing\|^Done\|^Reserving\|EnvironmentLight" | tail -150 ) > gpurun_out/s_pytest.log 2>&1
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/s_bench.log 2>&1
tail -4 gpurun_out/s_pytest.log; tail -1 gpurun_out/s_bench.log | cut -c1-600
