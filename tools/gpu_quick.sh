# quick GPU-box session: build, the GPU test suite, prebuilt variant sweep, optional source-level ncu capture
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s_build.log 2>&1
if [ "${QUICK_TESTS:-1}" = "1" ]; then ( time timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 ) > gpurun_out/s_pytest.log 2>&1; tail -5 gpurun_out/s_pytest.log; fi
if [ -d variants ]; then bash tools/sweep_prebuilt.sh ${QUICK_VARIANTS:-}; fi
if [ "${QUICK_NCU:-}" != "" ]; then bash tools/ncu_source.sh $QUICK_NCU; fi
