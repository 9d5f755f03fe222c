set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tools/setup_probe.py > gpurun_out/r2x_probe.json 2> gpurun_out/r2x_probe.err
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "specialised_kkt" > gpurun_out/r2x_tests.log 2>&1
cat gpurun_out/r2x_probe.json; tail -3 gpurun_out/r2x_probe.err; tail -3 gpurun_out/r2x_tests.log
