set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "kkt1" > gpurun_out/r2z2_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2z2_tests.log
tail -25 gpurun_out/r2z2_tests.log
