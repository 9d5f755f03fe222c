set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "mode_groups or nt159 or nt299 or (iterates_match_oracle and (159 or 299))" > gpurun_out/r2y_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2y_tests.log
tail -30 gpurun_out/r2y_tests.log
