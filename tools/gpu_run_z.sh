set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "kkt1 or kkt or drop_in or plugin" > gpurun_out/r2z_tests1.log 2>&1
echo "rc=$?" >> gpurun_out/r2z_tests1.log
tail -25 gpurun_out/r2z_tests1.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2z_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2z_gputests.log
for i in 1 2; do timeout 300 python bench.py --steps 50 --no-cpu --no-secondary >> gpurun_out/r2z_bench.json 2>> gpurun_out/r2z_bench.err; done
tail -4 gpurun_out/r2z_gputests.log; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2z_bench.json; grep -o '"time_to_tol_s": [0-9.]*' gpurun_out/r2z_bench.json
