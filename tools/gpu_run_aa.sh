cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "drop_in or plugin or replication" > gpurun_out/r2aa_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2aa_tests.log
tail -3 gpurun_out/r2aa_tests.log
for i in 1 2 3; do timeout 300 python bench.py --steps 50 --no-cpu --no-secondary >> gpurun_out/r2aa_bench.json 2>> gpurun_out/r2aa_bench.err; done
grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2aa_bench.json; grep -o '"time_to_tol_s": [0-9.]*' gpurun_out/r2aa_bench.json; grep -o '"timed_s": [0-9.]*' gpurun_out/r2aa_bench.json
