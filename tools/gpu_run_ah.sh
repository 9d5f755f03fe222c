cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2ah_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2ah_gputests.log
timeout 300 python tools/e2e_parts.py > gpurun_out/r2ah_parts.json 2> gpurun_out/r2ah_parts.err
for i in 1 2; do timeout 300 python bench.py --steps 50 --no-cpu --no-secondary >> gpurun_out/r2ah_bench.json 2>> gpurun_out/r2ah_bench.err; done
tail -4 gpurun_out/r2ah_gputests.log; cat gpurun_out/r2ah_parts.json; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2ah_bench.json; grep -o '"time_to_tol_s": [0-9.]*' gpurun_out/r2ah_bench.json
