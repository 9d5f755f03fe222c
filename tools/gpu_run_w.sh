set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2w_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2w_gputests.log
timeout 300 python tools/e2e_parts.py > gpurun_out/r2w_parts.json 2> gpurun_out/r2w_parts.err
timeout 300 python bench.py --steps 50 --no-cpu --no-secondary > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
tail -3 gpurun_out/r2w_gputests.log; cat gpurun_out/r2w_parts.json; grep -o '"setup_breakdown_s": {[^}]*}' gpurun_out/r2w_bench.json; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2w_bench.json
