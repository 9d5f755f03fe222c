# single GPU: setup-path changes (native front maps, trimmed driver): full suite + two benches for the setup breakdown
set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2u_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2u_gputests.log
for i in 1 2; do
timeout 300 python bench.py --steps 50 --no-cpu --no-secondary >> gpurun_out/r2u_bench.json 2>> gpurun_out/r2u_bench.err
done
tail -3 gpurun_out/r2u_gputests.log; grep -o '"setup_breakdown_s": {[^}]*}' gpurun_out/r2u_bench.json
