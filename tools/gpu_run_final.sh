# final single-GPU verification of the round: full GPU suite, smoke, the default bench line
set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2g_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2g_gputests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2g_smoke.log
timeout 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "bench rc=$?" >> gpurun_out/r2g_bench.err
tail -3 gpurun_out/r2g_gputests.log; tail -2 gpurun_out/r2g_smoke.log; tail -c 300 gpurun_out/r2g_bench.json
