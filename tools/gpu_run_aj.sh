cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tma_triangle or plane or kkt1 or refplane" > gpurun_out/r2aj_tests1.log 2>&1
echo "rc=$?" >> gpurun_out/r2aj_tests1.log
tail -15 gpurun_out/r2aj_tests1.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2aj_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2aj_gputests.log
for i in 1 2; do timeout 300 python bench.py --steps 50 --no-cpu --no-secondary >> gpurun_out/r2aj_bench.json 2>> gpurun_out/r2aj_bench.err; done
timeout 200 python tools/time_iter.py plane100 31 > gpurun_out/r2aj_plane100.json 2>&1
DOTS_TRI_PLAIN=1 timeout 200 python tools/time_iter.py plane100 31 > gpurun_out/r2aj_plane100_plain.json 2>&1
tail -4 gpurun_out/r2aj_gputests.log
