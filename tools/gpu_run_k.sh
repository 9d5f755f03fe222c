set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "symmetric_time or laplacian_inverse or both_sweep" > gpurun_out/r2k_tests1.log 2>&1
echo "tests1 rc=$?" >> gpurun_out/r2k_tests1.log
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2k_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2k_gputests.log
timeout 400 python bench.py --steps 50 --no-cpu > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
DOTS_TT_SYM=0 timeout 300 python bench.py --steps 50 --no-cpu --no-secondary > gpurun_out/r2k_bench_nosym.json 2> gpurun_out/r2k_bench_nosym.err
tail -4 gpurun_out/r2k_tests1.log gpurun_out/r2k_gputests.log
