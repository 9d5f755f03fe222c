set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "setup_factorisation or small_root or both_sweep or palm or cscale" > gpurun_out/r2i_tests1.log 2>&1
echo "tests1 rc=$?" >> gpurun_out/r2i_tests1.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2i_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2i_gputests.log
timeout 400 python bench.py --steps 50 --no-cpu > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
timeout 300 python bench.py --steps 50 --no-cpu --no-secondary --leaf 24 > gpurun_out/r2i_bench_leaf24.json 2> gpurun_out/r2i_bench_leaf24.err
timeout 300 python bench.py --steps 50 --no-cpu --no-secondary --leaf 32 > gpurun_out/r2i_bench_leaf32.json 2> gpurun_out/r2i_bench_leaf32.err
DOTS_FACTOR=mixed timeout 300 python bench.py --steps 20 --no-cpu --no-secondary > gpurun_out/r2i_bench_mixed.json 2> gpurun_out/r2i_bench_mixed.err
tail -5 gpurun_out/r2i_tests1.log
