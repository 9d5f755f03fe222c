cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tma_triangle or plane or kkt1 or refplane or headline_laplacian" > gpurun_out/r2ak_tests1.log 2>&1
echo "rc=$?" >> gpurun_out/r2ak_tests1.log
for i in 1 2; do timeout 300 python bench.py --steps 50 --no-cpu --no-secondary >> gpurun_out/r2ak_bench.json 2>> gpurun_out/r2ak_bench.err; done
timeout 200 python tools/time_iter.py plane100 31 > gpurun_out/r2ak_plane100.json 2>&1
tail -3 gpurun_out/r2ak_tests1.log
