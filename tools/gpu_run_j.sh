set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "both_sweep_kernels or ring_sweep_variants" > gpurun_out/r2j_tests1.log 2>&1
echo "tests1 rc=$?" >> gpurun_out/r2j_tests1.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2j_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2j_gputests.log
timeout 400 python bench.py --steps 50 --no-cpu > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
for w in knots5class_nt31 knots5class_nt127; do
  DOTS_SWEEP_MODE=4 timeout 120 python bench.py --workload $w --steps 200 --no-cpu --no-secondary > gpurun_out/r2j_${w}_m4.json 2> gpurun_out/r2j_${w}_m4.err
  DOTS_SWEEP_MODE=0 timeout 120 python bench.py --workload $w --steps 200 --no-cpu --no-secondary > gpurun_out/r2j_${w}_m0.json 2> gpurun_out/r2j_${w}_m0.err
  timeout 120 python bench.py --workload $w --steps 200 --no-cpu --no-secondary > gpurun_out/r2j_${w}_auto.json 2> gpurun_out/r2j_${w}_auto.err
  DOTS_PERSIST_BLOCKS_PER_SM=2 timeout 120 python bench.py --workload $w --steps 200 --no-cpu --no-secondary > gpurun_out/r2j_${w}_b2.json 2> gpurun_out/r2j_${w}_b2.err
  timeout 120 python bench.py --workload $w --steps 200 --no-cpu --no-secondary --leaf 32 > gpurun_out/r2j_${w}_leaf32.json 2> gpurun_out/r2j_${w}_leaf32.err
  timeout 120 python bench.py --workload $w --steps 200 --no-cpu --no-secondary --leaf 64 > gpurun_out/r2j_${w}_leaf64.json 2> gpurun_out/r2j_${w}_leaf64.err
done
tail -4 gpurun_out/r2j_tests1.log gpurun_out/r2j_gputests.log
