# single GPU: full GPU suite after the PDL / time-chunk change, benches, then ONE ncu --set full capture of the hot kernels
set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2o_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2o_gputests.log
timeout 600 python bench.py --no-cpu > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
echo "bench rc=$?" >> gpurun_out/r2o_bench.err
for pdl in 0 1; do
  DOTS_RING_PDL=$pdl timeout 200 python bench.py --workload knots5class_nt31 --steps 200 --no-cpu --no-secondary > gpurun_out/r2o_knots31_pdl$pdl.json 2> gpurun_out/r2o_knots31_pdl$pdl.err
done
DOTS_RING_PDL=0 timeout 300 python bench.py --steps 50 --no-cpu --no-secondary > gpurun_out/r2o_bench_pdl0.json 2> gpurun_out/r2o_bench_pdl0.err
timeout 300 python tools/ncu_target.py > gpurun_out/r2o_ncu_target_plain.log 2>&1 \
 && timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name "regex:^k_(phi|time|ring|vertex|tri|kkt|reduce)" -c 70 \
    -o gpurun_out/r2o_hot_kernels python tools/ncu_target.py > gpurun_out/r2o_ncu_target.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -4 gpurun_out/r2o_gputests.log; tail -c 300 gpurun_out/r2o_bench.json
