# single GPU: full GPU suite, bench (with cpu baseline), knots PDL A/B, ncu launch list of the short bench
set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2n_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2n_gputests.log
timeout 600 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
echo "bench rc=$?" >> gpurun_out/r2n_bench.err
for pdl in 0 1; do
  DOTS_RING_PDL=$pdl timeout 200 python bench.py --workload knots5class_nt31 --steps 200 --no-cpu --no-secondary > gpurun_out/r2n_knots31_pdl$pdl.json 2> gpurun_out/r2n_knots31_pdl$pdl.err
  DOTS_RING_PDL=$pdl timeout 200 python bench.py --workload knots5class_nt127 --steps 200 --no-cpu --no-secondary > gpurun_out/r2n_knots127_pdl$pdl.json 2> gpurun_out/r2n_knots127_pdl$pdl.err
done
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary > gpurun_out/r2n_bench_short.json 2> gpurun_out/r2n_bench_short.err \
 && timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name "regex:^k_(phi|time|ring|sweep|vertex|tri|kkt|reduce)" -c 800 --csv --log-file gpurun_out/r2n_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary > gpurun_out/r2n_ncu_bench.log 2>&1
tail -4 gpurun_out/r2n_gputests.log; tail -c 400 gpurun_out/r2n_bench.json
