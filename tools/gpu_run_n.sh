# single GPU: ncu launch list of the short bench (after the same command has run without ncu)
set -x
cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary > gpurun_out/r2n2_bench_short.json 2> gpurun_out/r2n2_bench_short.err \
 && timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name "regex:^k_(phi|time|ring|sweep|vertex|tri|kkt|reduce)" -c 800 --csv --log-file gpurun_out/r2n2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary > gpurun_out/r2n2_ncu_bench.log 2>&1
ls -la gpurun_out/r2n2_*
