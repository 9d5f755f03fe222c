# single GPU: benches after the PDL / time-chunk change, then ONE ncu --set full capture of the hot kernels (exported to CSV on the box)
set -x
cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --no-cpu > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
echo "bench rc=$?" >> gpurun_out/r2p_bench.err
for pdl in 0 1; do
  DOTS_RING_PDL=$pdl timeout 200 python bench.py --workload knots5class_nt31 --steps 200 --no-cpu --no-secondary > gpurun_out/r2p_knots31_pdl$pdl.json 2> gpurun_out/r2p_knots31_pdl$pdl.err
done
DOTS_RING_PDL=0 timeout 300 python bench.py --steps 50 --no-cpu --no-secondary > gpurun_out/r2p_bench_pdl0.json 2> gpurun_out/r2p_bench_pdl0.err
timeout 300 python tools/ncu_target.py > gpurun_out/r2p_ncu_target_plain.log 2>&1 \
 && timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name "regex:^k_(phi|time|ring|vertex|tri|kkt|reduce)" --launch-skip 49 -c 56 \
    -o /tmp/r2p_hot_kernels python tools/ncu_target.py > gpurun_out/r2p_ncu_target.log 2>&1
ls -la /tmp/r2p_hot_kernels.ncu-rep
ncu -i /tmp/r2p_hot_kernels.ncu-rep --page raw --csv > gpurun_out/r2p_hot_kernels_raw.csv 2> gpurun_out/r2p_export.err
ncu -i /tmp/r2p_hot_kernels.ncu-rep --page details --csv > gpurun_out/r2p_hot_kernels_details.csv 2>> gpurun_out/r2p_export.err
ncu -i /tmp/r2p_hot_kernels.ncu-rep --page source --csv --kernel-name "regex:k_ring_run" -c 3 > gpurun_out/r2p_ring_run_source.csv 2>> gpurun_out/r2p_export.err
ncu -i /tmp/r2p_hot_kernels.ncu-rep --page source --csv --kernel-name "regex:k_kkt_tri" -c 1 > gpurun_out/r2p_kkt_tri_source.csv 2>> gpurun_out/r2p_export.err
du -sh gpurun_out; ls -la gpurun_out | head -20
tail -c 300 gpurun_out/r2p_bench.json
