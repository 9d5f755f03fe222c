set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "both_sweep_kernels or ring_sweep_variants or headline_laplacian or iterates_match" > gpurun_out/r2d_sweeptests.log 2>&1
echo "sweeptests rc=$?" >> gpurun_out/r2d_sweeptests.log
timeout 900 python tools/sweep_ab.py icosphere7_nt63 0 4 4:sb=2048 4:sb=2048,stages=4 4:sb=2048,stages=2 4:stages=2 4:pdl=1,sb=2048 4:pdl=1,stages=2 4:sb=2048,tasks=128 4:sb=2048,split=64 > gpurun_out/r2d_ab.log 2>&1
timeout 300 python tools/level_times.py icosphere7_nt63 sb=2048 > gpurun_out/r2d_levels_sb2048.log 2>&1
DOTS_RING_STAGE_BYTES=2048 DOTS_RING_PDL=1 timeout 300 python bench.py --steps 50 --no-cpu --no-secondary > gpurun_out/r2d_bench_sb2048_pdl.json 2> gpurun_out/r2d_bench_sb2048_pdl.err
DOTS_RING_STAGES=2 DOTS_RING_PDL=1 timeout 300 python bench.py --steps 50 --no-cpu --no-secondary > gpurun_out/r2d_bench_st2_pdl.json 2> gpurun_out/r2d_bench_st2_pdl.err
timeout 300 python tools/sweep_ab.py icosphere7_nt63 4:sb=2048 > gpurun_out/r2d_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ring_run -s 0 -c 21 -o gpurun_out/r2d_ring_run python tools/sweep_ab.py icosphere7_nt63 4:sb=2048 > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_sweeptests.log; cut -c1-150 gpurun_out/r2d_ab.log
