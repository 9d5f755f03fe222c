set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "both_sweep_kernels or ring_sweep_variants or headline_laplacian" > gpurun_out/r2b_sweeptests.log 2>&1
echo "sweeptests rc=$?" >> gpurun_out/r2b_sweeptests.log
timeout 900 python tools/sweep_ab.py icosphere7_nt63 4 4:sb=2048 4:sb=2048,stages=4 4:sb=2048,stages=2 4:stages=2 4:stages=4 4:pdl=1,sb=2048 4:tasks=128 4:sb=2048,tasks=128 4:sb=2048,tasks=32 > gpurun_out/r2b_ab_pipe.log 2>&1
DOTS_LIB=$GRAFT_REPO_ROOT/dots_socp_b200/libdots_b200_nopipe.so timeout 600 python tools/sweep_ab.py icosphere7_nt63 4 4:sb=2048 4:stages=2 4:sb=2048,stages=4 > gpurun_out/r2b_ab_nopipe.log 2>&1
timeout 300 python tools/level_times.py icosphere7_nt63 sb=2048 > gpurun_out/r2b_levels_sb2048.log 2>&1
timeout 300 python tools/level_times.py icosphere7_nt63 > gpurun_out/r2b_levels.log 2>&1
timeout 300 python bench.py --steps 50 --no-cpu --no-secondary > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err
DOTS_RING_PDL=1 timeout 300 python bench.py --steps 50 --no-cpu --no-secondary > gpurun_out/r2b_bench_pdl.json 2> gpurun_out/r2b_bench_pdl.err
DOTS_RING_STAGE_BYTES=2048 timeout 300 python bench.py --steps 50 --no-cpu --no-secondary > gpurun_out/r2b_bench_sb2048.json 2> gpurun_out/r2b_bench_sb2048.err
DOTS_RING_STAGE_BYTES=2048 DOTS_RING_PDL=1 timeout 300 python bench.py --steps 50 --no-cpu --no-secondary > gpurun_out/r2b_bench_sb2048_pdl.json 2> gpurun_out/r2b_bench_sb2048_pdl.err
DOTS_SWEEP_MODE=0 timeout 300 python bench.py --steps 50 --no-cpu --no-secondary > gpurun_out/r2b_bench_mode0.json 2> gpurun_out/r2b_bench_mode0.err
tail -3 gpurun_out/r2b_sweeptests.log; cat gpurun_out/r2b_ab_pipe.log | cut -c1-200
