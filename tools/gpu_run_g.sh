set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "both_sweep_kernels or ring_sweep_variants or headline_laplacian or iterates_match" > gpurun_out/r2g_sweeptests.log 2>&1
echo "sweeptests rc=$?" >> gpurun_out/r2g_sweeptests.log
timeout 900 python tools/sweep_ab.py icosphere7_nt63 0 4 4:pdl=0 4:stages=3 4:tasks=128 4:tasks=256,tmin=8 4:tmax=48 4:split=96 4:sb=2048,stages=3 4:stages=3,tasks=128 > gpurun_out/r2g_ab.log 2>&1
timeout 300 python tools/level_times.py icosphere7_nt63 > gpurun_out/r2g_levels.log 2>&1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2g_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2g_gputests.log
timeout 400 python bench.py --steps 50 --no-cpu > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -3 gpurun_out/r2g_sweeptests.log gpurun_out/r2g_gputests.log; cut -c1-140 gpurun_out/r2g_ab.log
