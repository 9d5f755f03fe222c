set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2a_smi.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "both_sweep_kernels or ring_sweep_variants" > gpurun_out/r2a_sweeptests.log 2>&1
echo "sweeptests rc=$?" >> gpurun_out/r2a_sweeptests.log
timeout 900 python tools/sweep_ab.py icosphere7_nt63 0 4 4:pdl=1 4:stages=2 4:split=48 4:split=256 4:tasks=24 4:tmax=32 > gpurun_out/r2a_ab.log 2>&1
echo "ab rc=$?" >> gpurun_out/r2a_ab.log
timeout 300 python tools/level_times.py icosphere7_nt63 > gpurun_out/r2a_levels.log 2>&1
timeout 300 python tools/level_times.py icosphere7_nt63 pdl=1 > gpurun_out/r2a_levels_pdl.log 2>&1
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2a_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2a_gputests.log
tail -5 gpurun_out/r2a_sweeptests.log gpurun_out/r2a_ab.log gpurun_out/r2a_gputests.log
