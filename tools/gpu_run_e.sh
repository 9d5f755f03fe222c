set -x
cd $GRAFT_REPO_ROOT
DOTS_LIB=$GRAFT_REPO_ROOT/dots_socp_b200/libdots_b200_vec0.so timeout 300 python tools/sweep_ab.py icosphere7_nt63 4:stages=2,pdl=1 4:stages=2 > gpurun_out/r2e_ab_vec0.log 2>&1
DOTS_LIB=$GRAFT_REPO_ROOT/dots_socp_b200/libdots_b200_vec0.so timeout 300 python tools/level_times.py icosphere7_nt63 stages=2 > gpurun_out/r2e_levels_vec0.log 2>&1
timeout 300 python tools/level_times.py icosphere7_nt63 stages=2 > gpurun_out/r2e_levels_st2.log 2>&1
timeout 900 python tools/sweep_ab.py icosphere7_nt63 4:stages=2,pdl=1 4:stages=2,pdl=1,tasks=48 4:stages=2,pdl=1,tasks=96 4:stages=2,pdl=1,tasks=128 4:stages=2,pdl=1,tmax=48 4:stages=2,pdl=1,tmax=192,tasks=32 4:stages=2,pdl=1,split=64 4:stages=2,pdl=1,split=128 4:stages=2,pdl=1,split=48,wpr=4 4:stages=2,pdl=1,tmin=32 > gpurun_out/r2e_ab.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2e_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2e_gputests.log
DOTS_RING_STAGES=2 DOTS_RING_PDL=1 timeout 400 python bench.py --steps 50 > gpurun_out/r2e_bench_full.json 2> gpurun_out/r2e_bench_full.err
tail -3 gpurun_out/r2e_gputests.log
