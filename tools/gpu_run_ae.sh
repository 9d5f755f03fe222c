cd $GRAFT_REPO_ROOT
timeout 600 python tools/sweep_ab.py icosphere7_nt63 4:tasks=64 4:tasks=48 4:tasks=72 4:tasks=96 4:tasks=144 4:tasks=64 4:tasks=72,tmax=64 4:tasks=36,tmax=96 > gpurun_out/r2ae_ab.log 2>&1
cut -c1-140 gpurun_out/r2ae_ab.log
