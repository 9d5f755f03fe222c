cd $GRAFT_REPO_ROOT
timeout 300 python tools/profile_e2e.py > gpurun_out/r2ag_profile.log 2>&1
tail -70 gpurun_out/r2ag_profile.log
