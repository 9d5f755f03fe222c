cd $GRAFT_REPO_ROOT
timeout 300 python tools/diag_modes.py > gpurun_out/r2z3_diag.log 2>&1
GRAPHS=0 timeout 300 python tools/diag_modes.py > gpurun_out/r2z3_diag_eager.log 2>&1
cat gpurun_out/r2z3_diag.log; echo ----; cat gpurun_out/r2z3_diag_eager.log
