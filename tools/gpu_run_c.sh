set -x
cd $GRAFT_REPO_ROOT
DOTS_LIB=$GRAFT_REPO_ROOT/dots_socp_b200/libdots_b200_vec0.so timeout 300 python tools/sweep_ab.py icosphere7_nt63 4:sb=2048 4 > gpurun_out/r2c_ab_vec0.log 2>&1
DOTS_LIB=$GRAFT_REPO_ROOT/dots_socp_b200/libdots_b200_vec0.so timeout 300 python tools/level_times.py icosphere7_nt63 sb=2048 > gpurun_out/r2c_levels_vec0.log 2>&1
timeout 300 python tools/sweep_ab.py icosphere7_nt63 4:sb=2048 > gpurun_out/r2c_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ring_run -s 0 -c 21 -o gpurun_out/r2c_ring_run python tools/sweep_ab.py icosphere7_nt63 4:sb=2048 > gpurun_out/r2c_ncu.log 2>&1
ls -la gpurun_out/ | tail -5
