cd $GRAFT_REPO_ROOT
timeout 300 python tools/level_times.py icosphere7_nt63 > gpurun_out/r2ad_levels.log 2>&1
timeout 300 python tools/level_times.py icosphere7_nt63 pdl=0 > gpurun_out/r2ad_levels_nopdl.log 2>&1
timeout 200 python tools/time_iter.py knot 255 > gpurun_out/r2ad_knot255.json 2>&1
timeout 200 python tools/time_iter.py knot 127 > gpurun_out/r2ad_knot127.json 2>&1
timeout 300 python tools/time_iter.py icosphere6 255 20 > gpurun_out/r2ad_ico6_255.json 2>&1
cat gpurun_out/r2ad_levels.log; cat gpurun_out/r2ad_knot255.json gpurun_out/r2ad_knot127.json gpurun_out/r2ad_ico6_255.json
