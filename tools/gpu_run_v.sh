set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tools/e2e_parts.py > gpurun_out/r2v_parts.json 2> gpurun_out/r2v_parts.err
timeout 300 python bench.py --steps 50 --no-cpu --no-secondary > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
cat gpurun_out/r2v_parts.json; tail -3 gpurun_out/r2v_parts.err; grep -o '"setup_breakdown_s": {[^}]*}' gpurun_out/r2v_bench.json
