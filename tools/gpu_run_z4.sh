cd $GRAFT_REPO_ROOT
timeout 300 python tools/diag_modes.py > gpurun_out/r2z4_diag.log 2>&1
cat gpurun_out/r2z4_diag.log | grep b_mid
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2z4_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2z4_gputests.log
tail -5 gpurun_out/r2z4_gputests.log
timeout 300 python bench.py --steps 50 --no-cpu --no-secondary > gpurun_out/r2z4_bench.json 2> gpurun_out/r2z4_bench.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2z4_bench.json; grep -o '"status": "[a-zA-Z]*"' gpurun_out/r2z4_bench.json;  grep -o '"checksums_vs_single_gpu_record": [0-9.e-]*' gpurun_out/r2z4_bench.json
