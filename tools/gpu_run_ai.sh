cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "kkt or fixture" > gpurun_out/r2ai_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2ai_tests.log
timeout 300 python tools/e2e_parts.py > gpurun_out/r2ai_parts.json 2> gpurun_out/r2ai_parts.err
tail -3 gpurun_out/r2ai_tests.log; cat gpurun_out/r2ai_parts.json
