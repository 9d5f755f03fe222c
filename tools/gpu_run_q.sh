# single GPU: new gather kernel + L2 hint A/B
set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "sweep or ring or headline_laplacian or iterates_match" > gpurun_out/r2q_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2q_tests.log
timeout 400 python tools/sweep_ab.py icosphere7_nt63 4 4:l2hint=1 4 4:l2hint=1 > gpurun_out/r2q_ab.log 2>&1
timeout 300 python bench.py --steps 50 --no-cpu --no-secondary > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
DOTS_RING_L2HINT=1 timeout 300 python bench.py --steps 50 --no-cpu --no-secondary > gpurun_out/r2q_bench_hint.json 2> gpurun_out/r2q_bench_hint.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2q_smoke.log
tail -3 gpurun_out/r2q_tests.log; cat gpurun_out/r2q_ab.log | cut -c1-200; tail -2 gpurun_out/r2q_smoke.log
