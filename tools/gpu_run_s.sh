# single GPU: reference arm on the box, L2-hint A/B for the triangle kernel, full suite, final bench line
set -x
cd $GRAFT_REPO_ROOT
free -g > gpurun_out/r2s_mem.log; nproc >> gpurun_out/r2s_mem.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s_reference_arm.json 2> gpurun_out/r2s_reference_arm.err
echo "ref rc=$?" >> gpurun_out/r2s_reference_arm.err
for h in 1 3 1 3; do
  DOTS_RING_L2HINT=$h timeout 300 python bench.py --steps 50 --no-cpu --no-secondary >> gpurun_out/r2s_hint$h.json 2>> gpurun_out/r2s_hint$h.err
done
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2s_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2s_gputests.log
timeout 900 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
echo "bench rc=$?" >> gpurun_out/r2s_bench.err
tail -3 gpurun_out/r2s_gputests.log; tail -c 300 gpurun_out/r2s_bench.json; tail -c 600 gpurun_out/r2s_reference_arm.json
