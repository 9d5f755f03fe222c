# usage: bash tools/gpu_run_multi.sh N      (inside gpurun --gpus N)
set -x
cd $GRAFT_REPO_ROOT
N=$1
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2m_smi_n$N.log 2>&1
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2m_tests_n$N.log 2>&1
  echo "rc=$?" >> gpurun_out/r2m_tests_n$N.log
fi
if [ "$N" = "8" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "ico3_nt31" > gpurun_out/r2m_tests_n$N.log 2>&1
  echo "rc=$?" >> gpurun_out/r2m_tests_n$N.log
  for k in 2 4; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 2950$k bench.py --gpus $k --steps 50 --warmup 5 > gpurun_out/r2m_bench_n$k.json 2> gpurun_out/r2m_bench_n$k.err
  done
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2m_bench_n$N.json 2> gpurun_out/r2m_bench_n$N.err
tail -3 gpurun_out/r2m_tests_n$N.log; tail -c 600 gpurun_out/r2m_bench_n$N.json
