# usage: bash tools/gpu_run_multi.sh N [tag]      (inside gpurun --gpus N): world-size tests up to N, then benches at 2..N
set -x
cd $GRAFT_REPO_ROOT
N=$1
TAG=${2:-r2m}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_smi_n$N.log 2>&1
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/${TAG}_tests_n$N.log 2>&1
else
  timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "ico3_nt31" > gpurun_out/${TAG}_tests_n$N.log 2>&1
fi
echo "rc=$?" >> gpurun_out/${TAG}_tests_n$N.log
for k in 2 4 8; do
  if [ $k -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 2950$k bench.py --gpus $k --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_n$k.json 2> gpurun_out/${TAG}_bench_n$k.err
  fi
done
tail -3 gpurun_out/${TAG}_tests_n$N.log; tail -c 600 gpurun_out/${TAG}_bench_n$N.json
