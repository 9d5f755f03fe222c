# inside gpurun --gpus 8: world-8 test on the final code, benches at N = 4 and 8
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-r2ac}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_smi_n8.log 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "ico3_nt31 and 8" > gpurun_out/${TAG}_tests_n8.log 2>&1
echo "rc=$?" >> gpurun_out/${TAG}_tests_n8.log
for k in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 2950$k bench.py --gpus $k --steps 50 --warmup 5 --no-secondary > gpurun_out/${TAG}_bench_n$k.json 2> gpurun_out/${TAG}_bench_n$k.err
done
tail -3 gpurun_out/${TAG}_tests_n8.log; tail -c 400 gpurun_out/${TAG}_bench_n8.json
