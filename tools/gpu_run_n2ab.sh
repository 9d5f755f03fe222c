# 2 GPUs: A/B of the programmatic chaining on the sharded path (A B A B)
set -x
cd $GRAFT_REPO_ROOT
for i in 1 2; do
for pdl in 0 1; do
  DOTS_RING_PDL=$pdl timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$pdl bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu --no-secondary >> gpurun_out/r2t_n2_pdl$pdl.json 2>> gpurun_out/r2t_n2_pdl$pdl.err
done
done
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2t_n2_pdl0.json gpurun_out/r2t_n2_pdl1.json
