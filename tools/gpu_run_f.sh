set -x
cd $GRAFT_REPO_ROOT
for args in "knots5class_nt63_c0 16 4" "knots5class_nt63_c0 16 0" "knots5class_nt63_c0 24 4" "knots5class_nt127_c0 16 4" "knots5class_nt127_c0 16 0"; do
  timeout 300 python tools/diag_iters.py $args >> gpurun_out/r2f_diag.log 2>&1
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cscale or kkt_and_objective or stepwise" > gpurun_out/r2f_cscale.log 2>&1
echo "rc=$?" >> gpurun_out/r2f_cscale.log
grep -v "^$" gpurun_out/r2f_diag.log | grep -v Iter | tail -60
tail -15 gpurun_out/r2f_cscale.log
