mkdir -p gpurun_out/r02
for d in 0 1; do
  for w in fwd bwd; do
    RLREP_CHAIN_DIST=$d RLREP_CHAIN_VERBOSE=1 timeout 120 python tests/gpu_chain_probe.py $w > gpurun_out/r02/probe_dist${d}_$w.log 2>&1
  done
done
tail -12 gpurun_out/r02/probe_dist1_fwd.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "device_noise or row_operations" > gpurun_out/r02/pytest_dn.log 2>&1; tail -4 gpurun_out/r02/pytest_dn.log
