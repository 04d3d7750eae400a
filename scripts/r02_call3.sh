mkdir -p gpurun_out/r02
for cfg in "fwd 0 0" "fwd 32 1" "fwd 64 2" "bwd 0 0" "bwd 64 1"; do
  timeout 120 python tests/gpu_chain_probe.py $cfg 2>&1 | tee -a gpurun_out/r02/chain_probe.log
done
for fl in 1 2 4 7; do
  echo "== RLREP_CHAIN_FLAGS=$fl" | tee -a gpurun_out/r02/chain_probe.log
  RLREP_CHAIN_FLAGS=$fl timeout 120 python tests/gpu_chain_probe.py fwd 0 0 2>&1 | head -3 | tee -a gpurun_out/r02/chain_probe.log
done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -s -k "single_update or baseline_config" > gpurun_out/r02/pytest_single.log 2>&1; tail -5 gpurun_out/r02/pytest_single.log; grep -E "single update|b1024|b2048" gpurun_out/r02/pytest_single.log | cut -c1-250 | grep -E "diffsr|speder|b1024"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "checkpoint or batch_size_may or population or batched_select or select_action" 2>&1 | tail -15
