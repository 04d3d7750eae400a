mkdir -p gpurun_out/r02
for f in 0 1; do
  RLREP_HALO_FLAGS=$f timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -s -k "64" > gpurun_out/r02/pytest_halo_f$f.log 2>&1
  echo "flags=$f"; grep "B=64\|passed\|failed" gpurun_out/r02/pytest_halo_f$f.log | cut -c1-260
done
