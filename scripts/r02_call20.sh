mkdir -p gpurun_out/r02
for w in fwd bwd; do
  timeout 120 python tests/gpu_chain_probe.py $w > gpurun_out/r02/probe_entry_$w.log 2>&1
  grep "chain:\|kernel entry" gpurun_out/r02/probe_entry_$w.log
done
