mkdir -p gpurun_out/r02
for w in spedersac_hc_b256 diffsrsac_hc_b256 sac_hc_b256; do
  timeout 300 python tests/gpu_timeline.py $w > gpurun_out/r02/timeline_$w.csv 2> gpurun_out/r02/timeline_$w.err
done
wc -l gpurun_out/r02/timeline_*.csv
