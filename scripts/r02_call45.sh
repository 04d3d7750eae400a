mkdir -p gpurun_out/r02
timeout 300 python tests/gpu_timeline.py vlsac_hum_b1024 > gpurun_out/r02/timeline_vlsac.csv 2> gpurun_out/r02/timeline_vlsac.err
wc -l gpurun_out/r02/timeline_vlsac.csv
