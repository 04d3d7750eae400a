mkdir -p gpurun_out/r02
timeout 600 python tests/gpu_pixel_timeline.py ldiffsr_pixels_b256 150 > gpurun_out/r02/ldiffsr_timeline.log 2>&1
tail -5 gpurun_out/r02/ldiffsr_timeline.log
