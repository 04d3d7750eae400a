mkdir -p gpurun_out/r02
timeout 900 ncu --set full --clock-control none --import-source on -k regex:out_conv_ -c 3 -o gpurun_out/r02/ncu_outconv -f python bench.py --workload ldiffsr_pixels_b256 --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline --no-alt-precision > gpurun_out/r02/ncu_outconv.log 2>&1
tail -3 gpurun_out/r02/ncu_outconv.log
ls -la gpurun_out/r02/ncu_outconv.ncu-rep
