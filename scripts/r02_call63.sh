mkdir -p gpurun_out/r02
timeout 900 python bench.py --workload mulvdrq_population --steps 5 --warmup 3 --repeats 2 --no-cpu-baseline --no-alt-precision > gpurun_out/r02/bench_population_final.json 2> gpurun_out/r02/bench_population_final.err
tail -c 1200 gpurun_out/r02/bench_population_final.json; tail -3 gpurun_out/r02/bench_population_final.err
