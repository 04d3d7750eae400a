mkdir -p gpurun_out/final
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q > gpurun_out/final/pytest_sharded_n2.log 2>&1; tail -3 gpurun_out/final/pytest_sharded_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/final/r02_bench_default_n2.json 2> gpurun_out/final/bench_n2.err
tail -c 600 gpurun_out/final/r02_bench_default_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/final/r02_bench_reference_n2.json 2> gpurun_out/final/bench_ref_n2.err; tail -c 300 gpurun_out/final/r02_bench_reference_n2.json
