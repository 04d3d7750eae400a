mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02/pytest_all4.log 2>&1; tail -12 gpurun_out/r02/pytest_all4.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
