#!/bin/bash
# Round-2 final check on one GPU: GPU tests, smoke, bench.
mkdir -p gpurun_out
( time timeout 200 python -m pytest tests -m gpu -x -q --timeout 60 ) > gpurun_out/r2c_pytest.log 2>&1; tail -5 gpurun_out/r2c_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 110 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; tail -c 300 gpurun_out/r2c_bench_n1.json; tail -4 gpurun_out/r2c_bench_n1.err
