mkdir -p gpurun_out
( time timeout 85 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; tail -c 400 gpurun_out/r2d_bench_n1.json; tail -4 gpurun_out/r2d_bench_n1.err
