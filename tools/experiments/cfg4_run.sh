mkdir -p gpurun_out
( time timeout 100 python bench.py --steps 3 --warmup 3 --configs cfg4 --no-cpu --no-e2e --no-fit --no-strong --no-parity --sustain-seconds 0 ) > gpurun_out/r2c_cfg4.json 2> gpurun_out/r2c_cfg4.err
tail -c 2500 gpurun_out/r2c_cfg4.json; tail -3 gpurun_out/r2c_cfg4.err
