mkdir -p gpurun_out
timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --configs cfg4 --no-cpu --no-e2e --no-fit --no-strong --no-parity --sustain-seconds 0 > gpurun_out/r2c_n2_cfg4.json 2> gpurun_out/r2c_n2_cfg4.err
tail -c 1200 gpurun_out/r2c_n2_cfg4.json; tail -3 gpurun_out/r2c_n2_cfg4.err
