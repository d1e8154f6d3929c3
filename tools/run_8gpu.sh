#!/bin/bash
# 8-GPU run behind profiles/r01c_bench_n8*.json: parity check, weak scaling (100M rows per GPU), the fixed 100M-row problem
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/multi_gpu_check.py > gpurun_out/mgpu${N}_p2p.log 2>&1; tail -1 gpurun_out/mgpu${N}_p2p.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n${N}_r1c.json 2> gpurun_out/bench_n${N}_r1c.err; cut -c1-260 gpurun_out/bench_n${N}_r1c.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --rows $((100000000 / N)) --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_n${N}_strong_r1c.json 2> gpurun_out/bench_n${N}_strong_r1c.err; cut -c1-260 gpurun_out/bench_n${N}_strong_r1c.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29524 bench.py --impl reference --gpus $N --steps 3 --warmup 1 | cut -c1-200
