#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/mgpu2_p2p.log 2>&1; tail -2 gpurun_out/mgpu2_p2p.log
APOP_B200_ALLREDUCE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/multi_gpu_check.py > gpurun_out/mgpu2_nccl.log 2>&1; tail -2 gpurun_out/mgpu2_nccl.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2_r1c.json 2> gpurun_out/bench_n2_r1c.err; cat gpurun_out/bench_n2_r1c.json | cut -c1-400
