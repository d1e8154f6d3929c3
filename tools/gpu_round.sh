#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
timeout 600 python tools/bench_configs.py chains > gpurun_out/configs_r1c.jsonl 2> gpurun_out/configs_r1b.err; cat gpurun_out/configs_r1c.jsonl
timeout 300 python bench.py --family logit --k 24 --rows 50000000 --steps 30 --warmup 5 --no-cpu --no-e2e --no-fit 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['config']['k'], j['ms_per_step'], j['roofline']['achieved'])"
