#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_glm.py -m gpu -x -q -k "logit" > gpurun_out/pytest_logit.txt 2>&1; tail -3 gpurun_out/pytest_logit.txt
timeout 300 python tools/bench_configs.py logit8 > gpurun_out/logit8_hybrid.json 2> gpurun_out/logit8_hybrid.err
cat gpurun_out/logit8_hybrid.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logit_hybrid -c 1 -s 3 -o gpurun_out/prof_logit8_hyb4 -f python tools/bench_configs.py logit8 > gpurun_out/ncu_hyb4.log 2>&1
tail -2 gpurun_out/ncu_hyb4.log
