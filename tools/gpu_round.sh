#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_glm.py -m gpu -x -q -k "logit" > gpurun_out/pytest_logit.txt 2>&1; tail -3 gpurun_out/pytest_logit.txt
for st in 2 2; do
APOP_B200_LOGIT_STAGES=$st timeout 300 python tools/bench_configs.py logit8 > gpurun_out/logit8_hybrid_st$st.json 2> gpurun_out/logit8_hybrid.err
cat gpurun_out/logit8_hybrid_st$st.json
done
