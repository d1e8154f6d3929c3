#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
APOP_B200_LOGIT_VARIANT=mma timeout 600 python -m pytest tests/test_gpu_parity_glm.py -m gpu -x -q -k "logit" 2>&1 | tail -1
