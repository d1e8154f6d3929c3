#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q ) > gpurun_out/pytest_full.txt 2>&1; tail -25 gpurun_out/pytest_full.txt
