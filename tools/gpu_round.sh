#!/bin/bash
# The consolidated one-GPU measurement run behind profiles/r01c_* (run with gpurun from the repo root):
# both bench arms, the other BASELINE configurations, the launch list and the ncu captures.
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 3000 gpurun_out/bench_r1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r1.json 2> gpurun_out/bench_ref_r1.err; cat gpurun_out/bench_ref_r1.json
timeout 1500 python tools/bench_configs.py ingest prep logit8 bootstrap chains poisson newton bootstrap_cov metropolis config1 > gpurun_out/configs_r1d.jsonl 2> gpurun_out/configs_r1d.err; cat gpurun_out/configs_r1d.jsonl
# launch list of the bench command (ncu, cold-cache per-launch times: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-fit > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:glm_fused -c 1 -s 3 -o gpurun_out/prof_probit_r1c -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-fit > gpurun_out/ncu_probit.log 2>&1; tail -1 gpurun_out/ncu_probit.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:logit_hybrid -c 1 -s 3 -o gpurun_out/prof_logit8_r1c -f python tools/bench_configs.py logit8 > gpurun_out/ncu_logit.log 2>&1; tail -1 gpurun_out/ncu_logit.log
