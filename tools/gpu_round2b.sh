#!/bin/bash
# Round-2 closing run on one GPU: GPU tests, smoke, both bench arms, launch list, ncu of the
# multinomial-logit streaming kernel and of the information kernel.  Outputs under gpurun_out/.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b_pytest.log 2>&1; tail -5 gpurun_out/r2b_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; tail -c 1500 gpurun_out/r2b_bench_n1.json; tail -4 gpurun_out/r2b_bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err; tail -c 600 gpurun_out/r2b_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-fit --no-strong --no-configs --no-parity --sustain-seconds 0 > gpurun_out/r2b_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logit_stream -c 1 -s 3 -o gpurun_out/r2b_prof_logit_stream -f python tools/bench_configs.py logit8 > gpurun_out/r2b_ncu_logit.log 2>&1; tail -2 gpurun_out/r2b_ncu_logit.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logit_info -c 1 -s 8 -o gpurun_out/r2b_prof_logit_info -f python tools/bench_configs.py logit8_info > gpurun_out/r2b_ncu_info.log 2>&1; tail -2 gpurun_out/r2b_ncu_info.log
ls -la gpurun_out
