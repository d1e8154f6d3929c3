mkdir -p gpurun_out
timeout 70 ncu --set full --clock-control none --import-source on -k regex:glm_cols -c 1 -s 5 -o gpurun_out/r2c_prof_wide1024 -f python tools/kbench.py wide:1024 > gpurun_out/r2c_ncu_wide1024.log 2>&1; tail -2 gpurun_out/r2c_ncu_wide1024.log
timeout 70 ncu --set full --clock-control none --import-source on -k regex:glm_cols -c 1 -s 5 -o gpurun_out/r2c_prof_wide256 -f python tools/kbench.py wide:256 > gpurun_out/r2c_ncu_wide256.log 2>&1; tail -2 gpurun_out/r2c_ncu_wide256.log
