mkdir -p gpurun_out
L=gpurun_out/wide_run3.log; : > $L
echo "cols RH=32 from 49" >> $L; APOP_B200_COOP_MAX_K=48 timeout 200 python tools/kbench.py wide:49 wide:56 wide:96 2>&1 | tail -5 >> $L
echo "coop to 256, defaults beyond" >> $L; timeout 200 python tools/kbench.py wide:49 wide:56 wide:64 wide:96 wide:257 wide:640 wide:768 2>&1 | tail -8 >> $L
cat $L
