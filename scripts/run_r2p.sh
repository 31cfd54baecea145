mkdir -p gpurun_out
L=ibwa_b200/libb200aln.so
AB_STEPS=4 timeout 600 scripts/ab2.sh 10000000 "pin10|$L|--in-flight 1" 2>&1 | tee gpurun_out/r2p_ab.txt
B200ALN_LUT_PIN=11 AB_STEPS=4 timeout 600 scripts/ab2.sh 10000000 "pin11|$L|--in-flight 1" 2>&1 | tee -a gpurun_out/r2p_ab.txt
B200ALN_LUT_PIN=0 AB_STEPS=4 timeout 600 scripts/ab2.sh 10000000 "pin0|$L|--in-flight 1" 2>&1 | tee -a gpurun_out/r2p_ab.txt
B200ALN_L2_FETCH=64 AB_STEPS=4 timeout 600 scripts/ab2.sh 10000000 "fetch64|$L|--in-flight 1" 2>&1 | tee -a gpurun_out/r2p_ab.txt
B200ALN_L2_FETCH=128 AB_STEPS=4 timeout 600 scripts/ab2.sh 10000000 "fetch128|$L|--in-flight 1" 2>&1 | tee -a gpurun_out/r2p_ab.txt
AB_STEPS=4 timeout 600 scripts/ab2.sh 10000000 "a4096|$L|--in-flight 1 --set arena_cap=1024" 2>&1 | tee -a gpurun_out/r2p_ab.txt
