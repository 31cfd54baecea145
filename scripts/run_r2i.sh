# parking ring with in-kernel refill + blocking host waits: parity, then the small-launch matrix
mkdir -p gpurun_out
L=ibwa_b200/libb200aln.so
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_checked.py -m gpu -x -q > gpurun_out/r2i_pytest.txt 2>&1; tail -4 gpurun_out/r2i_pytest.txt
AB_STEPS=24 timeout 600 scripts/ab2.sh 1250000 "ring_k6|$L|--in-flight 6" "ring_k4|$L|--in-flight 4" "ring_k3s|$L|--in-flight 3 --set susp=16" "ring_k8|$L|--in-flight 8" 2>&1 | tee gpurun_out/r2i_ab.txt
AB_STEPS=18 timeout 600 scripts/ab2.sh 2500000 "ring_k6|$L|--in-flight 6" "ring_k4|$L|--in-flight 4" 2>&1 | tee -a gpurun_out/r2i_ab.txt
AB_STEPS=36 timeout 600 scripts/ab2.sh 262144 "ring_k6|$L|--in-flight 6" "ring_k8|$L|--in-flight 8" 2>&1 | tee -a gpurun_out/r2i_ab.txt
AB_STEPS=6 timeout 600 scripts/ab2.sh 10000000 "ring_s16|$L|--set susp=16" "ring_auto|$L|" 2>&1 | tee -a gpurun_out/r2i_ab.txt
B200ALN_TIMELINE=1 timeout 300 python bench.py --reads 1250000 --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_tl.json 2> gpurun_out/r2i_tl.err; grep timeline gpurun_out/r2i_tl.err | tail -14
