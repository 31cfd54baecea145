mkdir -p gpurun_out
L=ibwa_b200/libb200aln.so
AB_STEPS=5 timeout 900 scripts/ab2.sh 10000000 "base|ab/base.so|" "earlyq|$L|" "base2|ab/base.so|" "earlyq2|$L|" 2>&1 | tee gpurun_out/r2m_ab.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or parked or random" > gpurun_out/r2m_pytest.txt 2>&1; tail -3 gpurun_out/r2m_pytest.txt
