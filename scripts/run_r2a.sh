python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.txt 2>&1; tail -5 gpurun_out/r2a_pytest.txt
scripts/ab2.sh 10000000 "base|ab/base.so|" "jump|ibwa_b200/libb200aln.so|--set order=0" "jumpord|ibwa_b200/libb200aln.so|" 2>&1 | tee gpurun_out/r2a_ab.txt
scripts/ab2.sh 1250000 "base|ab/base.so|" "jump|ibwa_b200/libb200aln.so|--set order=0" "jumpord|ibwa_b200/libb200aln.so|" 2>&1 | tee -a gpurun_out/r2a_ab.txt
scripts/ab2.sh 262144 "base|ab/base.so|" "jumpord|ibwa_b200/libb200aln.so|" 2>&1 | tee -a gpurun_out/r2a_ab.txt
