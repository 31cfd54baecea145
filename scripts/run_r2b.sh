python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.txt 2>&1; tail -5 gpurun_out/r2b_pytest.txt
L=ibwa_b200/libb200aln.so
scripts/ab2.sh 10000000 "q32ord|$L|--set q8=0" "q8ord|$L|" "q8noord|$L|--set order=0" 2>&1 | tee gpurun_out/r2b_ab.txt
scripts/ab2.sh 1250000 "q32ord|$L|--set q8=0" "q8ord|$L|" 2>&1 | tee -a gpurun_out/r2b_ab.txt
ncu --set full --clock-control none --import-source on -k regex:k_search -c 1 -o gpurun_out/r2b_search python bench.py --reads 4000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_width -c 1 -o gpurun_out/r2b_width python bench.py --reads 4000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_ncu_w.log 2>&1
ls -la gpurun_out/*.ncu-rep
