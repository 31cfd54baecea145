# baseline of HEAD at session start: gpu tests, headline bench, small-launch A/B
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.txt 2>&1; tail -5 gpurun_out/r2c_pytest.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 3000 gpurun_out/r2c_bench.json
L=ibwa_b200/libb200aln.so
scripts/ab2.sh 1250000 "ord|$L|" "noord|$L|--set order=0" 2>&1 | tee gpurun_out/r2c_ab.txt
scripts/ab2.sh 262144 "ord|$L|" 2>&1 | tee -a gpurun_out/r2c_ab.txt
