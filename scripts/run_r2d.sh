# new bench.py (configs, strong scaling, pipelined value), block size A/B, cfg 3/4/5 baselines, ncu of the current kernel
mkdir -p gpurun_out
L=ibwa_b200/libb200aln.so
python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_cfg2.json 2> gpurun_out/r2d_cfg2.err; tail -c 2500 gpurun_out/r2d_cfg2.json; tail -3 gpurun_out/r2d_cfg2.err
scripts/ab2.sh 10000000 "b128|$L|" "b32|$L|--set search_block=32" 2>&1 | tee gpurun_out/r2d_ab.txt
scripts/ab2.sh 1250000 "b128|$L|" "b32|$L|--set search_block=32" "b32k1|$L|--set search_block=32 --in-flight 1" 2>&1 | tee -a gpurun_out/r2d_ab.txt
scripts/ab2.sh 262144 "b128|$L|" "b32|$L|--set search_block=32" 2>&1 | tee -a gpurun_out/r2d_ab.txt
python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/r2d_cfg3.json 2> gpurun_out/r2d_cfg3.err; tail -c 1500 gpurun_out/r2d_cfg3.json; tail -3 gpurun_out/r2d_cfg3.err
python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/r2d_cfg4.json 2> gpurun_out/r2d_cfg4.err; tail -c 1500 gpurun_out/r2d_cfg4.json; tail -3 gpurun_out/r2d_cfg4.err
python bench.py --config 5 --steps 3 --warmup 3 > gpurun_out/r2d_cfg5.json 2> gpurun_out/r2d_cfg5.err; tail -c 1500 gpurun_out/r2d_cfg5.json; tail -3 gpurun_out/r2d_cfg5.err
ncu --set full --clock-control none --import-source on -k regex:k_search -c 1 -o gpurun_out/r2d_search python bench.py --reads 4000000 --steps 1 --warmup 1 --no-cpu-baseline --in-flight 1 > gpurun_out/r2d_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
