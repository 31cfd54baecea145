# Q16 A/B, pipelining at small launches with enough steps, timeline, checked build over the golden set
mkdir -p gpurun_out
L=ibwa_b200/libb200aln.so
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.txt 2>&1; tail -3 gpurun_out/r2e_pytest.txt
AB_STEPS=5 scripts/ab2.sh 10000000 "q16|$L|" "q32|$L|--set q16=0" 2>&1 | tee gpurun_out/r2e_ab.txt
AB_STEPS=18 scripts/ab2.sh 1250000 "q16b128|$L|" "q16b32|$L|--set search_block=32" "q16b32k2|$L|--set search_block=32 --in-flight 2" "q16b32k4|$L|--set search_block=32 --in-flight 4" 2>&1 | tee -a gpurun_out/r2e_ab.txt
AB_STEPS=24 scripts/ab2.sh 262144 "q16b128|$L|" "q16b32|$L|--set search_block=32" "q16b32k6|$L|--set search_block=32 --in-flight 6" 2>&1 | tee -a gpurun_out/r2e_ab.txt
B200ALN_TIMELINE=1 python bench.py --reads 1250000 --steps 12 --warmup 3 --no-cpu-baseline --set search_block=32 > gpurun_out/r2e_tl.json 2> gpurun_out/r2e_tl.err; grep timeline gpurun_out/r2e_tl.err | tail -40
python bench.py --config 3 --steps 9 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_cfg3.json 2> gpurun_out/r2e_cfg3.err; tail -c 600 gpurun_out/r2e_cfg3.json
python bench.py --config 3 --steps 9 --warmup 3 --no-cpu-baseline --set arena_cap=4096 > gpurun_out/r2e_cfg3_a4k.json 2> gpurun_out/r2e_cfg3_a4k.err; tail -c 600 gpurun_out/r2e_cfg3_a4k.json
