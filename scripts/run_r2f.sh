# stream priorities A/B (short kernels first), new config tests
mkdir -p gpurun_out
L=ibwa_b200/libb200aln.so
python -m pytest tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/r2f_pytest.txt 2>&1; tail -15 gpurun_out/r2f_pytest.txt
AB_STEPS=18 scripts/ab2.sh 1250000 "prio_b32|$L|--set search_block=32" "prio_b128|$L|" 2>&1 | tee gpurun_out/r2f_ab.txt
B200ALN_PRIO=0 AB_STEPS=18 scripts/ab2.sh 1250000 "noprio_b32|$L|--set search_block=32" 2>&1 | tee -a gpurun_out/r2f_ab.txt
AB_STEPS=18 scripts/ab2.sh 1250000 "prio_b32k4|$L|--set search_block=32 --in-flight 4" 2>&1 | tee -a gpurun_out/r2f_ab.txt
AB_STEPS=36 scripts/ab2.sh 262144 "prio_b32k6|$L|--set search_block=32 --in-flight 6" "prio_b128k6|$L|--in-flight 6" 2>&1 | tee -a gpurun_out/r2f_ab.txt
AB_STEPS=6 scripts/ab2.sh 10000000 "prio|$L|" 2>&1 | tee -a gpurun_out/r2f_ab.txt
B200ALN_TIMELINE=1 python bench.py --reads 1250000 --steps 12 --warmup 3 --no-cpu-baseline --set search_block=32 > gpurun_out/r2f_tl.json 2> gpurun_out/r2f_tl.err; grep timeline gpurun_out/r2f_tl.err | tail -24
python bench.py --config 3 --steps 9 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_cfg3.json 2> gpurun_out/r2f_cfg3.err; tail -c 300 gpurun_out/r2f_cfg3.json; python -c "
import json;d=json.loads(open('gpurun_out/r2f_cfg3.json').read().strip().splitlines()[-1]);print('cfg3 value %.2fM seq %.2fM e2e %.2fM'%(d['value']/1e6,d['sequential']['value']/1e6,d['e2e']['value']/1e6))"
