# parking of sparse warps: parity (golden + checked build), then A/B at the launch sizes that matter
mkdir -p gpurun_out
L=ibwa_b200/libb200aln.so
python -m pytest tests/test_gpu_parity.py tests/test_gpu_checked.py -m gpu -x -q -k "parked or golden or checked or three_passes or random_genome" > gpurun_out/r2g_pytest.txt 2>&1; tail -8 gpurun_out/r2g_pytest.txt
AB_STEPS=18 scripts/ab2.sh 1250000 "susp16|$L|" "susp0|$L|--set susp=0" "susp24|$L|--set susp=24" "susp8|$L|--set susp=8" "susp16k5|$L|--in-flight 5" 2>&1 | tee gpurun_out/r2g_ab.txt
AB_STEPS=36 scripts/ab2.sh 262144 "susp16k6|$L|--in-flight 6" "susp0k6|$L|--set susp=0 --in-flight 6" 2>&1 | tee -a gpurun_out/r2g_ab.txt
AB_STEPS=6 scripts/ab2.sh 10000000 "susp16|$L|" "susp0|$L|--set susp=0" 2>&1 | tee -a gpurun_out/r2g_ab.txt
B200ALN_TIMELINE=1 python bench.py --reads 1250000 --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_tl.json 2> gpurun_out/r2g_tl.err; grep timeline gpurun_out/r2g_tl.err | tail -16
python bench.py --config 3 --steps 9 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_cfg3.json 2> gpurun_out/r2g_cfg3.err; python -c "
import json;d=json.loads(open('gpurun_out/r2g_cfg3.json').read().strip().splitlines()[-1]);print('cfg3 value %.2fM seq %.2fM e2e %.2fM'%(d['value']/1e6,d['sequential']['value']/1e6,d['e2e']['value']/1e6), d['kernel_ms'], d['parity'])"
