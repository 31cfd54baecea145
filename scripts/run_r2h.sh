# after the mid-pass parking fix: parity, then the in-flight x parking matrix at the per-GPU sizes of strong scaling
mkdir -p gpurun_out
L=ibwa_b200/libb200aln.so
python -m pytest tests/test_gpu_parity.py tests/test_gpu_checked.py -m gpu -x -q > gpurun_out/r2h_pytest.txt 2>&1; tail -4 gpurun_out/r2h_pytest.txt
AB_STEPS=24 scripts/ab2.sh 1250000 "s0k6|$L|--set susp=0 --in-flight 6" "s16k6|$L|--in-flight 6" "s16k8|$L|--in-flight 8" "s12k8b32|$L|--in-flight 8 --set susp=12 --set search_block=32" 2>&1 | tee gpurun_out/r2h_ab.txt
AB_STEPS=18 scripts/ab2.sh 2500000 "s0k3|$L|--set susp=0" "s0k6|$L|--set susp=0 --in-flight 6" "s16k6|$L|--in-flight 6" 2>&1 | tee -a gpurun_out/r2h_ab.txt
AB_STEPS=12 scripts/ab2.sh 5000000 "s0k3|$L|--set susp=0" "s16k4|$L|--in-flight 4" "s0k4|$L|--set susp=0 --in-flight 4" 2>&1 | tee -a gpurun_out/r2h_ab.txt
python bench.py --config 3 --steps 9 --warmup 3 --no-cpu-baseline --set susp=0 > gpurun_out/r2h_cfg3.json 2> gpurun_out/r2h_cfg3.err; python -c "
import json;d=json.loads(open('gpurun_out/r2h_cfg3.json').read().strip().splitlines()[-1]);print('cfg3 susp0 value %.2fM seq %.2fM e2e %.2fM'%(d['value']/1e6,d['sequential']['value']/1e6,d['e2e']['value']/1e6), d['kernel_ms'], d['parity'])"
python bench.py --config 3 --steps 9 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_cfg3s.json 2> gpurun_out/r2h_cfg3s.err; python -c "
import json;d=json.loads(open('gpurun_out/r2h_cfg3s.json').read().strip().splitlines()[-1]);print('cfg3 susp16 value %.2fM seq %.2fM e2e %.2fM'%(d['value']/1e6,d['sequential']['value']/1e6,d['e2e']['value']/1e6), d['kernel_ms'], d['parity'])"
python bench.py --config 1 --steps 5 --warmup 3 > gpurun_out/r2h_cfg1.json 2> gpurun_out/r2h_cfg1.err; tail -c 900 gpurun_out/r2h_cfg1.json
