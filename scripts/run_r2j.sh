mkdir -p gpurun_out
L=ibwa_b200/libb200aln.so
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_checked.py -m gpu -x -q -s > gpurun_out/r2j_pytest.txt 2>&1; tail -4 gpurun_out/r2j_pytest.txt; grep -n "Abort\|internal" gpurun_out/r2j_pytest.txt | head -5
AB_STEPS=24 timeout 600 scripts/ab2.sh 1250000 "ring_k6|$L|--in-flight 6" "ring_k8|$L|--in-flight 8" "nopark_k6|$L|--in-flight 6 --set susp=0" 2>&1 | tee gpurun_out/r2j_ab.txt
AB_STEPS=36 timeout 600 scripts/ab2.sh 262144 "ring_k6|$L|--in-flight 6" 2>&1 | tee -a gpurun_out/r2j_ab.txt
timeout 600 python bench.py --config 3 --steps 9 --warmup 3 --no-cpu-baseline --in-flight 4 > gpurun_out/r2j_cfg3.json 2> gpurun_out/r2j_cfg3.err; python -c "
import json;d=json.loads(open('gpurun_out/r2j_cfg3.json').read().strip().splitlines()[-1]);print('cfg3 K4 value %.2fM seq %.2fM e2e %.2fM'%(d['value']/1e6,d['sequential']['value']/1e6,d['e2e']['value']/1e6), d['kernel_ms'], d['parity'])"
