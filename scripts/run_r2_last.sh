# last check of the committed code: all GPU tests, smoke, the default bench line, and the two bench paths with eight steps in flight
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.txt 2>&1; tail -2 gpurun_out/r2l_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2l_default.json 2> gpurun_out/r2l_default.err; tail -c 400 gpurun_out/r2l_default.json; echo
timeout 600 python bench.py --reads 1250000 --steps 20 --no-cpu-baseline > gpurun_out/r2l_1250k.json 2> gpurun_out/r2l_1250k.err; python -c "
import json;d=json.loads(open('gpurun_out/r2l_1250k.json').read().strip().splitlines()[-1]);print('1.25M', 'value %.2fM e2e %.2fM K=%d'%(d['value']/1e6,d['e2e']['value']/1e6,d['in_flight']), d['parity'])"
timeout 600 python bench.py --config 1 > gpurun_out/r2l_cfg1.json 2> gpurun_out/r2l_cfg1.err; python -c "
import json;d=json.loads(open('gpurun_out/r2l_cfg1.json').read().strip().splitlines()[-1]);print('cfg1', 'value %.2fM e2e %.2fM K=%d'%(d['value']/1e6,d['e2e']['value']/1e6,d['in_flight']), d['parity'], d['cpu_baseline']['value'])"
