# last check of the committed code: all GPU tests, smoke, the default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.txt 2>&1; tail -2 gpurun_out/r2l_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2l_default.json 2> gpurun_out/r2l_default.err; python -c "
import json;d=json.loads(open('gpurun_out/r2l_default.json').read().strip().splitlines()[-1]);print('default', 'value %.2fM e2e %.2fM K=%d launches %d'%(d['value']/1e6,d['e2e']['value']/1e6,d['in_flight'],d['gpu_launches']), d['parity'], d['clocks'])"
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -c 400
