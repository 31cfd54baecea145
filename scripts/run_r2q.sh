mkdir -p gpurun_out
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2q_cfg2.json 2> gpurun_out/r2q_cfg2.err; echo "exit $?"; python -c "
import json;d=json.loads(open('gpurun_out/r2q_cfg2.json').read().strip().splitlines()[-1]);print('cfg2 value %.2fM e2e %.2fM'%(d['value']/1e6,d['e2e']['value']/1e6), d['steady_state'], d['hbm_in_use_gib'], d['parity'], {k:d['roofline'][k] for k in ('frac','frac_of_transaction_roof_reads','traffic')})"
timeout 600 python bench.py --reads 1250000 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_small.json 2> gpurun_out/r2q_small.err; python -c "
import json;d=json.loads(open('gpurun_out/r2q_small.json').read().strip().splitlines()[-1]);print('1.25M value %.2fM e2e %.2fM K=%d'%(d['value']/1e6,d['e2e']['value']/1e6,d['in_flight']), d['steady_state'], d['hbm_in_use_gib'])"
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2q_ref.json 2> gpurun_out/r2q_ref.err; echo "exit $?"; tail -c 700 gpurun_out/r2q_ref.json
