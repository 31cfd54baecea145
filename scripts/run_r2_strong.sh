mkdir -p gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_strong_n$N.json 2> gpurun_out/r2_strong_n$N.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_strong_n$N.json').read().strip().splitlines()[-1]);print('N=$N', d['scaling'], 'value %.1fM seq %.1fM e2e %.1fM K=%d'%(d['value']/1e6,d['sequential']['value']/1e6,d['e2e']['value']/1e6,d['in_flight']))"
