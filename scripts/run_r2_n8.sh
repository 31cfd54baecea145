# strong (default) and weak scaling on 8 GPUs of one box, one process per GPU
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) bench.py --gpus $1 --steps 18 --warmup 3 --no-cpu-baseline "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; tail -c 1400 gpurun_out/$2.json | head -c 1400; echo; }
N=${1:-8}
run $N r2_strong_n$N
run $N r2_weak_n$N --scaling weak
