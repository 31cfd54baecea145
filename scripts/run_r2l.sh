# repeat-rich genome line, engine-only launch list, refreshed config lines
mkdir -p gpurun_out
timeout 1200 python bench.py --repeats 0.05 --steps 3 --warmup 3 > gpurun_out/r2l_repeats.json 2> gpurun_out/r2l_repeats.err; tail -c 1800 gpurun_out/r2l_repeats.json; tail -4 gpurun_out/r2l_repeats.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --reads 10000000 --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 > gpurun_out/r2l_ncu_l.log 2>&1
timeout 600 python bench.py --config 1 --steps 20 --warmup 3 > gpurun_out/r2l_cfg1.json 2> gpurun_out/r2l_cfg1.err; tail -c 400 gpurun_out/r2l_cfg1.json
