# final captures: ncu --set full of k_search and k_width at the benchmarked read count, and the launch list
mkdir -p gpurun_out
python bench.py --reads 10000000 --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 > gpurun_out/r2k_plain.json 2> gpurun_out/r2k_plain.err; tail -c 300 gpurun_out/r2k_plain.json
ncu --set full --clock-control none --import-source on -k regex:k_search -c 1 -o gpurun_out/r2k_search python bench.py --reads 10000000 --steps 1 --warmup 1 --no-cpu-baseline --in-flight 1 > gpurun_out/r2k_ncu_s.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_width -c 1 -o gpurun_out/r2k_width python bench.py --reads 10000000 --steps 1 --warmup 1 --no-cpu-baseline --in-flight 1 > gpurun_out/r2k_ncu_w.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --reads 10000000 --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 > gpurun_out/r2k_ncu_l.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2k_launches.csv
