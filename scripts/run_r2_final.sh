# final validation of the committed code: all GPU tests, smoke, the five config lines, the CLI end to end
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.txt 2>&1; tail -3 gpurun_out/r2f_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_cfg2.json 2> gpurun_out/r2f_cfg2.err; tail -c 600 gpurun_out/r2f_cfg2.json; echo
for c in 3 4 5; do timeout 900 python bench.py --config $c --steps 6 --warmup 3 > gpurun_out/r2f_cfg$c.json 2> gpurun_out/r2f_cfg$c.err; tail -c 300 gpurun_out/r2f_cfg$c.json; echo; done
rm -f gpurun_out/r2f_trace.txt
CLI_E2E_TRACE=gpurun_out/r2f_trace.txt timeout 600 python tests/tools/cli_e2e.py 10000000 500000 > gpurun_out/r2f_cli.txt 2>&1; tail -16 gpurun_out/r2f_cli.txt
