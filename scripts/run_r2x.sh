# round-2 session 3: the CLI end to end with the driver's timeline
mkdir -p gpurun_out
rm -f gpurun_out/r2x_trace.txt
CLI_E2E_QUICK=1 CLI_E2E_TRACE=gpurun_out/r2x_trace.txt timeout 900 python tests/tools/cli_e2e.py 10000000 500000 > gpurun_out/r2x_cli.txt 2>&1; tail -14 gpurun_out/r2x_cli.txt
