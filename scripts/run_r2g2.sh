# two GPUs, one process: the command line's driver sends launches round the workers of both devices
mkdir -p gpurun_out
nvidia-smi -L
rm -f gpurun_out/r2g2_trace.txt
CLI_E2E_TRACE=gpurun_out/r2g2_trace.txt timeout 900 python tests/tools/cli_e2e.py 10000000 500000 > gpurun_out/r2g2_cli.txt 2>&1; tail -16 gpurun_out/r2g2_cli.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cli" 2>&1 | tail -3
