# repeated CLI runs under different driver settings: every run must write the same bytes
mkdir -p gpurun_out
CLI_STRESS_ENVS="B200ALN_INFLIGHT=1;B200ALN_INFLIGHT=6,B200ALN_MERGE=1;B200ALN_INFLIGHT=6,B200ALN_MERGE=2;B200ALN_MERGE=1;;B200ALN_INFLIGHT=8,B200ALN_MERGE=1" timeout 1200 python tests/tools/cli_stress.py 10000000 31 > gpurun_out/r2y_stress.txt 2>&1; tail -40 gpurun_out/r2y_stress.txt | cut -c1-400
