# per-GPU share of the strong-scaling step at 8 GPUs (1.25 M reads), 20 steps: launches in flight and parking thresholds
mkdir -p gpurun_out
AB_STEPS=20 bash scripts/ab2.sh 1250000 "k6|ibwa_b200/libb200aln.so|--in-flight 6" "k8|ibwa_b200/libb200aln.so|--in-flight 8" "k8_s24|ibwa_b200/libb200aln.so|--in-flight 8 --set susp=-24" "k6_s8|ibwa_b200/libb200aln.so|--in-flight 6 --set susp=-8" "k6_min1k|ibwa_b200/libb200aln.so|--in-flight 6 --set susp_min=1024" 2>&1 | tee gpurun_out/r2s2_ab.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grow_while or first_batches or cli" 2>&1 | tail -3
