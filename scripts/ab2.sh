#!/bin/bash
# A/B of engine builds / knobs on one box: scripts/ab2.sh <reads> "tag|lib.so|extra bench flags" ...
# each run: bench.py --steps 3 --warmup 3 --no-cpu-baseline.  Prints value / kernel_ms / parity per run.
reads=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  IFS='|' read -r tag lib flags <<< "$spec"
  B200ALN_LIB=$PWD/$lib python bench.py --steps ${AB_STEPS:-3} --warmup 3 --no-cpu-baseline --reads $reads $flags > gpurun_out/ab_${tag}_$reads.json 2> gpurun_out/ab_${tag}_$reads.err
  python - "$tag" $reads gpurun_out/ab_${tag}_$reads.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
    print(sys.argv[1], sys.argv[2], "value %.2fM seq %.2fM e2e %.2fM" % (d["value"] / 1e6, d["sequential"]["value"] / 1e6, d["e2e"]["value"] / 1e6),
          {k: round(v, 2) for k, v in d["kernel_ms"].items()}, "parity", d["parity"], flush=True)
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e, flush=True)
PY
done
