#!/bin/bash
# A/B of engine builds on one box: scripts/ab.sh <reads> lib1.so lib2.so ...   (each run: bench.py --steps 3 --warmup 3)
# extra bench flags through AB_FLAGS.  Prints value / kernel_ms per build.
reads=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  tag=$(basename $lib .so)
  B200ALN_LIB=$PWD/$lib python bench.py --steps 3 --warmup 3 --no-cpu-baseline --reads $reads $AB_FLAGS > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - "$tag" gpurun_out/ab_$tag.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.2fM e2e %.2fM" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6), d["kernel_ms"], "parity", d["parity"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
