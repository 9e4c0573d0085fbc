#!/bin/bash
# bench every experimental library variant (tools/build_variants.sh) + the default one
cd "$(dirname "$0")/.."
for lib in default contexture-nerf_b200/ctxnerf/variants/*.so; do
  if [ "$lib" = default ]; then unset CTXNERF_LIB; else export CTXNERF_LIB=$PWD/$lib; fi
  timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/bv.json
  python - "$lib" <<'PY'
import json, sys
try:
    d = json.load(open("/tmp/bv.json"))
    k = d["kernels"]
    print(f"{sys.argv[1].split('/')[-1]:32s} step {d['ms_per_step']:.3f} ms | " + " ".join(f"{n[4:]} {v['ms_per_launch']:.3f}" for n, v in k.items()))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
