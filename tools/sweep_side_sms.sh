#!/bin/bash
# bench.py at several SM budgets of the side-stream (coarse) backward chain: prints ms per step (device / e2e)
cd "$(dirname "$0")/.."
for sm in "$@"; do
  CTXNERF_SIDE_SMS=$sm timeout 100 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/sweep.json
  python - "$sm" <<'PY'
import json, sys
d = json.load(open("/tmp/sweep.json"))
print("side_sms", sys.argv[1], "step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "group", round(d["roofline"]["ms_per_launch"], 3))
PY
done
