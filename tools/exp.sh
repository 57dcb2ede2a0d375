#!/bin/bash
# Usage: tools/exp.sh "<defs variant 1>" "<defs variant 2>" ...   (run on the GPU box; rebuilds per variant)
mkdir -p gpurun_out
: > gpurun_out/exp.log
for defs in "$@"; do
  echo "=== variant: [$defs]" >> gpurun_out/exp.log
  IMPOP_NVCC_DEFS="$defs" python -c "from impop_b200 import build; build.build(force=True)" >> gpurun_out/exp.log 2>&1 || { echo "build failed" >> gpurun_out/exp.log; continue; }
  timeout 90 python - >> gpurun_out/exp.log 2>&1 <<'PY'
import json, subprocess, sys
from impop_b200.engine import Context
ctx = Context(0)
print("selftest mismatches:", ctx.selftest_division(1 << 24, 3))
ctx.close()
out = subprocess.run([sys.executable, "bench.py", "--steps", "10", "--warmup", "3", "--no-cpu"], capture_output=True, text=True)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    r = d["roofline"]
    print("ms_per_step %.3f pairs_ms %.3f prep_ms %.3f e2e_ms %.3f same %s" % (d["ms_per_step"], r["kernel_ms"], r["step_share"]["prep"] * d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["matches_resident_run"]))
except Exception as exc:
    print("bench failed", exc, out.stderr[-2000:])
PY
done
python -c "from impop_b200 import build; build.build(force=True)" > /dev/null 2>&1
cat gpurun_out/exp.log
