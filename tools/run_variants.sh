#!/bin/bash
# On the GPU box: time every variants/*.so (or the names given) with a short resident-only bench; restores the default.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp impop_b200/libimpop_b200.so /tmp/default.so
names="$@"; [ -z "$names" ] && names=$(ls variants/*.so | xargs -n1 basename | sed 's/\.so$//')
: > gpurun_out/variants.log
for v in $names; do
  cp "variants/$v.so" impop_b200/libimpop_b200.so
  echo "=== $v" >> gpurun_out/variants.log
  timeout 120 python - >> gpurun_out/variants.log 2>&1 <<'PY'
import json, subprocess, sys
from impop_b200.engine import Context
ctx = Context(0)
print("selftest mismatches:", ctx.selftest_division(1 << 24, 3))
ctx.close()
dbg = subprocess.run([sys.executable, "tools/gpu_debug.py"], capture_output=True, text=True).stdout
print("debug: exact tc lines", dbg.count("tc: A ok=True I mismatches=0") , "of 5;", "pi exact" , dbg.count("pi exact=True"), "of 10")
out = subprocess.run([sys.executable, "bench.py", "--steps", "10", "--warmup", "3", "--no-cpu", "--no-others"], capture_output=True, text=True)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    r = d["roofline"]
    print("ms_per_step %.3f pairs_ms %.3f prep_ms %.3f e2e_ms %.3f same %s" % (d["ms_per_step"], r["kernel_ms"], r["step_share"]["prep"] * d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["matches_resident_run"]))
except Exception as exc:
    print("bench failed", exc, out.stdout[-500:], out.stderr[-2000:])
PY
done
cp /tmp/default.so impop_b200/libimpop_b200.so
cat gpurun_out/variants.log
