#!/bin/bash
# On the GPU box: tools/compact_bench.py (original + compacted columns, resident timing) for every variants/*.so.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp impop_b200/libimpop_b200.so /tmp/default.so
names="$@"; [ -z "$names" ] && names=$(ls variants/*.so | xargs -n1 basename | sed 's/\.so$//')
: > gpurun_out/variants2.log
for v in $names; do
  cp "variants/$v.so" impop_b200/libimpop_b200.so
  echo "=== $v" >> gpurun_out/variants2.log
  timeout 120 python tools/compact_bench.py ${CB_WINDOWS:-4854} 2>&1 | grep -E "ms_per_step|counts equal" >> gpurun_out/variants2.log
done
cp /tmp/default.so impop_b200/libimpop_b200.so
cat gpurun_out/variants2.log
