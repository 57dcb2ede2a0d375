#!/bin/bash
# On the GPU box: role-time breakdown (tools/role_times.py) for every variants/roles*.so; restores the default library.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp impop_b200/libimpop_b200.so /tmp/default.so
: > gpurun_out/roles.log
for v in $(ls variants/roles*.so | xargs -n1 basename | sed 's/\.so$//'); do
  cp "variants/$v.so" impop_b200/libimpop_b200.so
  echo "=== $v" >> gpurun_out/roles.log
  timeout 120 python tools/role_times.py $ROLE_ARGS >> gpurun_out/roles.log 2>&1
done
cp /tmp/default.so impop_b200/libimpop_b200.so
cat gpurun_out/roles.log
