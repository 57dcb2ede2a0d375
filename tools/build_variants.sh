#!/bin/bash
# Build experiment variants of libimpop_b200.so HERE (no GPU needed) into variants/<name>.so; they travel to the
# GPU box with the snapshot.  Usage: tools/build_variants.sh name1="-DFOO=1 -DBAR" name2="" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
for spec in "$@"; do
  name="${spec%%=*}"; defs="${spec#*=}"
  [ "$name" = "$spec" ] && defs=""
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3,-pthread -shared $defs \
    -o "variants/$name.so" impop_b200/csrc/api.cu impop_b200/csrc/window_kernels.cu impop_b200/csrc/aux_kernels.cu impop_b200/csrc/ingest.cpp &
done
wait
ls -la variants/
