#!/bin/bash
# On an N-GPU box: bench.py for the multi-GPU configs (5: tile-grid split, strong; 2: windows sharded, weak; 3 --strong: the
# whole genome's windows split over the ranks).  Usage: tools/scale_run.sh N "5 2 3"
N=$1; CFGS=${2:-"5 2 3"}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in $CFGS; do
  extra=""; [ "$c" = 3 ] && extra="--strong"
  if [ "$N" = 1 ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --config $c $extra --no-others --no-cpu > gpurun_out/r2_scale_c${c}_n$N.json 2> gpurun_out/r2_scale_c${c}_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --config $c $extra > gpurun_out/r2_scale_c${c}_n$N.json 2> gpurun_out/r2_scale_c${c}_n$N.err
  fi
  echo "config $c N=$N rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_scale_c${c}_n$N.json").read().strip().splitlines()[-1])
    print("   value %.3e  ms_per_step %.3f  e2e %.3e (%.3f ms)  pairs_ms %.3f  %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], d["scaling"]))
except Exception as exc:
    print("   no line:", exc); print(open("gpurun_out/r2_scale_c${c}_n$N.err").read()[-800:])
PY
done
