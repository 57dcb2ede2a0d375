#!/bin/bash
# On an N-GPU box: config 2 (weak) end-to-end leg with the presence rows tight / aligned and several sub-batch counts.
N=$1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/e2e_ab_n$N.log
for spec in ${SPECS:-"tight 8" "aligned 4" "aligned 8"}; do
  set -- $spec
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --transfer $1 --sub-batches $2 2>gpurun_out/b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('N=$N $1 sub $2: ms_per_step %.3f e2e_ms %.3f host_enqueue_ms %.3f h2d %d same %s' % (d['ms_per_step'], e['ms_per_step'], e['host_enqueue_ms_per_step'], e['h2d_bytes_per_step'], e['matches_resident_run']))" >> gpurun_out/e2e_ab_n$N.log
done
cat gpurun_out/e2e_ab_n$N.log; tail -2 gpurun_out/b.err
