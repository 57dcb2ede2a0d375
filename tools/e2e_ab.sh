#!/bin/bash
# On the GPU box: the end-to-end leg of bench.py with the presence rows transferred tight / aligned, at several sub-batch counts.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/e2e_ab.log
for tr in ${TRANSFERS:-aligned tight}; do for sb in ${SUBS:-4 8}; do
  python bench.py --steps 10 --warmup 3 --no-cpu --no-others --transfer $tr --sub-batches $sb 2>gpurun_out/b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('$tr sub $sb: ms_per_step %.3f e2e_ms %.3f host_enqueue_ms %.3f h2d %d same %s' % (d['ms_per_step'], e['ms_per_step'], e['host_enqueue_ms_per_step'], e['h2d_bytes_per_step'], e['matches_resident_run']))" >> gpurun_out/e2e_ab.log
done; done
cat gpurun_out/e2e_ab.log; tail -2 gpurun_out/b.err
