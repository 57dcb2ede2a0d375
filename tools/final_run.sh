#!/bin/bash
# Round-end validation on the GPU box: gpu tests, smoke, both bench arms, then (each only after its plain run exited 0)
# the ncu launch list of the bench command and one full capture each of the pairs and prep_rows kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
set -o pipefail
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 400 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; rc=$?; echo "bench rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:'prep_|seg_count|window_|finalize' -c 60 --csv --log-file gpurun_out/launches_final.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_l.log 2>&1
  timeout 100 python tools/profile_step.py --windows 1184 --reps 2 > gpurun_out/plain.log 2>&1 && {
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:window_pairs_tc_kernel -c 1 -s 1 -f \
        -o gpurun_out/prof_pairs_final python tools/profile_step.py --windows 1184 --reps 2 > gpurun_out/ncu_pairs.log 2>&1
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:prep_rows_kernel -c 1 -s 1 -f \
        -o gpurun_out/prof_prep_final python tools/profile_step.py --windows 1184 --reps 2 > gpurun_out/ncu_prep.log 2>&1
  }
fi
tail -2 gpurun_out/pytest.log; cat gpurun_out/smoke.log | tail -1; tail -c 600 gpurun_out/bench_ref.log; echo; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
r = d["roofline"]
print("ms_per_step %.3f value %.3e pairs_ms %.3f frac %.3f fp64 %.3f e2e_ms %.3f e2e %.3e cpu %.3e clocks %s" % (
    d["ms_per_step"], d["value"], r["kernel_ms"], r["frac"], r["fp64"]["frac"], d["e2e"]["ms_per_step"], d["e2e"]["value"],
    d["cpu_baseline"]["value"], d["clocks"]))
PY
ls -la gpurun_out/*.ncu-rep | tail -3
