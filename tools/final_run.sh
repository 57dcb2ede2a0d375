#!/bin/bash
# Round-end validation on the GPU box: gpu tests, smoke, both bench arms, then (each only after its plain run exited 0)
# the ncu launch list of the bench command and one full capture each of the pairs and prep_rows kernels.
# Outputs under gpurun_out/ with the round tag R (default r2).
cd "$(dirname "$0")/.."
R=${1:-r2}
mkdir -p gpurun_out
set -o pipefail
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${R}_bench_reference_line.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/${R}_bench_line.json 2> gpurun_out/bench.err; rc=$?; echo "bench rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:'prep_|seg_count|window_|finalize' -s 60 -c 36 --csv --log-file gpurun_out/${R}_launches.csv \
      python bench.py --steps 10 --warmup 3 --no-cpu --no-others > gpurun_out/ncu_l.log 2>&1
  timeout 100 python tools/profile_step.py --windows 1184 --reps 2 --compact > gpurun_out/plain.log 2>&1 && {
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:window_pairs_tc_kernel -c 1 -s 1 -f \
        -o gpurun_out/prof_pairs_${R} python tools/profile_step.py --windows 1184 --reps 2 --compact > gpurun_out/ncu_pairs.log 2>&1
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:prep_rows_kernel -c 1 -s 1 -f \
        -o gpurun_out/prof_prep_${R} python tools/profile_step.py --windows 1184 --reps 2 --compact --full-pitch > gpurun_out/ncu_prep.log 2>&1
  }
fi
tail -2 gpurun_out/pytest.log; cat gpurun_out/smoke.log | tail -1; python - <<PY
import json
for name in ("gpurun_out/${R}_bench_reference_line.json", "gpurun_out/${R}_bench_line.json"):
    try:
        d = json.loads(open(name).read().strip().splitlines()[-1])
    except Exception as exc:
        print(name, "no line", exc); continue
    if d.get("impl") == "reference":
        print("reference: value %.3e ms_per_step %.1f cores %s" % (d["value"], d["ms_per_step"], d["cpu_baseline"]["cores"])); continue
    r = d["roofline"]
    print("ms_per_step %.3f value %.3e pairs_ms %.3f tensor.frac %.3f (executed %.3f) fp64 %.3f e2e_ms %.3f e2e %.3e h2d %d cpu %.3e clocks %s" % (
        d["ms_per_step"], d["value"], r["kernel_ms"], r["tensor"]["frac"], r["tensor"]["frac_executed"], r["fp64"]["frac"], d["e2e"]["ms_per_step"],
        d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"], d["cpu_baseline"]["value"], d["clocks"]))
    print("step_share", r["step_share"])
    print("parity", json.dumps(d["cpu_baseline"]["gpu_vs_oracle_on_original_columns"])[:600])
    for k, v in (d.get("other_configs") or {}).items():
        print(k, json.dumps(v)[:400])
PY
ls -la gpurun_out/*${R}*.ncu-rep 2>/dev/null | tail -3
