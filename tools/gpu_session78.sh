#!/bin/bash
# new defaults (half / quarter-occupancy trace launches, 128-thread blocks, 4 wavefronts on L2-resident scenes): validation + bench lines + evidence
mkdir -p gpurun_out
echo "== gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench N=1"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s78_n1.json 2> gpurun_out/bench_s78_n1.err; python - <<PY
import json
j=json.loads(open("gpurun_out/bench_s78_n1.json").read().strip().splitlines()[-1])
print("n1 value", round(j["value"],1), "ms", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["value"],1), "frac", round(j["roofline"]["frac"],3), "pipes", j["pipelines"], j["clocks"], j["cpu_baseline"]["value"])
PY
for w in c1 c3 c3i c4; do
timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s78_$w.json 2> gpurun_out/bench_s78_$w.err; python - <<PY
import json
j=json.loads(open("gpurun_out/bench_s78_$w.json").read().strip().splitlines()[-1])
print("$w value", round(j["value"],1), "ms", round(j["ms_per_step"],2), "e2e ms", round(j["e2e"]["ms_per_step"],1), "pipes", j["pipelines"])
PY
done
echo "== launch list c2m"
timeout 600 python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/plain_s78_c2m.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_s78_c2m.csv python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/ncu_s78_l.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/plain_s78_c2m.log
echo "== ncu full c2m"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_trace' -s 8 -c 4 -o gpurun_out/prof_s78_c2m python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/ncu_s78_f.log 2>&1
echo "rc=$?"
