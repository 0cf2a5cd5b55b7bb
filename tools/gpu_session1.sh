#!/bin/bash
# First GPU session: smoke, gpu tests, golden fixtures from the reference, reference + own bench, ncu.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== golden"; timeout 1500 python tools/make_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/golden.log
echo "== profile_run c1/c2"; timeout 600 python tools/profile_run.py --workload c1 --reps 3 > gpurun_out/prun_c1.log 2>&1; tail -4 gpurun_out/prun_c1.log
timeout 600 python tools/profile_run.py --workload c2 --reps 2 > gpurun_out/prun_c2.log 2>&1; tail -3 gpurun_out/prun_c2.log
timeout 600 python tools/profile_run.py --workload c2 --reps 2 --flags 4 > gpurun_out/prun_c2_flat.log 2>&1; tail -3 gpurun_out/prun_c2_flat.log
echo "== bench reference"; timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json
echo "== bench ours"; timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?"; cat gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
echo "== ncu launches"
timeout 600 python tools/profile_run.py --workload c1 --reps 2 > gpurun_out/plain_c1.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c1.csv python tools/profile_run.py --workload c1 --reps 2 > gpurun_out/ncu_launches.log 2>&1
echo "rc=$?"
echo "== ncu full k_extend"
timeout 600 python tools/profile_run.py --workload c1 --reps 2 > gpurun_out/plain_c1b.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 12 -c 3 -o gpurun_out/prof_extend_r1 python tools/profile_run.py --workload c1 --reps 2 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; ls -la gpurun_out
