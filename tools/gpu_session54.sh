#!/bin/bash
mkdir -p gpurun_out
for w in c1 c4; do timeout 900 python bench.py --impl reference --workload $w --steps 3 --warmup 1 > gpurun_out/bench_s54_ref_$w.json 2> gpurun_out/bench_s54_ref_$w.err; echo "$w rc=$?"; python -c "
import json
j=json.loads(open('gpurun_out/bench_s54_ref_$w.json').read().strip().splitlines()[-1])
print('$w', 'value', round(j.get('value',0),1), 'ms', round(j.get('ms_per_step',0),2), 'e2e', j.get('e2e'), j.get('reference_harness_error'))
"; done
timeout 1500 python bench.py --impl reference --workload c3 --steps 1 --warmup 0 > gpurun_out/bench_s54_ref_c3.json 2> gpurun_out/bench_s54_ref_c3.err; echo "c3 rc=$?"; python -c "
import json
j=json.loads(open('gpurun_out/bench_s54_ref_c3.json').read().strip().splitlines()[-1])
print('c3', 'value', round(j.get('value',0),1), 'ms', round(j.get('ms_per_step',0),2), 'e2e', j.get('e2e'), j.get('reference_harness_error'), j.get('reference_bvh_build_ms_host'))
"
