#!/bin/bash
for w in c2 c3; do
echo -n "$w tb256: "; timeout 900 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | cut -c60-170
for tb in 128 512; do echo -n "$w tb$tb: "; RTB_LIB=$PWD/tools/_exp/librtb_tb$tb.so timeout 900 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | cut -c60-170; done
done
