#!/bin/bash
for fs in 1 8 15 16 17 32 48 64 96; do echo -n "fs $fs: "; timeout 900 python tools/sweep.py --workload c3 --first-sample $fs --reps 2 2>&1 | tail -1 | cut -c60-140; done
for fs in 0 64 128; do echo -n "c2 fs $fs: "; timeout 900 python tools/sweep.py --workload c2 --first-sample $fs --reps 2 2>&1 | tail -1 | cut -c60-140; done
