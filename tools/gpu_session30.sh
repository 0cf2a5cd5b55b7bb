#!/bin/bash
for w in c2 c3 c4 c1; do for c in 1 0; do echo -n "$w collapse $c: "; timeout 900 python tools/sweep.py --workload $w --collapse $c --reps 3 --count 2>&1 | tail -2 | cut -c75-250 | tr '\n' ' '; echo; done; done
