#!/bin/bash
run() { echo -n "$1 minb=$2 sblocks=$3 tblocks=$4: "; RTB_SHADE_MINB=$2 RTB_SHADE_BLOCKS=$3 RTB_TRACE_BLOCKS=$4 timeout 900 python tools/sweep.py --workload $1 --reps 3 2>&1 | tail -1 | cut -c75-170; }
run c2 3 2 0
run c2 2 2 0
run c2 2 1 0
run c2 2 3 0
run c2 2 2 3
run c2 2 2 5
run c3 3 2 0
run c3 2 2 0
run c4 3 2 0
run c4 2 2 0
