#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sanitize_run.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/sanitize_plain.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_run.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/sanitize_memcheck.log
