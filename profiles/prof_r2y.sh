#!/bin/bash
# round 2, pass y: compute-sanitizer memcheck over every kernel family incl. the round-2 kernels
mkdir -p gpurun_out
timeout 300 python profiles/sanitize.py > gpurun_out/r2y_plain.log 2>&1 || { tail -20 gpurun_out/r2y_plain.log; exit 1; }
tail -3 gpurun_out/r2y_plain.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python profiles/sanitize.py > gpurun_out/sanitize_memcheck_r2y.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|pass . done" gpurun_out/sanitize_memcheck_r2y.log | tail -12
