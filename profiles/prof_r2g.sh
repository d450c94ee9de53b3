#!/bin/bash
# round 2, pass g: GPU tests, the restructured bench (default = 256 x 100k lnL target shape) and its reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2g_tests.log
tail -5 gpurun_out/r2g_tests.log
( time python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err ) 2>&1 | tail -3
tail -c 600 gpurun_out/bench_r2g.err
( time PLG_BENCH_REF_SECONDS=30 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2g_ref.json 2> gpurun_out/bench_r2g_ref.err ) 2>&1 | tail -3
head -c 1500 gpurun_out/bench_r2g.json; echo; cat gpurun_out/bench_r2g_ref.json | head -c 1200
