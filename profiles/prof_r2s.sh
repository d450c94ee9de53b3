#!/bin/bash
# round 2, pass s: the default bench line (N = 1) with every leg, and the reference arm
mkdir -p gpurun_out
( time python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2s.json 2> gpurun_out/bench_r2s.err ) 2>&1 | tail -3
tail -c 400 gpurun_out/bench_r2s.err
( time PLG_BENCH_REF_SECONDS=30 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2s_ref.json 2> gpurun_out/bench_r2s_ref.err ) 2>&1 | tail -3
head -c 600 gpurun_out/bench_r2s.json; echo; head -c 400 gpurun_out/bench_r2s_ref.json
