#!/bin/bash
# round 2, pass p: clv20_kernel with the per-tile scaling decision; full GPU test pass; C5 bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2p_tests.log
tail -6 gpurun_out/r2p_tests.log
python profiles/t_aa.py 2>&1 | tail -2
python bench.py --workload aa_c5 --steps 5 --warmup 3 --no-extra > gpurun_out/bench_r2p_aa_c5.json 2> gpurun_out/bench_r2p_aa_c5.err; echo "bench rc=$?"
head -c 2500 gpurun_out/bench_r2p_aa_c5.json
