#!/bin/bash
# round 2, pass a: GPU tests, walk4m variants (deferred log / rescale pre-vote), ncu of the new default
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
python profiles/t_walk_variants.py variants/r1.so "" variants/n_rt.so variants/n_nopv.so variants/n_rt_nopv.so variants/n_b2.so > gpurun_out/r2a_variants.log 2>&1
cat gpurun_out/r2a_variants.log
NT=256 NS=100000 REPS=2 python profiles/t_walk_variants.py variants/r1.so "" >> gpurun_out/r2a_variants.log 2>&1
tail -2 gpurun_out/r2a_variants.log
REPS=2 python profiles/t_walk_variants.py "" > gpurun_out/r2a_plain.log 2>&1 &&
REPS=2 ncu --set full --clock-control none --import-source on -k regex:walk4m -s 2 -c 1 -o gpurun_out/prof_walk4m_r2a python profiles/t_walk_variants.py "" > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_ncu.log
