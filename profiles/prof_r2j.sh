#!/bin/bash
# round 2, pass j: full GPU test pass after lock support / oracle frequency normalisation / PHYLIP forms
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2j_tests.log
tail -8 gpurun_out/r2j_tests.log
