# Round-1 profiling recipe (run under gpurun on one B200).  Every ncu pass is preceded by the same
# command run plain (must exit 0).  Outputs land in gpurun_out/; summaries are copied to profiles/.
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:clv4 -s 27 -c 5 -o gpurun_out/prof_clv4 $CMD > gpurun_out/ncu3.log 2>&1
$CMD > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fitch_level -s 28 -c 8 -o gpurun_out/prof_fitch $CMD > gpurun_out/ncu4.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_r1c_ref.json 2>> gpurun_out/bench_r1c.err
ls -la gpurun_out
