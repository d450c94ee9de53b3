# Round-1 (second pass) profiling recipe for the SHIPPED kernels (visit4b_kernel, fitch_level_kernel).
# Every ncu pass is preceded by the same command run plain (must exit 0).
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/plainb1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncub1.log 2>&1
$CMD > gpurun_out/plainb2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:visit4b -s 20 -c 4 -o gpurun_out/prof_visit4b $CMD > gpurun_out/ncub2.log 2>&1
$CMD > gpurun_out/plainb3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fitch_level -s 28 -c 6 -o gpurun_out/prof_fitch_b $CMD > gpurun_out/ncub3.log 2>&1
tail -3 gpurun_out/ncub1.log gpurun_out/ncub2.log gpurun_out/ncub3.log
ls -la gpurun_out | tail -12
