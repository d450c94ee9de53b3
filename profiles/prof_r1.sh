set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:root_eval -s 3 -c 1 -o gpurun_out/prof_root_eval $CMD > gpurun_out/ncu2.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:clv_update -s 150 -c 3 -o gpurun_out/prof_clv_update $CMD > gpurun_out/ncu3.log 2>&1
$CMD > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fitch_level -s 120 -c 3 -o gpurun_out/prof_fitch $CMD > gpurun_out/ncu4.log 2>&1
tail -3 gpurun_out/ncu*.log
ls -la gpurun_out
