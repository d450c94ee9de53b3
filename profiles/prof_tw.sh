# Tree-walk kernels (twalk4_kernel, fitch_walk_kernel): plain run, then ncu --set full on the same command.
set -x
CMD="python profiles/t_stream.py walk"
$CMD > gpurun_out/plain_tw.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:twalk4 -s 1 -c 1 -o gpurun_out/prof_twalk4_b $CMD > gpurun_out/ncu_tw1.log 2>&1
cat gpurun_out/plain_tw.log
tail -n 3 gpurun_out/ncu_tw1.log
