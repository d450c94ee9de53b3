set -x
CMD="python profiles/t_lnl.py"
$CMD > gpurun_out/plain_tl.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:clv4 -s 30 -c 2 -o gpurun_out/prof_clv4_b $CMD > gpurun_out/ncu_tl.log 2>&1
tail -2 gpurun_out/ncu_tl.log
