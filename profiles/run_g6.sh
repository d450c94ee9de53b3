python profiles/t_modes.py walkm > /dev/null 2>&1 &&
REPS=1 ncu --set full --clock-control none --import-source on -k regex:walk4m -s 2 -c 1 -o gpurun_out/prof_walk4m python profiles/t_modes.py walkm > gpurun_out/ncu_g6.log 2>&1
tail -n 2 gpurun_out/ncu_g6.log
