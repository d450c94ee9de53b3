for v in old new; do echo "== $v"; PLG_LIB_PATH=variants/$v.so python profiles/t_modes.py walk 2>&1 | tail -1; done
PLG_LIB_PATH=variants/new.so python profiles/t_modes.py walk > /dev/null 2>&1 &&
PLG_LIB_PATH=variants/new.so REPS=1 ncu --set full --clock-control none --import-source on -k regex:walk4 -s 2 -c 1 -o gpurun_out/prof_walk4_pipe python profiles/t_modes.py walk > gpurun_out/ncu_g4.log 2>&1
tail -n 2 gpurun_out/ncu_g4.log
