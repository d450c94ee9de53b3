# Round-1 (fifth pass) recipe: device-planned SPR neighbourhoods (spr_plan_kernel, spr_lik_plan_kernel),
# walk kernels unchanged.  Every ncu pass is preceded by the same command run plain (exit 0).
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_r1k_ref.json 2>> gpurun_out/bench_r1k.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/plaine1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r1e.csv $CMD > gpurun_out/ncue1.log 2>&1
$CMD > gpurun_out/plaine2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:walk4m -s 3 -c 1 -o gpurun_out/prof_walk4m_e $CMD > gpurun_out/ncue2.log 2>&1
$CMD > gpurun_out/plaine3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spr_.*plan_kernel -s 3 -c 4 -o gpurun_out/prof_spr_plan_e $CMD > gpurun_out/ncue3.log 2>&1
$CMD > gpurun_out/plaine4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fitch_search_walk -s 2 -c 1 -o gpurun_out/prof_fitch_search_walk_e $CMD > gpurun_out/ncue4.log 2>&1
CMD2="python profiles/t_aa.py"
$CMD2 > gpurun_out/plaine5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval20_rows -s 1 -c 1 -o gpurun_out/prof_eval20_rows_e $CMD2 > gpurun_out/ncue5.log 2>&1
tail -n 2 gpurun_out/ncue1.log gpurun_out/ncue2.log gpurun_out/ncue3.log gpurun_out/ncue4.log gpurun_out/ncue5.log | cut -c1-300
