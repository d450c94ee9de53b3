# Round-1 (sixth pass) recipe: walk4m_kernel with normal-form visits and straight-line neighbour chains,
# plan kernels with shared-memory edge records.  Every ncu pass is preceded by the same command run plain (exit 0).
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_r1m_ref.json 2>> gpurun_out/bench_r1m.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/plainf1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r1f.csv $CMD > gpurun_out/ncuf1.log 2>&1
$CMD > gpurun_out/plainf2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:walk4m -s 3 -c 1 -o gpurun_out/prof_walk4m_f $CMD > gpurun_out/ncuf2.log 2>&1
$CMD > gpurun_out/plainf3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spr_.*plan_kernel -s 3 -c 4 -o gpurun_out/prof_spr_plan_f $CMD > gpurun_out/ncuf3.log 2>&1
$CMD > gpurun_out/plainf4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fitch_search_walk -s 2 -c 1 -o gpurun_out/prof_fitch_search_walk_f $CMD > gpurun_out/ncuf4.log 2>&1
CMD2="python profiles/t_aa.py"
$CMD2 > gpurun_out/plainf5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval20_rows -s 1 -c 1 -o gpurun_out/prof_eval20_rows_f $CMD2 > gpurun_out/ncuf5.log 2>&1
tail -n 2 gpurun_out/ncuf1.log gpurun_out/ncuf2.log gpurun_out/ncuf3.log gpurun_out/ncuf4.log gpurun_out/ncuf5.log | cut -c1-300
