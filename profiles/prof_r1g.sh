# Round-1 (seventh pass) recipe: walk4m_kernel (one scaling counter per lane), Fitch plan kernel on a side stream,
# batched branch-length smoothing leg in the bench.  Every ncu pass is preceded by the same command run plain (exit 0).
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1p.json 2> gpurun_out/bench_r1p.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_r1p_ref.json 2>> gpurun_out/bench_r1p.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/plaing1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r1g.csv $CMD > gpurun_out/ncug1.log 2>&1
$CMD > gpurun_out/plaing2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:walk4m -s 3 -c 1 -o gpurun_out/prof_walk4m_g $CMD > gpurun_out/ncug2.log 2>&1
$CMD > gpurun_out/plaing3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spr_.*plan_kernel -s 3 -c 4 -o gpurun_out/prof_spr_plan_g $CMD > gpurun_out/ncug3.log 2>&1
$CMD > gpurun_out/plaing4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fitch_search_walk -s 2 -c 1 -o gpurun_out/prof_fitch_search_walk_g $CMD > gpurun_out/ncug4.log 2>&1
CMD2="python profiles/t_aa.py"
$CMD2 > gpurun_out/plaing5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval20_rows -s 1 -c 1 -o gpurun_out/prof_eval20_rows_g $CMD2 > gpurun_out/ncug5.log 2>&1
tail -n 2 gpurun_out/ncug1.log gpurun_out/ncug2.log gpurun_out/ncug3.log gpurun_out/ncug4.log gpurun_out/ncug5.log | cut -c1-300
CMD3="python profiles/t_pll_batch.py"
$CMD3 > gpurun_out/plaing6.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:opt_batch -s 200 -c 1 -o gpurun_out/prof_opt4_batch_g $CMD3 > gpurun_out/ncug6.log 2>&1
tail -n 2 gpurun_out/ncug6.log | cut -c1-300
