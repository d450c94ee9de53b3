# Round-1 (fourth pass) recipe: walk kernels shipped everywhere (walk4m_kernel, fitch_search_walk_kernel,
# twalk4_kernel, fitch_walk_kernel).  Every ncu pass is preceded by the same command run plain (exit 0).
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_r1j_ref.json 2>> gpurun_out/bench_r1j.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/plaind1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1d.csv $CMD > gpurun_out/ncud1.log 2>&1
$CMD > gpurun_out/plaind2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:walk4m -s 3 -c 1 -o gpurun_out/prof_walk4m_d $CMD > gpurun_out/ncud2.log 2>&1
$CMD > gpurun_out/plaind3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fitch_search_walk -s 2 -c 1 -o gpurun_out/prof_fitch_search_walk $CMD > gpurun_out/ncud3.log 2>&1
tail -n 2 gpurun_out/ncud1.log gpurun_out/ncud2.log gpurun_out/ncud3.log | cut -c1-300
