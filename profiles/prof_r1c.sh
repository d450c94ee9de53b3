# Round-1 (third pass) recipe: depth-first walk kernel shipped (walk4_kernel).  Every ncu pass is
# preceded by the same command run plain (must exit 0).  Outputs land in gpurun_out/.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1g.log 2>&1; tail -3 gpurun_out/pytest_r1g.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_r1g_ref.json 2>> gpurun_out/bench_r1g.err
python profiles/t_host.py 2> gpurun_out/t_host_r1g.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/plainc1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1c.csv $CMD > gpurun_out/ncuc1.log 2>&1
$CMD > gpurun_out/plainc2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:walk4 -s 3 -c 1 -o gpurun_out/prof_walk4 $CMD > gpurun_out/ncuc2.log 2>&1
tail -3 gpurun_out/ncuc1.log gpurun_out/ncuc2.log
cat gpurun_out/t_host_r1g.log
