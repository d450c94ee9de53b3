for v in p1 p0 p0h p1h; do echo "== $v"; PLG_LIB_PATH=variants/$v.so python profiles/t_modes.py walkm 2>&1 | tail -1; done
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; tail -3 gpurun_out/bench_r1h.err
