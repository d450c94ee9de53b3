python -m pytest tests/test_neighbourhood_gpu.py -x -q -k planners 2>&1 | tail -2
for v in a b c d e; do echo "== $v"; PLG_LIB_PATH=variants/$v.so python profiles/t_modes.py walkm 2>&1 | tail -1; done
