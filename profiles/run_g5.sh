python -m pytest tests/test_neighbourhood_gpu.py -x -q 2>&1 | tail -15
python profiles/t_modes.py walkm walk 2>&1 | tail -2
for v in wm2 wm4; do echo "== $v"; PLG_LIB_PATH=variants/$v.so python profiles/t_modes.py walkm 2>&1 | tail -1; done
