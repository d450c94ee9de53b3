python -m pytest tests/test_neighbourhood_gpu.py tests/test_configs_gpu.py -x -q 2>&1 | tail -5
PLG_TIMING=1 python profiles/t_host.py 2>&1 | head -24
PLG_FITCH_NB_MODE=ops PLG_TIMING=1 python profiles/t_host.py 2>&1 | head -20 | tail -7
