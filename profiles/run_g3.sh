python -m pytest tests/test_treewalk_gpu.py tests/test_neighbourhood_gpu.py tests/test_lik_gpu.py -x -q 2>&1 | tail -5
python profiles/t_modes.py walk visits 2>&1 | tail -3
python profiles/t_stream.py walk 2>&1 | tail -3
PLG_TIMING=1 python profiles/t_host.py 2>&1 | tail -22
