# pair-fused visit kernel variants vs the shipped per-neighbour kernel
for v in "-DPLG_USE_VISITS=0" "-DPLG_USE_VISITS=1 -DPLG_VB_MINB=5" "-DPLG_USE_VISITS=1 -DPLG_VB_MINB=4" "-DPLG_USE_VISITS=1 -DPLG_VB_MINB=6"; do
  PLG_NVCC_EXTRA="$v" python -m pylogeny_b200.build >/dev/null 2>&1 || { echo "build failed $v"; continue; }
  touch pylogeny_b200/csrc/lik.cu pylogeny_b200/csrc/nb_api.cpp
  echo "== $v"
  python profiles/t_lnl.py 2>&1 | tail -1
done
PLG_NVCC_EXTRA="-DPLG_USE_VISITS=1 -DPLG_VB_MINB=5" python -m pylogeny_b200.build >/dev/null 2>&1
python -m pytest tests/test_neighbourhood_gpu.py -m gpu -x -q 2>&1 | tail -2
