# clv4_kernel experiments (rebuilds the library on the GPU box for each variant)
for v in "-DPLG_D4_PREFETCH=0" "-DPLG_D4_PREFETCH=1" "-DPLG_D4_PREFETCH=1 -DPLG_D4_MINB=5"; do
  PLG_NVCC_EXTRA="$v" python -m pylogeny_b200.build >/dev/null 2>&1 || { echo "build failed $v"; continue; }
  touch pylogeny_b200/csrc/lik.cu
  echo "== $v"
  python profiles/t_lnl.py 2>&1 | tail -1
done
