# clv4_kernel block-size / occupancy sweep (rebuilds the library on the GPU box for each variant)
for v in "128 6" "128 7" "128 8" "64 10" "64 12" "64 14" "256 3" "96 7"; do
  set -- $v
  PLG_NVCC_EXTRA="-DPLG_D4_THREADS=$1 -DPLG_D4_MINB=$2" python -m pylogeny_b200.build >/dev/null 2>&1 || { echo "build failed $v"; continue; }
  touch pylogeny_b200/csrc/lik.cu
  echo "== threads=$1 minblocks=$2"
  python profiles/t_lnl.py 2>&1 | tail -2
done
