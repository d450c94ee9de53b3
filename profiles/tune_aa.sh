for v in 1 2 4; do
  PLG_NVCC_EXTRA="-DPLG_A20_UNROLL=$v" python -m pylogeny_b200.build >/dev/null 2>&1 || { echo "build failed $v"; continue; }
  touch pylogeny_b200/csrc/lik.cu
  echo "== unroll=$v"
  python profiles/t_aa.py 2>&1 | tail -2
done
