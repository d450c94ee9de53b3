# fitch_level_kernel residency sweep (rebuilds the library on the GPU box for each variant)
for v in 1 3 4 5 6 8; do
  PLG_NVCC_EXTRA="-DPLG_FITCH_MINB=$v" python -m pylogeny_b200.build >/dev/null 2>&1 || { echo "build failed $v"; continue; }
  touch pylogeny_b200/csrc/fitch.cu
  echo "== minblocks(256 thr)=$v"
  python profiles/t_host.py 2>&1 | grep "run " | head -3 | tail -1
done
