#!/bin/bash
# Build tuning variants of the library locally (nvcc cross-compiles) into variants/<name>.so;
# run one on the GPU box with PLG_LIB_PATH=variants/<name>.so.  usage: build_variants.sh name "flags" [name "flags" ...]
set -e
mkdir -p variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( PLG_LIB_PATH=$PWD/variants/$name.so PLG_NVCC_EXTRA="$flags" python -c "from pylogeny_b200 import build; build.build(force=True)" && echo "built $name: $flags" ) &
done
wait
