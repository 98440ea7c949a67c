#!/bin/bash
# Dev utility: build tuning variants of libcosmob200.so into variants/lib_<name>.so (run tools/variants.py on the GPU).
# usage: tools/build_variants.sh name1="-DFLAG=.. -DFLAG2=.." name2="..."
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/variants"
rm -f "$ROOT"/variants/lib_*.so "$ROOT"/variants/*.log
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  ( cd "$ROOT/cosmomc_b200/csrc" && /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
      -Xcompiler -fPIC,-O2,-ffp-contract=off -ccbin /usr/bin/g++ --fmad=true -Xptxas -v $flags -shared \
      -o "$ROOT/variants/lib_$name.so" cosmob200.cu > "$ROOT/variants/$name.log" 2>&1 \
      && echo "built $name: $(grep -A3 'project4_kernelILi11ELb0ELi6E' "$ROOT/variants/$name.log" | grep -E 'spill' | head -1)" \
      || echo "FAILED $name (see variants/$name.log)" ) &
  while [ "$(jobs -r | wc -l)" -ge 6 ]; do sleep 1; done
done
wait
