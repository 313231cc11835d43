#!/bin/bash
# Development aid: builds libhalmd_b200 variants of force.cu with extra -D flags into scripts/_build/<name>.so
# usage: scripts/build_variants.sh name1 "-DROW2_MINB=5" name2 "-DROW2_MINB=8" ...   (select with HALMD_B200_LIB=...)
set -e
cd "$(dirname "$0")/../halmd_b200/csrc"
make -s >/dev/null
mkdir -p ../../scripts/_build
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O3 -Xptxas -v $flags \
      -c force.cu -o ../../scripts/_build/force_$name.o 2> ../../scripts/_build/force_$name.log
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../scripts/_build/$name.so \
      core.o binning.o order.o neighbour.o ../../scripts/_build/force_$name.o external.o integrate.o observables.o dd.o modules.o -cudart static
  grep -A2 "k_force_row2ILi2ELi2ELi4" ../../scripts/_build/force_$name.log | grep -E "Used" | sed "s/^/$name: /"
done
