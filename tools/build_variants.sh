#!/bin/bash
# tools/build_variants.sh — experiment libraries: the radix-sort TU rebuilt with other tile shapes, linked with the
# stock objects into houdini-gsplat-renderer_b200/_exp/lib_<name>.so (git-ignored, travels with gpurun).
# usage: tools/build_variants.sh name:ITEMS:MINB ...     then: tools/sweep.py --libs name,...
set -e
cd "$(dirname "$0")/../houdini-gsplat-renderer_b200/csrc"
make -j8 >/dev/null
mkdir -p ../_exp _build/exp
NVCC=/usr/local/cuda/bin/nvcc
for v in "$@"; do
  IFS=: read name items minb <<< "$v"
  $NVCC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -ccbin /usr/bin/g++ \
        -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden -Xptxas -v -DGSB_RS_ITEMS=$items -DGSB_RS_MINB=$minb \
        -c radix_sort.cu -o _build/exp/radix_sort_$name.o 2> _build/exp/radix_sort_$name.log
  grep -A1 "os_pass_kernelILi8E" _build/exp/radix_sort_$name.log | grep -E "registers|spill" | head -2
  $NVCC -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -cudart static -o ../_exp/lib_$name.so \
        _build/renderer.o _build/project.o _build/exp/radix_sort_$name.o _build/scan.o _build/binning.o _build/blend.o _build/ingest.o
  echo "built _exp/lib_$name.so (ITEMS=$items MINB=$minb)"
done
