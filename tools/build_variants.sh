#!/bin/bash
# Build search-kernel variants into build_variants/ (git-ignored; travels to the GPU box) for A/B runs:
#   tools/build_variants.sh name "-DEMM_SEARCH_THREADS=640" [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/../enzymm_b200/csrc"
mkdir -p ../../build_variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
while [ $# -ge 2 ]; do
  name=$1; extra=$2; shift 2
  d=../../build_variants/$name; mkdir -p $d
  for f in emm_prepare emm_search emm_api; do nvcc $FLAGS $extra -Xptxas -v -c -o $d/$f.o $f.cu 2> $d/$f.ptxas.log & done
  g++ -O3 -std=c++17 -fPIC -pthread -c -o $d/emm_pdb.o emm_pdb.cpp &
  g++ -O3 -std=c++17 -fPIC -c -o $d/emm_tsv.o emm_tsv.cpp &
  wait
  nvcc $ARCH -shared -o ../../build_variants/lib_$name.so $d/*.o -lcudart -lpthread
  grep -A2 "emm_search_kernelILb0ELb1ELb0" $d/emm_search.ptxas.log | tail -2
  echo "built build_variants/lib_$name.so"
done
