#!/bin/bash
# Developer aid: link a copy of libalaya_b200.so whose search.cu is compiled with extra -D flags.
#   scripts/build_variant.sh NAME -DAB_INFLIGHT=2 -DAB_MINBLOCKS=4   ->  alayalite_b200/_variants/libNAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p alayalite_b200/_variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3 --expt-relaxed-constexpr \
  -Iinclude "$@" -c alayalite_b200/csrc/search.cu -o alayalite_b200/_variants/search_$name.o
objs=$(ls alayalite_b200/csrc/_obj/*.o | grep -v search.cu.o)
nvcc -shared -o alayalite_b200/_variants/lib$name.so $objs alayalite_b200/_variants/search_$name.o -lcudart_static -lpthread -ldl -lrt
echo built alayalite_b200/_variants/lib$name.so
