#!/bin/bash
# build a variant of libdesc_b200.so with extra -D flags on ONE translation unit:
#     profiles/build_variant.sh <name> <file.cu> <flags...>
# result: desc_b200/libdesc_b200_<name>.so (load it with DESC_B200_LIB=...)
set -e
cd "$(dirname "$0")/../desc_b200/csrc"
name=$1; src=$2; shift 2
base=${src%.cu}
make -j8 >/dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function "$@" -c $src -o build/${base}_v_$name.o
objs=$(ls build/*.o | grep -v "_v_" | grep -v "build/$base.o" | tr '\n' ' ')
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libdesc_b200_$name.so $objs build/${base}_v_$name.o -ldl
echo built ../libdesc_b200_$name.so
