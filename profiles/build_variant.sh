#!/bin/bash
# build a variant of libdesc_b200.so with extra -D flags on pgd.cu only:  profiles/build_variant.sh <name> <flags...>
# result: desc_b200/libdesc_b200_<name>.so (load it with DESC_B200_LIB=...)
set -e
cd "$(dirname "$0")/../desc_b200/csrc"
name=$1; shift
make -j8 >/dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function "$@" -c pgd.cu -o build/pgd_$name.o
objs=$(ls build/*.o | grep -v "build/pgd" | tr '\n' ' ')
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libdesc_b200_$name.so $objs build/pgd_$name.o -ldl
echo built ../libdesc_b200_$name.so
