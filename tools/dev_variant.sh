#!/bin/bash
# Build a VARIANT of the library for A/B measurements in one gpurun call:
#   tools/dev_variant.sh NAME FILE.cu "-DSOME_KNOB=1"      -> precise-mrd-mini_b200/lib/variants/libpmrd_b200_NAME.so
# (only FILE.cu is recompiled with the extra flags; every other object is the regular build's).  Use with
#   PMRD_B200_LIB=precise-mrd-mini_b200/lib/variants/libpmrd_b200_NAME.so python bench.py ...
set -e
cd "$(dirname "$0")/../precise-mrd-mini_b200/csrc"
name=$1; file=$2; flags=$3
make -j8 >/dev/null
mkdir -p ../build/variants ../lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O3,-Wall,-Wno-unused-function --expt-relaxed-constexpr $flags -c $file -o ../build/variants/${file%.cu}_$name.o
objs=$(ls ../build/*.o | grep -v "/${file%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/variants/libpmrd_b200_$name.so $objs ../build/variants/${file%.cu}_$name.o
echo built ../lib/variants/libpmrd_b200_$name.so
