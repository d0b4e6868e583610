#!/bin/bash
# Builds tuning variants of the library (only capi.cu is recompiled; the other objects come from build/obj):
#   bash tools/build_variants.sh name "-DMDCT_F64_STAGES=2 -DMDCT_F64_MAXFT=8 -DMDCT_F64_MINB_FWD=5 -DMDCT_F64_MINB_INV=3" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $defs -c -o build/variants/capi_$name.o mdctgan_b200/csrc/capi.cu
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_$name.so build/variants/capi_$name.o build/obj/nn_capi.o build/obj/train_capi.o
  echo built build/variants/lib_$name.so
done
