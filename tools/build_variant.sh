#!/bin/bash
# Build a variant of libzb200.so with other compile-time constants in zb_inflate.cu, for A/B runs on
# the GPU box:  tools/build_variant.sh NAME -DZB_INF_CTAS_PER_SM=5 -DZB_ROUND_LG_MAX=4
#   -> tools/_variants/libzb200_NAME.so   (python tools/prof_legs.py reads $ZB_LIB)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p tools/_variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr --expt-extended-lambda \
  -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v -I include "$@" -c zlib_wasm_b200/csrc/zb_inflate.cu -o tools/_variants/zb_inflate_$name.o 2>&1 | grep -A2 "14inflate_kernel" | grep "Used\|spill"
objs=$(ls zlib_wasm_b200/build/*.o | grep -v zb_inflate.cu.o)
nvcc -shared -o tools/_variants/libzb200_$name.so $objs tools/_variants/zb_inflate_$name.o -cudart static -Xlinker -Bsymbolic -Xlinker --exclude-libs,ALL -lpthread -ldl -lrt
