#!/bin/bash
# usage: tools/build_variant.sh NAME [nvcc flags, e.g. -DME_FFS_R32_CTAS=4]  -- builds tools/variants/libme_cuda_NAME.so for tools/sweep_variants.sh
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
name=$1; shift
mkdir -p "$ROOT/tools/variants"
cd "$ROOT/h.264_b200/csrc"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -cudart static "$@" \
  me_cuda_host.cu planes.cu ffs_basic.cu ffs_tuned.cu subpel.cu subpel_tuned.cu mc.cu deblock.cu int_peak.cu -o "$ROOT/tools/variants/libme_cuda_$name.so"
echo "built tools/variants/libme_cuda_$name.so"
