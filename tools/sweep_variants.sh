#!/bin/bash
# usage: tools/sweep_variants.sh TAG name1 name2 ...  -- bench.py once per prebuilt tools/variants/libme_cuda_<name>.so (timing experiments only)
tag=$1; shift
cp h.264_b200/libme_cuda.so /tmp/libme_cuda_orig.so
for v in "$@"; do
  cp tools/variants/libme_cuda_$v.so h.264_b200/libme_cuda.so
  python bench.py --no-cpu-baseline --steps 30 --warmup 3 > gpurun_out/${tag}_$v.log 2>&1
  echo "$v rc=$? $(grep -o '"share_of_step": {[^}]*}' gpurun_out/${tag}_$v.log) $(grep -o '"frac": [0-9.]*' gpurun_out/${tag}_$v.log | head -1)"
done
cp /tmp/libme_cuda_orig.so h.264_b200/libme_cuda.so
