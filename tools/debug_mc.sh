#!/bin/bash
# debugging aid: one configuration through lencod_cuda with the mode-decision / plane-paging switches one at a time
# usage: ARGS="-d encoder.cfg ..." KIND=fade bash tools/debug_mc.sh
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
T=$(mktemp -d); cp "$ROOT"/oracle/_ref/fixtures/* "$T"/ 2>/dev/null; cd "$T"
python - <<PY
import sys
sys.path.insert(0, "$ROOT")
import h264_b200_pkg
from h264_b200 import synth
synth.write_yuv420("syn.yuv", 176, 144, 7, kind="${KIND:-pan}")
PY
export ME_CUDA_LIB="$ROOT/h.264_b200/libme_cuda.so"
ARGS="${ARGS:--d encoder.cfg}"
"$ROOT"/oracle/_ref/lencod $ARGS > cpu.log 2>&1; md5sum test.264 | cut -c1-32
for envs in "ME_CUDA_MC=0 ME_CUDA_HOST_PLANES=1" "ME_CUDA_MC=1 ME_CUDA_HOST_PLANES=1" "ME_CUDA_MC=0 ME_CUDA_HOST_PLANES=0" "ME_CUDA_MC=1 ME_CUDA_HOST_PLANES=1 ME_CUDA_VERIFY=1"; do
  env $envs "$ROOT"/oracle/_ref/lencod_cuda $ARGS > gpu.log 2> gpu.err
  echo "$envs -> rc $? $(md5sum test.264 | cut -c1-32)"
  grep -c VERIFY gpu.err; grep VERIFY gpu.err | head -6; grep "mode-decision\|verify mode" gpu.err
done
