#!/bin/bash
# Side measurement: the unmodified reference encoder (oracle/_ref/lencod) against the same objects with the hot path routed
# through libme_cuda.so (oracle/_ref/lencod_cuda) on a synthetic CIF clip, 4 pictures, SR 32, 5 references.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
T=$(mktemp -d)
cp "$ROOT"/oracle/_ref/fixtures/*.cfg "$T"/
cd "$T"
python - <<PY
import sys
sys.path.insert(0, "$ROOT")
import h264_b200_pkg
from h264_b200 import synth
synth.write_yuv420("syn.yuv", 352, 288, 4)
PY
ARGS="-d encoder_main.cfg -p InputFile=syn.yuv -p SourceWidth=352 -p SourceHeight=288 -p FramesToBeEncoded=4 -p FrameSkip=0 -p NumberBFrames=0 -p SearchRange=32 -p NumberReferenceFrames=5 -p LevelIDC=51"
export ME_CUDA_LIB="$ROOT/h.264_b200/libme_cuda.so"
"$ROOT"/oracle/_ref/lencod $ARGS > cpu.log 2>&1; cp test.264 cpu.264
"$ROOT"/oracle/_ref/lencod_cuda $ARGS > gpu.log 2>&1
echo "CPU : $(grep 'Total ME time' cpu.log)  $(grep 'Total encoding time' cpu.log)"
echo "GPU : $(grep 'Total ME time' gpu.log)  $(grep 'Total encoding time' gpu.log)"
cmp cpu.264 test.264 && echo "bitstreams identical"
