#!/bin/bash
# Side measurement: the unmodified reference encoder (oracle/_ref/lencod) against the same objects with the hot path routed
# through libme_cuda.so (oracle/_ref/lencod_cuda) on a synthetic CIF clip, 8 pictures (IPPP), SR 32, 5 references -- pictures
# 5..7 search five reference pictures each (BASELINE config 2).  Three runs of the GPU binding:
#   pms   : one device launch per PartitionMotionSearch call (default)
#   block : one launch per block search (ME_CUDA_PMS=0, the round-1 binding)
#   verify: pms with the reference function run next to every call (ME_CUDA_VERIFY=1; mismatch count must be 0)
# Prints the encoder's own "Total ME time" / "Total encoding time" lines and the shim's service report; one JSON line at the end.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
FRAMES=${FRAMES:-8}
T=$(mktemp -d)
cp "$ROOT"/oracle/_ref/fixtures/*.cfg "$T"/
cd "$T"
python - <<PY
import sys
sys.path.insert(0, "$ROOT")
import h264_b200_pkg
from h264_b200 import synth
synth.write_yuv420("syn.yuv", 352, 288, $FRAMES)
PY
# EXTRA="-p MbInterlace=2 -p ReferenceReorder=0 -p PocMemoryManagement=0 -p LevelIDC=40": the same clip as MBAFF frames (any further -p overrides)
ARGS="-d encoder_main.cfg -p InputFile=syn.yuv -p SourceWidth=352 -p SourceHeight=288 -p FramesToBeEncoded=$FRAMES -p FrameSkip=0 -p NumberBFrames=0 -p SearchRange=32 -p NumberReferenceFrames=5 -p LevelIDC=51 $EXTRA"
export ME_CUDA_LIB="$ROOT/h.264_b200/libme_cuda.so"
unset ME_CUDA_DISABLE ME_CUDA_VERIFY ME_CUDA_PMS
"$ROOT"/oracle/_ref/lencod $ARGS > cpu.log 2>&1; cp test.264 cpu.264
"$ROOT"/oracle/_ref/lencod_cuda $ARGS > gpu.log 2> gpu.err; cp test.264 gpu.264
ME_CUDA_PMS=0 "$ROOT"/oracle/_ref/lencod_cuda $ARGS > blk.log 2> blk.err; cp test.264 blk.264
ME_CUDA_VERIFY=1 "$ROOT"/oracle/_ref/lencod_cuda $ARGS > ver.log 2> ver.err
echo "CPU        : $(grep 'Total ME time' cpu.log)  $(grep 'Total encoding time' cpu.log)"
echo "GPU (pms)  : $(grep 'Total ME time' gpu.log)  $(grep 'Total encoding time' gpu.log)"
echo "GPU (block): $(grep 'Total ME time' blk.log)  $(grep 'Total encoding time' blk.log)"
grep 'me_cuda shim' gpu.err
grep 'verify mode' ver.err
cmp cpu.264 gpu.264 && cmp cpu.264 blk.264 && echo "bitstreams identical"
python - <<PY
import json, re
def me(f):
    return float(re.search(r"Total ME time for sequence\s*:\s*([0-9.]+)", open(f).read()).group(1))
def tot(f):
    return float(re.search(r"Total encoding time for the seq\.\s*:\s*([0-9.]+)", open(f).read()).group(1))
ver = re.search(r"verify mode: (\d+) calls checked, (\d+) mismatches", open("ver.err").read())
pms = re.search(r"partition-search launches (\d+), block searches answered from them (\d+), argument mismatches (\d+)", open("gpu.err").read())
print(json.dumps({"clip": "CIF 352x288, $FRAMES pictures IPPP, SR 32, 5 references", "extra_args": "$EXTRA", "me_seconds": {"lencod": me("cpu.log"), "lencod_cuda_pms": me("gpu.log"), "lencod_cuda_per_block": me("blk.log")},
                  "encode_seconds": {"lencod": tot("cpu.log"), "lencod_cuda_pms": tot("gpu.log"), "lencod_cuda_per_block": tot("blk.log")},
                  "me_ratio_pms_over_cpu": round(me("gpu.log") / me("cpu.log"), 4),
                  "verify_calls": int(ver.group(1)), "verify_mismatches": int(ver.group(2)),
                  "pms_launches": int(pms.group(1)), "pms_served": int(pms.group(2)), "pms_missed": int(pms.group(3)),
                  "identical_bitstreams": open("cpu.264", "rb").read() == open("gpu.264", "rb").read()}))
PY
