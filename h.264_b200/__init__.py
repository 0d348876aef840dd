"""B200-native motion-estimation path for the JM 12.4 H.264 reference encoder (package directory `h.264_b200`).

The directory name is not a Python identifier; import it through `h264_b200_pkg` at the repository root, which
registers this package as module `h264_b200`.

  csrc/      CUDA kernels (sm_100a) + the C ABI host code -> libme_cuda.so   (include/me_cuda.h)
  shim/      the reference-side binding: --wrap shim that routes JM lencod's hot path through libme_cuda.so
  me_cuda.py ctypes binding of the C ABI (tests, bench)
  synth.py   deterministic synthetic YUV clips
  build.py   nvcc build of libme_cuda.so
  mc.py      GetSkipCostMB / GetDirectCost8x8 / BIDPartitionCost on the library's prediction primitives
  gop_shard.py GOP sharding across GPUs: the assignment rule and the one-encoder-process-per-GOP driver
"""
from . import build as _build  # noqa: F401
from .me_cuda import (MeCuda, MeCudaError, JOB_DTYPE, RESULT_DTYPE, BIPRED_JOB_DTYPE, BIPRED_RESULT_DTYPE,  # noqa: F401
                      NUM_PART, PART_TYPE, PART_BLK)
