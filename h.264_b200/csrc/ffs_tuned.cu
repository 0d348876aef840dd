// ffs_tuned.cu -- placeholder until the tuned search kernel lands (returns 0 = "not applicable": caller uses ffs_basic).
#include "me_common.cuh"
extern "C" int me_launch_ffs_batch_tuned(const MeJob *, int, const MeRefDev *, const uint8_t *, int, const MeCfgDev *,
                                         MeResultInt *, cudaStream_t)
{
  return 0;
}
