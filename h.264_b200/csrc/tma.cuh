// tma.cuh -- mbarrier + TMA (cp.async.bulk.tensor) helpers shared by the tuned kernels, and the host-side tensor-map encoder.
//
// Measured on B200 (tools/probe/tma_probe.cu): a box of a byte tensor must start at a 16-byte aligned column -- any other
// start coordinate raises an illegal-instruction fault.  Callers therefore load boxes that start at x & ~15 and keep the
// residue x & 15 as an offset into the shared-memory tile.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstring>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy accesses to a buffer before an async-proxy (TMA) write to it
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a transfer that never completes (bad tensor map) traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
#pragma unroll 1
  for (unsigned spin = 0; spin < (1u << 24); spin++) {
    unsigned done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const void *tmap, int x, int y, unsigned bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const void *tmap, int x, int y, int z, unsigned bar)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}

// Byte tensor map of rank 2 or 3 (no swizzle, no interleave); dims / box innermost first, strides in bytes for dims 1..rank-1.
// out: 128 bytes, 64-byte aligned.  0 on success.
static inline int me_encode_tensor_map(void *out, const void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides,
                                       const cuuint32_t *box)
{
  typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) return 1;
    encode = (EncodeTiledFn)fn;
  }
  CUtensorMap m;
  const cuuint32_t estr[3] = { 1, 1, 1 };
  if (encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, (cuuint32_t)rank, const_cast<void *>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 2;
  memcpy(out, &m, sizeof m);
  return 0;
}
