// tma_probe.cu -- stand-alone check of the 2-D byte-tile TMA load used by ffs_tuned.cu (development aid, not product code)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <stdint.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int BW, int BH>
__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, uint8_t *out, int *status)
{
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ unsigned long long bar;
  unsigned b = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(BW * BH) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(sm)), "l"(&map), "r"(x), "r"(y), "r"(b) : "memory");
  }
  unsigned done = 0;
  for (unsigned spin = 0; spin < (1u << 22) && !done; spin++) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(b), "r"(0u) : "memory");
  }
  if (threadIdx.x == 0) *status = done ? 1 : -1;
  if (done) for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = sm[i];
}

int main(int argc, char **argv)
{
  int ax = argc > 1 ? atoi(argv[1]) : 0, ay = argc > 2 ? atoi(argv[2]) : 0;
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) { printf("no entry point\n"); return 1; }
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int W = 2112, H = 1280, PITCH = 2176;
  std::vector<uint8_t> h((size_t)PITCH * H);
  for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)((i * 2654435761u) >> 13);
  uint8_t *d, *dout; int *dst;
  cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  cudaMalloc(&dout, 256 * 256); cudaMalloc(&dst, 4);
  int boxes[3][2] = { { 16, 16 }, { 80, 80 }, { 144, 144 } };
  for (int bi = 0; bi < 3; bi++) {
    int bw = boxes[bi][0], bh = boxes[bi][1];
    CUtensorMap m;
    cuuint64_t dims[2] = { (cuuint64_t)W, (cuuint64_t)H };
    cuuint64_t strides[1] = { (cuuint64_t)PITCH };
    cuuint32_t box[2] = { (cuuint32_t)bw, (cuuint32_t)bh };
    cuuint32_t estr[2] = { 1, 1 };
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %dx%d encode rc=%d\n", bw, bh, (int)r);
    if (r) continue;
    int xs[3] = { ax, ax, ax }, ys[3] = { ay, ay + 5, ay + 777 };
    for (int t = 0; t < 3; t++) {
      cudaMemset(dst, 0, 4);
      if (bi == 0) probe<16, 16><<<1, 128, 16 * 16>>>(m, xs[t], ys[t], dout, dst);
      if (bi == 1) probe<80, 80><<<1, 128, 80 * 80>>>(m, xs[t], ys[t], dout, dst);
      if (bi == 2) { cudaFuncSetAttribute(probe<144, 144>, cudaFuncAttributeMaxDynamicSharedMemorySize, 144 * 144); probe<144, 144><<<1, 128, 144 * 144>>>(m, xs[t], ys[t], dout, dst); }
      cudaError_t e = cudaDeviceSynchronize();
      int st = 0; cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost);
      std::vector<uint8_t> o(bw * bh);
      cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int yy = 0; yy < bh; yy++) for (int xx = 0; xx < bw; xx++) bad += o[yy * bw + xx] != h[(size_t)(ys[t] + yy) * PITCH + xs[t] + xx];
      printf("  at (%d,%d): sync=%s status=%d mismatches=%d\n", xs[t], ys[t], cudaGetErrorString(e), st, bad);
      if (e != cudaSuccess) return 2;
    }
  }
  return 0;
}
