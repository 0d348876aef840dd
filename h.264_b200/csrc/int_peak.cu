// int_peak.cu -- on-device measurement of the integer-pipe issue rates the motion search is bounded by.
//
// the roofline of the search kernel is the INT-ALU pipe, and MEASURED_PEAKS.json carries no integer
// figure, so the denominator is measured here with dependency-free chains of the instructions the kernels use:
//   kind 0  VABSDIFF4.U8.ACC   (__vsadu4 with accumulate)        -- 4 pixel |a-b| + add per lane-op
//   kind 1  IADD3              (a + b + c)
//   kind 2  IMAD               (a * b + c)
//   kind 3  VIADDMNMX          (min(a + b, c), __viaddmin_u32)
//   kind 4  LOP3               ((a & b) ^ c)
//   kind 5  SHF (funnel shift)
//   kind 6  VABSDIFF4 and IMAD interleaved 1:1 (do the two pipes dual-issue?)
//   kind 7  VABSDIFF4 and IADD3 interleaved 1:1 (same pipe?)
// Result: lane-operations per second over the whole GPU (warp instructions * 32).
#include <cuda_runtime.h>
#include <stdint.h>

#define CHAINS 8

template <int KIND>
__global__ void __launch_bounds__(256)
int_peak_kernel(uint32_t *out, int iters, uint32_t s1, uint32_t s2)
{
  uint32_t a[CHAINS], x[CHAINS];
#pragma unroll
  for (int k = 0; k < CHAINS; k++) { a[k] = threadIdx.x * 2654435761u + k * 40503u + s1; x[k] = a[k] ^ (s2 * (k + 1)); }
  uint32_t b = s1 | 1u, c = s2;
#pragma unroll 4
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int k = 0; k < CHAINS; k++) {
        if (KIND == 0) a[k] = __vsadu4(x[k], a[k]) + a[k];
        else if (KIND == 1) a[k] = a[k] + b + x[k];
        else if (KIND == 2) a[k] = a[k] * b + x[k];
        else if (KIND == 3) a[k] = __viaddmin_u32(a[k], b, x[k]);
        else if (KIND == 4) a[k] = (a[k] & b) ^ x[k];
        else if (KIND == 5) a[k] = __funnelshift_r(a[k], x[k], b);
        else if (KIND == 6) { if (k & 1) a[k] = a[k] * b + x[k]; else a[k] = __vsadu4(x[k], a[k]) + a[k]; }
        else if (KIND == 7) { if (k & 1) a[k] = a[k] + b + x[k]; else a[k] = __vsadu4(x[k], a[k]) + a[k]; }
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < CHAINS; k++) s += a[k] + x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + c;
}

template <int KIND>
static double run_kind(uint32_t *d_out, int blocks, int iters)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int_peak_kernel<KIND><<<blocks, 256>>>(d_out, iters / 8, 12345u, 777u);   // warm-up
  cudaEventRecord(e0);
  int_peak_kernel<KIND><<<blocks, 256>>>(d_out, iters, 12345u, 777u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  double ops = (double)blocks * 256.0 * (double)iters * 4.0 * CHAINS;
  return ops / (ms * 1e-3);
}

extern "C" int me_cuda_measure_int_peak(int kind, double *lane_ops_per_s)
{
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 2;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int blocks = sms * 8, iters = 4096;
  uint32_t *d_out = nullptr;
  if (cudaMalloc((void **)&d_out, (size_t)blocks * 256 * sizeof(uint32_t)) != cudaSuccess) return 2;
  double r = 0;
  switch (kind) {
    case 0: r = run_kind<0>(d_out, blocks, iters); break;
    case 1: r = run_kind<1>(d_out, blocks, iters); break;
    case 2: r = run_kind<2>(d_out, blocks, iters); break;
    case 3: r = run_kind<3>(d_out, blocks, iters); break;
    case 4: r = run_kind<4>(d_out, blocks, iters); break;
    case 5: r = run_kind<5>(d_out, blocks, iters); break;
    case 6: r = run_kind<6>(d_out, blocks, iters); break;
    case 7: r = run_kind<7>(d_out, blocks, iters); break;
    default: cudaFree(d_out); return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  cudaFree(d_out);
  if (e != cudaSuccess) return 2;
  *lane_ops_per_s = r;
  return 0;
}
