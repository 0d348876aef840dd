// int_peak.cu -- on-device measurement of the integer-pipe issue rates the motion search is bounded by.
//
// the roofline of the search kernel is the INT-ALU pipe, and MEASURED_PEAKS.json carries no integer
// figure, so the denominator is measured here with dependency-free chains of the instructions the kernels use.
// Every operation is an `asm volatile` so that the compiler can neither fold nor reorder it away; 8 independent
// chains per thread hide the 4-cycle ALU latency.  Result: lane-operations per second over the whole GPU
// (warp instructions * 32).  Kinds:
//   0 VABSDIFF4 (+acc)   1 IADD3 (add)        2 IMAD (mad.lo)        3 VIADDMNMX (add+min, DPX)
//   4 LOP3 (xor/and)     5 SHF (funnel shift) 6 VIMNMX (min)         7 PRMT (byte permute)
//   8..14: the same as 1..7 interleaved 1:1 with VABSDIFF4 (do the two instructions issue on different pipes?)
//   15 LDS.32 (shared-memory load, conflict-free) interleaved 1:1 with VABSDIFF4
//   16 VABSDIFF4 + IMAD 1:1 with the operand pattern of the search kernel's sweep: every VABSDIFF4 of a group reads the SAME
//      reference word (operand-reuse cache) and the IMAD has an immediate multiplier, so an instruction fetches at most two
//      registers from the register file -- kind 9 fetches three per instruction (two distinct per chain + a shared one) and is
//      bound by the two register-file banks, not by the pipes (see profiles/r02_int_peak.md)
//   17 the same with VIADDMNMX in place of the IMAD (both on the ALU pipe)
//   18 the arithmetic of one window row of the search kernel's sweep, dependency-free: 32 VABSDIFF4.ACC : 21 add+min : 33 IMAD
//      (immediate multiplier).  What this mix reaches (0.73 instructions per cycle and scheduler on B200, at any occupancy:
//      register-file reads, profiles/r02_ffs_ceiling.md) is the practical ceiling of the sweep, below the sum of the two pipes
#include <cuda_runtime.h>
#include <stdint.h>

#define CHAINS 8

template <int OP>
__device__ __forceinline__ void one_op(uint32_t &a, uint32_t x, uint32_t b)
{
  if (OP == 0) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a) : "r"(x), "r"(b));
  else if (OP == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(x));
  else if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(x));
  else if (OP == 3) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; min.u32 %0, t, %2;}" : "+r"(a) : "r"(b), "r"(x));
  else if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(x), "r"(b));
  else if (OP == 5) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(x), "r"(b));
  else if (OP == 6) asm volatile("min.u32 %0, %0, %1;" : "+r"(a) : "r"(x));
  else if (OP == 7) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(x), "r"(b));
}

template <int KIND>
__global__ void __launch_bounds__(256)
int_peak_kernel(uint32_t *out, int iters, uint32_t s1, uint32_t s2)
{
  __shared__ uint32_t sm[256 * 2];
  uint32_t a[CHAINS], x[CHAINS];
#pragma unroll
  for (int k = 0; k < CHAINS; k++) { a[k] = threadIdx.x * 2654435761u + k * 40503u + s1; x[k] = a[k] ^ (s2 * (k + 1)); }
  sm[threadIdx.x] = a[0]; sm[threadIdx.x + 256] = x[0];
  __syncthreads();
  uint32_t b = s1 | 1u;
  const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(&sm[threadIdx.x]);
#pragma unroll 2
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int k = 0; k < CHAINS; k++) {
        if (KIND < 8) one_op<KIND>(a[k], x[k], b);
        else if (KIND < 15) { if (k & 1) one_op<KIND - 7>(a[k], x[k], b); else one_op<0>(a[k], x[k], b); }
        else {
          if (k & 1) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr + ((k & 2) ? 1024u : 0u))); x[k] ^= v; }
          else one_op<0>(a[k], x[k], b);
        }
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < CHAINS; k++) s += a[k] + x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// kinds 16 / 17: 8 accumulators fed by one shared reference word per group (like refL / refR of the sweep), 8 keys
template <int KIND>
__global__ void __launch_bounds__(256)
int_peak_reuse_kernel(uint32_t *out, int iters, uint32_t s1, uint32_t s2)
{
  uint32_t acc[CHAINS], cur[CHAINS], key[CHAINS];
#pragma unroll
  for (int k = 0; k < CHAINS; k++) { acc[k] = 0; cur[k] = threadIdx.x * 2654435761u + k * 40503u + s1; key[k] = cur[k] ^ s2; }
  uint32_t ref = s1 ^ threadIdx.x, best = 0xffffffffu;
#pragma unroll 2
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int k = 0; k < CHAINS; k++) {
        asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(ref), "r"(cur[k]));
        if (KIND == 16) asm volatile("mad.lo.u32 %0, %0, 4225, %1;" : "+r"(key[k]) : "r"(ref));
        else asm volatile("{.reg .u32 t; add.u32 t, %0, %1; min.u32 %0, t, %2;}" : "+r"(key[k]) : "r"(ref), "r"(best));
      }
      ref += 0x01010101u * (u + 1);
    }
  }
  uint32_t s = best;
#pragma unroll
  for (int k = 0; k < CHAINS; k++) s += acc[k] + key[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// kind 18: one window row per step (tools/probe/pipe_probe.cu, mix 11)
__global__ void __launch_bounds__(128)
int_peak_sweep_mix_kernel(uint32_t *out, int iters, uint32_t s1, uint32_t s2)
{
  uint32_t acc[32], cur[8], key[21];
#pragma unroll
  for (int k = 0; k < 32; k++) acc[k] = k * s1;
#pragma unroll
  for (int k = 0; k < 8; k++) cur[k] = threadIdx.x * 2654435761u + k * 40503u + s1;
#pragma unroll
  for (int k = 0; k < 21; k++) key[k] = cur[k % 8] ^ (s2 + k);
  uint32_t ref = s1 ^ threadIdx.x, best = 0xffffffffu;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int k = 0; k < 16; k++) {
        asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(ref), "r"(cur[k % 8]));
        asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[16 + k]) : "r"(best), "r"(cur[k % 8]));
      }
#pragma unroll
      for (int k = 0; k < 21; k++) {
        asm volatile("mad.lo.u32 %0, %0, 4225, %1;" : "+r"(key[k]) : "r"(ref));
        asm volatile("{.reg .u32 t; add.u32 t, %1, %2; min.u32 %0, t, %0;}" : "+r"(key[(k + 7) % 21]) : "r"(cur[k % 8]), "r"(ref));
      }
#pragma unroll
      for (int k = 0; k < 12; k++) asm volatile("mad.lo.u32 %0, %0, 4225, %1;" : "+r"(key[(k + 3) % 21]) : "r"(best));
      ref += 0x01010101u;
    }
  }
  uint32_t s = best;
#pragma unroll
  for (int k = 0; k < 32; k++) s += acc[k];
#pragma unroll
  for (int k = 0; k < 21; k++) s += key[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
static double run_kind(uint32_t *d_out, int blocks, int iters)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  if (KIND >= 16) int_peak_reuse_kernel<KIND><<<blocks, 256>>>(d_out, iters / 8, 12345u, 777u);
  else int_peak_kernel<KIND><<<blocks, 256>>>(d_out, iters / 8, 12345u, 777u);   // warm-up
  cudaEventRecord(e0);
  if (KIND >= 16) int_peak_reuse_kernel<KIND><<<blocks, 256>>>(d_out, iters, 12345u, 777u);
  else int_peak_kernel<KIND><<<blocks, 256>>>(d_out, iters, 12345u, 777u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  double per_op = (KIND == 3 || KIND == 10) ? 1.0 : 1.0;   // add+min is counted as ONE operation (it fuses to VIADDMNMX when it can)
  double ops = (double)blocks * 256.0 * (double)iters * 4.0 * CHAINS * per_op * (KIND >= 16 ? 2.0 : 1.0);   // kinds 16 / 17: two instructions per chain step
  return ops / (ms * 1e-3);
}

extern "C" int me_cuda_measure_int_peak(int kind, double *lane_ops_per_s)
{
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 2;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int blocks = sms * 8, iters = 2048;
  uint32_t *d_out = nullptr;
  if (cudaMalloc((void **)&d_out, (size_t)blocks * 256 * sizeof(uint32_t)) != cudaSuccess) return 2;
  double r = 0;
  switch (kind) {
#define K(n) case n: r = run_kind<n>(d_out, blocks, iters); break;
    K(0) K(1) K(2) K(3) K(4) K(5) K(6) K(7) K(8) K(9) K(10) K(11) K(12) K(13) K(14) K(15) K(16) K(17)
#undef K
    case 18: {
      // three CTAs of 128 threads per SM like the search kernel; 86 instructions per step and thread, 4 steps per iteration
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      const int b = sms * 3, it = 2048;
      int_peak_sweep_mix_kernel<<<b, 128>>>(d_out, it / 8, 12345u, 777u);
      cudaEventRecord(e0);
      int_peak_sweep_mix_kernel<<<b, 128>>>(d_out, it, 12345u, 777u);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      cudaEventDestroy(e0); cudaEventDestroy(e1);
      r = (double)b * 128.0 * (double)it * 4.0 * 86.0 / (ms * 1e-3);
      break;
    }
    default: cudaFree(d_out); return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  cudaFree(d_out);
  if (e != cudaSuccess) return 2;
  *lane_ops_per_s = r;
  return 0;
}
