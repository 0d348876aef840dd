// pipe_probe.cu -- what bounds a mix of ALU-pipe and FMA-pipe integer instructions on one SM sub-partition of a B200?
//
// The search kernel's sweep is 32 VABSDIFF4 : 21 VIADDMNMX : 33 IMAD per window row.  VABSDIFF4 alone and IMAD alone both issue
// every second cycle (16 lanes per cycle), but the two together reach 0.72 of the sum (me_cuda_measure_int_peak kinds 0, 2, 9).
// This probe runs dependency-free chains of several mixes at 1..8 resident warps per scheduler and prints warp instructions
// per cycle and scheduler (IPC) and lane-operations per second.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o
// tools/probe/pipe_probe tools/probe/pipe_probe.cu ; run on the GPU box.  Not part of the library.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define V(acc, a, b)      asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc) : "r"(a), "r"(b))
#define IMAD3(k, m, c)    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(k) : "r"(m), "r"(c))
#define IMADI(k, c)       asm volatile("mad.lo.u32 %0, %0, 4225, %1;" : "+r"(k) : "r"(c))
#define IMADI2(d, a, c)   asm volatile("mad.lo.u32 %0, %1, 4225, %2;" : "=r"(d) : "r"(a), "r"(c))
#define ADDMIN(k, a, b)   asm volatile("{.reg .u32 t; add.u32 t, %1, %2; min.u32 %0, t, %0;}" : "+r"(k) : "r"(a), "r"(b))
#define MINU(k, a)        asm volatile("min.u32 %0, %0, %1;" : "+r"(k) : "r"(a))
#define ADD2(k, a)        asm volatile("add.u32 %0, %0, %1;" : "+r"(k) : "r"(a))
#define FFMA3(f, a, b)    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(a), "f"(b))
#define FFMAI(f, b)       asm volatile("fma.rn.f32 %0, %0, 0f3F800001, %1;" : "+f"(f) : "f"(b))

constexpr int NC = 8;

// MIX: 0 V  1 IMAD3  2 IMADI  3 V+IMAD3 1:1  4 V+IMADI 1:1  5 V:IMADI 2:1  6 V:IMADI 3:1  7 ADD2+IMADI 1:1  8 V+FFMA3 1:1
//      9 V+FFMAI 1:1  10 MINU+IMADI 1:1  11 sweep mix 32 V : 21 ADDMIN : 33 IMADI (per 86)  12 ADDMIN alone  13 ADDMIN + IMADI 1:1
//      14 V + ADDMIN 3:2 (the sweep's ALU-pipe part alone)   15 sweep mix with the IMADs writing fresh registers (key = f(sad))
template <int MIX> __device__ constexpr int ops_per_step()
{
  return MIX == 0 || MIX == 1 || MIX == 2 || MIX == 12 ? NC : MIX == 5 ? NC + NC / 2 : MIX == 6 ? NC + NC / 3 + 1 : MIX == 11 || MIX == 15 ? 86
         : MIX == 14 ? 53 : 2 * NC;
}

template <int MIX>
__global__ void __launch_bounds__(128) probe(uint32_t *out, int iters, uint32_t s1, uint32_t s2)
{
  uint32_t acc[32], cur[NC], key[24];
  float f[NC];
#pragma unroll
  for (int k = 0; k < 32; k++) acc[k] = k * s1;
#pragma unroll
  for (int k = 0; k < NC; k++) { cur[k] = threadIdx.x * 2654435761u + k * 40503u + s1; f[k] = (float)k + s1; }
#pragma unroll
  for (int k = 0; k < 24; k++) key[k] = cur[k % NC] ^ (s2 + k);
  uint32_t ref = s1 ^ threadIdx.x, best = 0xffffffffu;
  float fb = (float)s2;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (MIX == 11 || MIX == 15 || MIX == 14) {
        // one window row of the sweep: 32 SADs on two reference words, 21 keys, 21 add+min, 12 table addresses
#pragma unroll
        for (int k = 0; k < 16; k++) { V(acc[k], ref, cur[k % NC]); V(acc[16 + k], best, cur[k % NC]); }
#pragma unroll
        for (int k = 0; k < 21; k++) {
          if (MIX == 11) IMADI(key[k], ref);
          if (MIX == 15) IMADI2(key[k], acc[k], ref);
          ADDMIN(key[(k + 7) % 21], cur[k % NC], ref);
        }
        if (MIX != 14) {
#pragma unroll
          for (int k = 0; k < 12; k++) IMADI(key[(k + 3) % 21], best);
        }
      } else {
#pragma unroll
        for (int k = 0; k < NC; k++) {
          if (MIX == 0 || MIX == 3 || MIX == 4 || MIX == 5 || MIX == 6 || MIX == 8 || MIX == 9) V(acc[k], ref, cur[k]);
          if (MIX == 1 || MIX == 3) IMAD3(key[k], cur[k], ref);
          if (MIX == 2 || MIX == 4 || MIX == 7 || MIX == 10 || MIX == 13) IMADI(key[k], ref);
          if (MIX == 5 && (k & 1)) IMADI(key[k], ref);
          if (MIX == 6 && (k % 3) == 0) IMADI(key[k], ref);
          if (MIX == 7) ADD2(acc[k], ref);
          if (MIX == 8) FFMA3(f[k], fb, fb);
          if (MIX == 9) FFMAI(f[k], fb);
          if (MIX == 10) MINU(acc[k], ref);
          if (MIX == 12 || MIX == 13) ADDMIN(acc[k], cur[k], ref);
        }
      }
      ref += 0x01010101u;
    }
  }
  uint32_t s = best;
#pragma unroll
  for (int k = 0; k < 32; k++) s += acc[k];
#pragma unroll
  for (int k = 0; k < 24; k++) s += key[k];
#pragma unroll
  for (int k = 0; k < NC; k++) s += (uint32_t)f[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MIX>
static void run(const char *name, uint32_t *d_out, int sms, double mhz)
{
  printf("%-52s", name);
  for (int wps : { 1, 2, 3, 4, 6, 8 }) {
    // wps warps per scheduler = wps CTAs of 128 threads per SM; the shared-memory request keeps more from becoming resident
    const int smem = (int)(200 * 1024 / wps) & ~1023;
    cudaFuncSetAttribute(probe<MIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 4096, blocks = sms * wps;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MIX><<<blocks, 128, smem>>>(d_out, iters / 8, 12345u, 777u);
    cudaEventRecord(e0);
    probe<MIX><<<blocks, 128, smem>>>(d_out, iters, 12345u, 777u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double winstr = (double)iters * 4.0 * ops_per_step<MIX>() * wps;        // per scheduler
    const double cycles = ms * 1e-3 * mhz * 1e6;
    printf("  %d:%5.3f", wps, winstr / cycles);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  printf("\n");
}

int main()
{
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1000.0;
  uint32_t *d_out;
  cudaMalloc(&d_out, (size_t)sms * 8 * 128 * 4);
  printf("SMs %d, clock %.0f MHz (the figures assume the GPU runs at it); columns = resident warps per scheduler : warp instructions per cycle and scheduler\n", sms, mhz);
  run<0>("VABSDIFF4.ACC (3 registers)", d_out, sms, mhz);
  run<1>("IMAD 3 registers", d_out, sms, mhz);
  run<2>("IMAD immediate multiplier (2 registers)", d_out, sms, mhz);
  run<12>("add+min (VIADDMNMX, 3 registers)", d_out, sms, mhz);
  run<3>("VABSDIFF4 + IMAD 3 registers 1:1", d_out, sms, mhz);
  run<4>("VABSDIFF4 + IMAD immediate 1:1", d_out, sms, mhz);
  run<5>("VABSDIFF4 : IMAD immediate 2:1", d_out, sms, mhz);
  run<6>("VABSDIFF4 : IMAD immediate 8:3", d_out, sms, mhz);
  run<7>("IADD (2 registers) + IMAD immediate 1:1", d_out, sms, mhz);
  run<10>("VIMNMX (2 registers) + IMAD immediate 1:1", d_out, sms, mhz);
  run<13>("add+min + IMAD immediate 1:1", d_out, sms, mhz);
  run<8>("VABSDIFF4 + FFMA 3 registers 1:1", d_out, sms, mhz);
  run<9>("VABSDIFF4 + FFMA immediate 1:1", d_out, sms, mhz);
  run<14>("sweep ALU part: 32 VABSDIFF4 : 21 add+min", d_out, sms, mhz);
  run<11>("sweep mix 32 VABSDIFF4 : 21 add+min : 33 IMAD imm.", d_out, sms, mhz);
  run<15>("sweep mix, keys = IMAD(fresh SAD)", d_out, sms, mhz);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
