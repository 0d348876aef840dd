// pipe_probe2.cu -- second round of tools/probe/pipe_probe.cu: WHY do VABSDIFF4 (ALU pipe) and IMAD (FMA pipe) not overlap fully?
// Variants of the 1:1 mix: interleave granularity, operand sharing, accumulate operand, and the rate of 2- / 3-register forms.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define V(acc, a, b)      asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc) : "r"(a), "r"(b))
#define V0(d, a, b)       asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(0u))
#define IMADI(k, c)       asm volatile("mad.lo.u32 %0, %0, 4225, %1;" : "+r"(k) : "r"(c))
#define IMAD3(k, m, c)    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(k) : "r"(m), "r"(c))
#define ADD3(k, a, b)     asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t;}" : "+r"(k) : "r"(a), "r"(b))
#define ADD2(k, a)        asm volatile("add.u32 %0, %0, %1;" : "+r"(k) : "r"(a))
#define ADDI(k)           asm volatile("add.u32 %0, %0, 12345;" : "+r"(k))
#define LOPI(k, a)        asm volatile("lop3.b32 %0, %0, %1, 0x0f0f0f0f, 0x96;" : "+r"(k) : "r"(a))
#define LOP3R(k, a, b)    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(k) : "r"(a), "r"(b))
#define FFMA3(f, a, b)    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(a), "f"(b))
#define PRMT(k, a, b)     asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(k) : "r"(a), "r"(b))
#define PRMTI(k, a)       asm volatile("prmt.b32 %0, %0, %1, 0x4240;" : "+r"(k) : "r"(a))
#define VMAX2(k, a)       asm volatile("max.u16x2 %0, %0, %1;" : "+r"(k) : "r"(a))

constexpr int NC = 8;
template <int MIX> __device__ constexpr int ops_per_step()
{
  return (MIX >= 20 && MIX < 30) ? NC : 2 * NC;
}
// 0 V,I interleaved 1  1: VV II  2: 4V 4I  3: 8V 8I  4: interleaved, IMAD addend private (no shared register)  5: V with zero
// accumulate (2 registers) + IMAD  6: V + IMAD 3 registers, all private  7: V + FFMA 3 distinct registers  8: V + IADD3 3 registers
// 9: V + LOP3 immediate (2 registers)    20: IADD3 3 registers alone  21: IADD immediate (1 register)  22: LOP3 2 registers + imm
// 23: LOP3 3 registers  24: PRMT 3 registers  25: PRMT 2 registers + immediate selector  26: packed 16-bit max (2 registers)  27: V zero accumulate alone
template <int MIX>
__global__ void __launch_bounds__(128) probe(uint32_t *out, int iters, uint32_t s1, uint32_t s2)
{
  uint32_t acc[NC], cur[NC], key[NC], prv[NC];
  float f[NC], fa[NC], fb[NC];
#pragma unroll
  for (int k = 0; k < NC; k++) {
    acc[k] = k * s1; cur[k] = threadIdx.x * 2654435761u + k * 40503u + s1; key[k] = cur[k] ^ (s2 + k); prv[k] = key[k] * 7u + s1;
    f[k] = (float)k + s1; fa[k] = 1.0f + 1e-6f * k; fb[k] = (float)(s2 + k);
  }
  uint32_t ref = s1 ^ threadIdx.x;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (MIX == 0 || MIX == 4 || MIX == 5 || MIX == 6 || MIX == 7 || MIX == 8 || MIX == 9) {
#pragma unroll
        for (int k = 0; k < NC; k++) {
          if (MIX == 5) V0(acc[k], acc[k], cur[k]); else V(acc[k], ref, cur[k]);
          if (MIX == 0 || MIX == 5) IMADI(key[k], ref);
          if (MIX == 4) IMADI(key[k], prv[k]);
          if (MIX == 6) IMAD3(key[k], prv[k], prv[(k + 1) % NC]);
          if (MIX == 7) FFMA3(f[k], fa[k], fb[k]);
          if (MIX == 8) ADD3(key[k], prv[k], ref);
          if (MIX == 9) LOPI(key[k], ref);
        }
      } else if (MIX >= 1 && MIX <= 3) {
        constexpr int GR = MIX == 1 ? 2 : MIX == 2 ? 4 : 8;
#pragma unroll
        for (int g = 0; g < NC; g += GR) {
#pragma unroll
          for (int k = g; k < g + GR; k++) V(acc[k], ref, cur[k]);
#pragma unroll
          for (int k = g; k < g + GR; k++) IMADI(key[k], ref);
        }
      } else {
#pragma unroll
        for (int k = 0; k < NC; k++) {
          if (MIX == 20) ADD3(key[k], prv[k], ref);
          if (MIX == 21) ADDI(key[k]);
          if (MIX == 22) LOPI(key[k], ref);
          if (MIX == 23) LOP3R(key[k], prv[k], ref);
          if (MIX == 24) PRMT(key[k], prv[k], ref);
          if (MIX == 25) PRMTI(key[k], ref);
          if (MIX == 26) VMAX2(key[k], ref);
          if (MIX == 27) V0(acc[k], acc[k], cur[k]);
        }
      }
      ref += 0x01010101u;
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < NC; k++) s += acc[k] + key[k] + (uint32_t)f[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MIX>
static void run(const char *name, uint32_t *d_out, int sms, double mhz)
{
  printf("%-64s", name);
  for (int wps : { 1, 2, 4, 8 }) {
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
    printf("  %d:%5.3f", wps, (double)iters * 4.0 * ops_per_step<MIX>() * wps / (ms * 1e-3 * mhz * 1e6));
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
  printf("columns = resident warps per scheduler : warp instructions per cycle and scheduler (%.0f MHz assumed)\n", mhz);
  run<20>("IADD3 3 registers alone", d_out, sms, mhz);
  run<21>("IADD immediate (1 register) alone", d_out, sms, mhz);
  run<22>("LOP3 2 registers + immediate alone", d_out, sms, mhz);
  run<23>("LOP3 3 registers alone", d_out, sms, mhz);
  run<24>("PRMT 3 registers alone", d_out, sms, mhz);
  run<25>("PRMT 2 registers + immediate selector alone", d_out, sms, mhz);
  run<26>("max.u16x2 (2 registers) alone", d_out, sms, mhz);
  run<27>("VABSDIFF4 zero accumulate (2 registers) alone", d_out, sms, mhz);
  run<0>("VABSDIFF4 + IMAD imm 1:1, V I V I", d_out, sms, mhz);
  run<1>("VABSDIFF4 + IMAD imm 1:1, VV II", d_out, sms, mhz);
  run<2>("VABSDIFF4 + IMAD imm 1:1, 4V 4I", d_out, sms, mhz);
  run<3>("VABSDIFF4 + IMAD imm 1:1, 8V 8I", d_out, sms, mhz);
  run<4>("VABSDIFF4 + IMAD imm 1:1, no shared register", d_out, sms, mhz);
  run<5>("VABSDIFF4 zero accumulate + IMAD imm 1:1", d_out, sms, mhz);
  run<6>("VABSDIFF4 + IMAD 3 private registers 1:1", d_out, sms, mhz);
  run<7>("VABSDIFF4 + FFMA 3 distinct registers 1:1", d_out, sms, mhz);
  run<8>("VABSDIFF4 + IADD3 3 registers 1:1", d_out, sms, mhz);
  run<9>("VABSDIFF4 + LOP3 imm 1:1", d_out, sms, mhz);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
