// deblock.cu -- in-loop deblocking filter of a coded picture on the device.
//
// Replaces the edge loops of DeblockFrame / DeblockMb (reference lencod/src/loopFilter.c:87-253): EdgeLoopLumaNormal (:529-672)
// and EdgeLoopChromaNormal (:814-921) with the tables ALPHA_TABLE / BETA_TABLE / CLIP_TAB (:42-53), for frame pictures without
// MBAFF, 8-bit samples, 4:0:0 and 4:2:0.  The boundary strengths stay the encoder's own (GetStrengthNormal, :262-392, reads the
// macroblock modes, coded-block flags, references and vectors on the host) and arrive per macroblock with the quantisers and
// the filter offsets (me_deblock_mb).
//
// Order: the standard filters macroblocks in raster order, left edges before top edges, and every macroblock reads samples its
// left and upper neighbours have already filtered -- and the upper-right neighbour's left-edge filter has touched the three
// right-most columns of the macroblock above.  So macroblock (x, y) may start once (x - 1, y) and (x + 1, y - 1) are done: a
// wavefront.  One warp per macroblock row walks its row and spins on the progress counter of the row above (all rows are
// resident: a 4K picture has 135).  Inside a macroblock the eight edges run one after the other; the 16 lines across a luma
// edge (lanes 0..15) and the 2 x 8 lines across the chroma edge that shares its strengths (lanes 16..31) are independent.
#include "me_common.cuh"

namespace {

struct DbMb {                      // me_deblock_mb (include/me_cuda.h)
  uint8_t strength[2][4][16];
  int8_t qp, qpc[2], alpha_off, beta_off;
  uint8_t luma_skip, reserved[10];
};
static_assert(sizeof(DbMb) == sizeof(me_deblock_mb) && sizeof(DbMb) == 144, "me_deblock_mb layout");

__constant__ uint8_t c_alpha[52] = { 0,0,0,0,0,0,0,0,0,0,0,0, 0,0,0,0,4,4,5,6, 7,8,9,10,12,13,15,17, 20,22,25,28,32,36,40,45, 50,56,63,71,80,90,101,113,
                                     127,144,162,182,203,226,255,255 };
__constant__ uint8_t c_beta[52] = { 0,0,0,0,0,0,0,0,0,0,0,0, 0,0,0,0,2,2,2,3, 3,3,3,4,4,4,6,6, 7,7,8,8,9,9,10,10, 11,11,12,12,13,13,14,14, 15,15,16,16,17,17,18,18 };
__constant__ uint8_t c_clip[52][5] = {
  {0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},
  {0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},
  {0,0,0,0,0},{0,0,0,1,1},{0,0,0,1,1},{0,0,0,1,1},{0,0,0,1,1},{0,0,1,1,1},{0,0,1,1,1},{0,1,1,1,1},
  {0,1,1,1,1},{0,1,1,1,1},{0,1,1,1,1},{0,1,1,2,2},{0,1,1,2,2},{0,1,1,2,2},{0,1,1,2,2},{0,1,2,3,3},
  {0,1,2,3,3},{0,2,2,3,3},{0,2,2,4,4},{0,2,3,4,4},{0,2,3,4,4},{0,3,3,5,5},{0,3,4,6,6},{0,3,4,6,6},
  {0,4,5,7,7},{0,4,5,8,8},{0,4,6,9,9},{0,5,7,10,10},{0,6,8,11,11},{0,6,8,13,13},{0,7,10,14,14},{0,8,11,16,16},
  {0,9,12,18,18},{0,10,13,20,20},{0,11,15,23,23},{0,13,17,25,25} };

// samples another warp may have written: read at L2
__device__ __forceinline__ int ld(const uint8_t *p) { return (int)__ldcg(p); }
__device__ __forceinline__ void st(uint8_t *p, int v) { __stcg(p, (uint8_t)v); }

// one line of 8 samples across a luma edge, q points at R0 (EdgeLoopLumaNormal, loopFilter.c:600-668)
__device__ __forceinline__ void luma_line(uint8_t *q, ptrdiff_t inc, int strng, int alpha, int beta, int C0)
{
  const int L0 = ld(q - inc), R0 = ld(q), L1 = ld(q - 2 * inc), R1 = ld(q + inc);
  const int delta = R0 - L0, absdelta = abs(delta);
  if (absdelta >= alpha) return;
  if (!(((abs(R0 - R1) - beta) & (abs(L0 - L1) - beta)) < 0)) return;
  const int L2 = ld(q - 3 * inc), R2 = ld(q + 2 * inc);
  if (strng == 4) {
    const int L3 = ld(q - 4 * inc), R3 = ld(q + 3 * inc);
    const int RL0 = L0 + R0;
    const int small_gap = absdelta < ((alpha >> 2) + 2);
    const int aq = (abs(R0 - R2) < beta) & small_gap, ap = (abs(L0 - L2) < beta) & small_gap;
    if (ap) {
      st(q - 3 * inc, (((L3 + L2) << 1) + L2 + L1 + RL0 + 4) >> 3);
      st(q - 2 * inc, (L2 + L1 + L0 + R0 + 2) >> 2);
      st(q - inc, (R1 + ((L1 + RL0) << 1) + L2 + 4) >> 3);
    } else st(q - inc, ((L1 << 1) + L0 + R1 + 2) >> 2);
    if (aq) {
      st(q, (L1 + ((R1 + RL0) << 1) + R2 + 4) >> 3);
      st(q + inc, (R2 + R0 + R1 + L0 + 2) >> 2);
      st(q + 2 * inc, (((R3 + R2) << 1) + R2 + R1 + RL0 + 4) >> 3);
    } else st(q, ((R1 << 1) + R0 + L1 + 2) >> 2);
  } else {
    const int RL0 = (L0 + R0 + 1) >> 1;
    const int aq = abs(R0 - R2) < beta, ap = abs(L0 - L2) < beta;
    const int c0 = C0 + ap + aq;
    const int dif = me_clamp(((delta << 2) + (L1 - R1) + 4) >> 3, -c0, c0);
    if (ap) st(q - 2 * inc, L1 + me_clamp((L2 + RL0 - (L1 << 1)) >> 1, -C0, C0));
    st(q - inc, me_clamp(L0 + dif, 0, 255));
    st(q, me_clamp(R0 - dif, 0, 255));
    if (aq) st(q + inc, R1 + me_clamp((R2 + RL0 - (R1 << 1)) >> 1, -C0, C0));
  }
}

// one line of 4 samples across a chroma edge (EdgeLoopChromaNormal, loopFilter.c:885-915)
__device__ __forceinline__ void chroma_line(uint8_t *q, ptrdiff_t inc, int strng, int alpha, int beta, int C0)
{
  const int L0 = ld(q - inc), R0 = ld(q), L1 = ld(q - 2 * inc), R1 = ld(q + inc);
  const int delta = R0 - L0;
  if (abs(delta) >= alpha) return;
  if (!(((abs(R0 - R1) - beta) & (abs(L0 - L1) - beta)) < 0)) return;
  if (strng == 4) {
    st(q, ((R1 << 1) + R0 + L1 + 2) >> 2);
    st(q - inc, ((L1 << 1) + L0 + R1 + 2) >> 2);
  } else {
    const int c0 = C0 + 1;
    const int dif = me_clamp(((delta << 2) + (L1 - R1) + 4) >> 3, -c0, c0);
    st(q - inc, me_clamp(L0 + dif, 0, 255));
    st(q, me_clamp(R0 - dif, 0, 255));
  }
}

__global__ void __launch_bounds__(32)
me_deblock_kernel(uint8_t *__restrict__ Y, int ypitch, uint8_t *__restrict__ U, uint8_t *__restrict__ V, int cpitch,
                  int mbw, int mbh, const DbMb *__restrict__ mbs, int *progress /* [mbh], zeroed */)
{
  __shared__ DbMb rec;
  const int row = blockIdx.x, lane = threadIdx.x;
  for (int mx = 0; mx < mbw; mx++) {
    const int addr = row * mbw + mx;
    // the upper-right neighbour (or the end of the row above) must be done
    if (row > 0) {
      const int need = mx + 2 < mbw ? mx + 2 : mbw;
      if (lane == 0) { while (*((volatile int *)&progress[row - 1]) < need) { } }
      __syncwarp();
      __threadfence();
    }
    for (int i = lane; i < (int)(sizeof(DbMb) / 4); i += 32) reinterpret_cast<uint32_t *>(&rec)[i] = reinterpret_cast<const uint32_t *>(&mbs[addr])[i];
    __syncwarp();
    for (int dir = 0; dir < 2; dir++) {
#pragma unroll 1
      for (int e = 0; e < 4; e++) {
        // the macroblock on the other side of the edge: the left / upper neighbour for edge 0, this one otherwise
        const DbMb *P = &rec;
        int qpP = rec.qp, qpcP0 = rec.qpc[0], qpcP1 = rec.qpc[1];
        if (e == 0) {
          const int pa = dir ? addr - mbw : addr - 1;
          if ((dir ? row > 0 : mx > 0)) { P = &mbs[pa]; qpP = P->qp; qpcP0 = P->qpc[0]; qpcP1 = P->qpc[1]; }
        }
        if (lane < 16) {
          const int strng = rec.strength[dir][e][lane];
          if (strng && !((rec.luma_skip >> (4 * dir + e)) & 1)) {
            const int QP = (qpP + rec.qp + 1) >> 1;
            const int ia = me_clamp(QP + rec.alpha_off, 0, 51), ib = me_clamp(QP + rec.beta_off, 0, 51);
            uint8_t *q = dir ? Y + (size_t)(16 * row + 4 * e) * ypitch + 16 * mx + lane
                             : Y + (size_t)(16 * row + lane) * ypitch + 16 * mx + 4 * e;
            luma_line(q, dir ? ypitch : 1, strng, c_alpha[ia], c_beta[ib], c_clip[ia][strng]);
          }
        } else if (U && !(e & 1)) {
          // 4:2:0: luma edges 0 and 2 are chroma edges 0 and 4 (chroma_edge, loopFilter.c:55-64); eight lines per component,
          // strength of line pel = Strength[((pel >> 1) << 2) + (pel & 1)] (:880)
          const int uv = (lane - 16) >> 3, pel = (lane - 16) & 7;
          const int strng = rec.strength[dir][e][((pel >> 1) << 2) + (pel & 1)];
          if (strng) {
            const int QP = ((uv ? qpcP1 : qpcP0) + rec.qpc[uv] + 1) >> 1;
            const int ia = me_clamp(QP + rec.alpha_off, 0, 51), ib = me_clamp(QP + rec.beta_off, 0, 51);
            uint8_t *C = uv ? V : U;
            uint8_t *q = dir ? C + (size_t)(8 * row + 2 * e) * cpitch + 8 * mx + pel
                             : C + (size_t)(8 * row + pel) * cpitch + 8 * mx + 2 * e;
            chroma_line(q, dir ? cpitch : 1, strng, c_alpha[ia], c_beta[ib], c_clip[ia][strng]);
          }
        }
        __syncwarp();
        (void)P;
      }
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) *((volatile int *)&progress[row]) = mx + 1;
  }
}

}  // namespace

extern "C" void me_launch_deblock(uint8_t *Y, int ypitch, uint8_t *U, uint8_t *V, int cpitch, int mbw, int mbh, const void *mbs, int *progress,
                                  cudaStream_t st)
{
  cudaMemsetAsync(progress, 0, sizeof(int) * (size_t)mbh, st);
  me_deblock_kernel<<<mbh, 32, 0, st>>>(Y, ypitch, U, V, cpitch, mbw, mbh, (const DbMb *)mbs, progress);
}
