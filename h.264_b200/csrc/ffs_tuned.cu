// ffs_tuned.cu -- integer-pel fast full search, the throughput kernel (SAD metric, RDO on, same range for every reference).
//
// Same results as ffs_basic.cu, i.e. as SetupFastFullPelSearch / SetupLargerBlocks / FastFullPelBlockMotionSearch of the
// reference (lencod/src/me_fullfast.c:491, :210, :833), with the work laid out for the B200 integer pipes.
//
//  * Persistent CTAs (3 per SM), each walking the (macroblock, list, ref) jobs blockIdx.x, +gridDim.x, ...  A CTA has
//    4R threads: thread (dx, h) owns ONE horizontal displacement dx in [0, 2R) and ONE 8-pixel column pair h of the
//    macroblock and streams down the (2R+16)-row search window.  Its 32 current-picture words (16 rows x 2 columns) stay
//    in registers; one window row costs 3 shared-memory words + 2 funnel shifts and feeds 32 VABSDIFF4.U8.ACC -- the 16
//    vertical displacements that overlap the row are in flight at once.  Their accumulators are addressed by
//    (band of the macroblock, dy & 3), which makes every register index a compile-time constant with the row loop
//    unrolled only four times: the interior body is ~500 instructions (8 KB) and stays in the instruction caches
//    (a 16x unrolled first version spent 2.3 warps per issue cycle waiting for instructions).  The first and last 16
//    window rows, where some displacements do not exist, are eight compile-time specialised four-row groups, so no
//    SAD and no evaluation is issued outside the (2R+1)^2 positions.
//  * The larger-block SADs are never materialised: key(sum) = a*K + key(b) folds every SetupLargerBlocks add into the
//    IMAD that builds the comparison key (FMA pipe); only the 8x8 sums and the two 16-wide sums (one warp shuffle each
//    between the lanes h = 0 / 1) exist as plain integers.
//  * Cost: key = (SAD + mvcost) * K + spiral_index with K = (2R+1)^2 > any spiral index, so ONE unsigned minimum per
//    (partition, position) reproduces JM's "first strict minimum in spiral order" (me_fullfast.c:879-897).
//    K * (lambda * (bits_x + bits_y) >> 16) comes from a shared table indexed by the sum of the Exp-Golomb length classes
//    of the two vector components: the x class is a per-thread constant (fixed dx), the y class is uniform and read
//    from a per-job table.  Slots of the same band are looked up in PAIRS: a 36 x 36 table of {value(s0), value(s1)}
//    gives both with one IMAD (address) + one LDS.64.  Per (partition, position): IMAD (key) + VIADDMNMX + half a lookup.
//  * The window column dx = 2R (the 65th of 65) does not fit the 2R-lane layout: thread pair i scores position (2R, i)
//    with the same key encoding after the sweep.
//  * Search windows (2R+16)^2 and the 16x16 current macroblock arrive by TMA (cp.async.bulk.tensor.2d, SASS UTMALDG) from
//    the wide-border search plane / the current picture, double-buffered one job ahead on two mbarriers; a window whose
//    box leaves the wide border (predictor drift beyond the 32-pixel guard) is loaded with clamped element reads instead.
//    TMA needs the first byte of a box row 16-byte aligned in global memory (measured on B200: any other start column of a
//    byte tensor raises an illegal-instruction fault), so the box is 16 bytes wider than the window and starts at
//    X0 & ~15; the residue X0 & 15 goes into each thread's word offset and funnel-shift count.
//
// ALU-pipe work per search position (both lanes): 64 VABSDIFF4 + 42 VIADDMNMX + 4 SHF + ~10 adds/selects.
#include <cstdlib>
#include <cstring>
#include <utility>
#include "me_common.cuh"
#include "tma.cuh"

namespace {

constexpr int CYW = 12;            // words per (window row, h) of the y-class table: one per lookup unit
constexpr int LUT_STRIDE_B = 12;   // bytes between table entries: odd word stride -> conflict-free and the address is an IMAD
constexpr int LUT_ENTRIES = 36;    // class sums 0..35 (classes up to 17 per component: |d| < 2^17)
constexpr int NSLOT = 21;          // partitions scored by one thread
// The 21 slots are looked up in 12 units: 9 PAIRS of slots of the same band share one 64-bit load from a table indexed by
// both class sums (36 x 36 entries of {K*mvcost(s0), K*mvcost(s1)}), 3 single slots use the 36-entry table.
constexpr int NUNIT = 12;
__host__ __device__ constexpr int unit_a(int u) { constexpr int t[NUNIT] = { 0, 2, 3, 5, 7, 9, 11, 12, 14, 16, 18, 20 }; return t[u]; }
__host__ __device__ constexpr int unit_b(int u) { constexpr int t[NUNIT] = { 1, -1, 4, 6, 8, 10, -1, 13, 15, 17, 19, -1 }; return t[u]; }
// the band whose completion evaluates the unit; its displacement is dy = window row - (4 * band + 3)
__host__ __device__ constexpr int unit_lag(int u) { return u < 2 ? 3 : u < 5 ? 7 : u < 7 ? 11 : 15; }
__constant__ int8_t c_unit_a[NUNIT] = { 0, 2, 3, 5, 7, 9, 11, 12, 14, 16, 18, 20 };
__constant__ int8_t c_unit_b[NUNIT] = { 1, -1, 4, 6, 8, 10, -1, 13, 15, 17, 19, -1 };
__constant__ int8_t c_unit_lag[NUNIT] = { 3, 3, 7, 7, 7, 11, 11, 15, 15, 15, 15, 15 };

template <int R> struct Geo {
  static constexpr int N = 2 * R + 16;            // window side
  static constexpr int NP = N + 16;               // row pitch in shared memory: the TMA box starts at a 16-byte aligned column
  static constexpr int SIDE = 2 * R + 1;
  static constexpr int NT = 4 * R;                // threads per CTA
  static constexpr int NW = NT / 32;
  static constexpr unsigned KMUL = SIDE * SIDE;   // key = cost * KMUL + spiral index
};

// partition (ABI order, include/me_cuda.h) scored in slot e by the lanes of column pair h
__constant__ int8_t c_slot_part[2][NSLOT] = {
  { 25, 26,  9, 29, 30, 11, 17, 18, 5, 33, 34, 13, 37, 38, 15, 21, 22, 7, 3, 1, 0 },
  { 27, 28, 10, 31, 32, 12, 19, 20, 6, 35, 36, 14, 39, 40, 16, 23, 24, 8, 4, 2, 0 } };

template <int R> struct __align__(128) Smem {
  uint8_t  win[2][Geo<R>::NP * Geo<R>::N];      // search windows, N rows of NP bytes (TMA box, dense); the window starts at byte X0 & 15
  uint32_t cur[2][64];                          // current macroblocks, 16 rows x 4 words
  uint32_t cyt[Geo<R>::N][2][CYW];              // per (window row r, h, unit): y class word (single: class, pair: 36 * class_a +
                                                // class_b) of the displacement dy = r - lag(unit) the unit is evaluated for in row r
  uint2    lut2[LUT_ENTRIES * LUT_ENTRIES];     // {K * mvcost(s0), K * mvcost(s1)} at index 36 * s0 + s1
  uint32_t lut[LUT_ENTRIES * 3];                // K * mvcost(class sum), entry stride 12 bytes
  uint32_t ua_last[2][NUNIT];                   // unit table addresses of column 2R per (h, unit)
  uint32_t red[Geo<R>::NW][2][NSLOT];
  uint32_t pslot[2][24];                        // per (h, slot): predictor word (x | y << 16) of the job about to be searched, 21 used
  uint32_t job[3][48];                          // me_mb_job ring (188 bytes each)
  int      jid[4];                              // job ids of the ring slots (>= n_jobs: no such job), 3 used
  uint16_t sptab[Geo<R>::SIDE * Geo<R>::SIDE];  // spiral index of (dy, dx)
  MeRefDev refs[2 * ME_MAX_REFS];               // copy of the reference table
  unsigned long long mbar[2];
  int lut_lambda, lut_valid;
};

// shared-memory loads at [address register + compile-time offset]: the unrolled sweep addresses everything relative to three
// per-group base registers, so no address arithmetic is issued inside it
template <int OFF = 0> __device__ __forceinline__ unsigned lds32(unsigned a)
{
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
  return v;
}
template <int OFF = 0> __device__ __forceinline__ unsigned lds16(unsigned a)
{
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1+%2];" : "=h"(v) : "r"(a), "n"(OFF));
  return v;
}
template <int OFF = 0> __device__ __forceinline__ uint2 lds64(unsigned a)
{
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(a), "n"(OFF));
  return v;
}
template <int OFF = 0> __device__ __forceinline__ uint4 lds128(unsigned a)
{
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a), "n"(OFF));
  return v;
}
// sum of the four byte |a - b| plus c: one VABSDIFF4.U8.ACC (the intrinsic __vsadu4 leaves the accumulate to a separate add)
__device__ __forceinline__ unsigned vsad(unsigned a, unsigned b, unsigned c)
{
  unsigned d;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// a * m + c on the FMA pipe
__device__ __forceinline__ unsigned imad(unsigned a, unsigned m, unsigned c)
{
  unsigned d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(m), "r"(c));
  return d;
}
__device__ __forceinline__ int len_class(int d)   // (mvbits(d) - 1) / 2: 0 for 0, floor(log2|d|) + 1 otherwise (mv-search.c:333-341)
{
  return 32 - __clz(d < 0 ? -d : d);
}

struct Regs {
  unsigned accL[4][4], accR[4][4];  // [band b][dy & 3]: 4x4 SAD in progress, left / right 4-pixel column of the pair
  unsigned curL[16], curR[16];      // current-picture words of the 16 rows
  unsigned h0L[4], h0R[4];          // band-0 4x4 SADs waiting for band 1, by dy & 3
  unsigned h2L[4], h2R[4];          // band-2 4x4 SADs waiting for band 3
  unsigned hT1[4], hT2[4];          // top 8x8 SAD waiting 8 rows for the bottom one: two-stage queue per dy & 3
  unsigned best[NSLOT];             // running minima of the keys
  unsigned ua[NUNIT];               // per unit: shared address of the table entry of this thread's x class(es), y class 0
};

struct Thr {
  unsigned sp_addr;     // shared address of sptab[0][dx]
  unsigned cy_addr;     // shared address of cyt[0][h][0]
  unsigned shift;       // (dx & 3) * 8
  unsigned k8;          // 8, as a run-time value: keeps the pair-table address an IMAD (FMA pipe) instead of an ALU-pipe LEA
  int h;
};

// single slot E of unit U: one 32-bit lookup
template <int U, int E>
__device__ __forceinline__ void eval1(Regs &rg, unsigned key, unsigned cy)
{
  const unsigned m = lds32(imad(cy, LUT_STRIDE_B, rg.ua[U]));
  rg.best[E] = __viaddmin_u32(key, m, rg.best[E]);
}
// slots EA, EB of pair unit U: one 64-bit lookup
template <int U, int EA, int EB>
__device__ __forceinline__ void eval2(Regs &rg, const Thr &t, unsigned keyA, unsigned keyB, unsigned cyp)
{
  const uint2 m = lds64(imad(cyp, t.k8, rg.ua[U]));
  rg.best[EA] = __viaddmin_u32(keyA, m.x, rg.best[EA]);
  rg.best[EB] = __viaddmin_u32(keyB, m.y, rg.best[EB]);
}

// per 4-row group: shared addresses every load of the group is an immediate offset from
struct GroupBase {
  unsigned win;     // window row r0, this thread's first word
  unsigned sp;      // sptab[r0][dx]
  unsigned cy;      // cyt[r0][h][0]
};

// Band B (rows 4B..4B+3 of the macroblock) of vertical displacement dy = r0 + DYO has just been accumulated: L / Rr are
// the two 4x4 SADs, J = dy & 3, hL / hR the 4x4 SADs of the band above (bands 1, 3).  cA, cB, cC: this row's y-class words.
template <int R, int B, int DYO, int J>
__device__ __forceinline__ void band_done(Regs &rg, const Thr &t, const GroupBase &g, unsigned L, unsigned Rr, unsigned hL, unsigned hR,
                                          const uint4 &cA, const uint4 &cB, const uint4 &cC)
{
  constexpr unsigned K = Geo<R>::KMUL;
  constexpr int SPO = DYO * Geo<R>::SIDE * 2;
  const unsigned sp = lds16<SPO>(g.sp);
  const unsigned kL = imad(L, K, sp), kR = imad(Rr, K, sp);
  const unsigned k84 = imad(L, K, kR);                  // 8x4
  if (B == 0) { eval2<0, 0, 1>(rg, t, kL, kR, cA.x); eval1<1, 2>(rg, k84, cA.y); }
  else if (B == 2) { eval2<5, 9, 10>(rg, t, kL, kR, cB.y); eval1<6, 11>(rg, k84, cB.z); }
  else {
    const unsigned k48L = imad(hL, K, kL), k48R = imad(hR, K, kR);
    const unsigned PL = L + hL;
    const unsigned k88 = imad(PL, K, k48R);             // 8x8
    const unsigned S8 = PL + Rr + hR;
    if (B == 1) {
      eval2<2, 3, 4>(rg, t, kL, kR, cA.z);
      eval2<3, 5, 6>(rg, t, k84, k48L, cA.w);
      eval2<4, 7, 8>(rg, t, k48R, k88, cB.x);
      rg.hT1[J] = S8;
    } else {
      eval2<7, 12, 13>(rg, t, kL, kR, cB.w);
      eval2<8, 14, 15>(rg, t, k84, k48L, cC.x);
      eval2<9, 16, 17>(rg, t, k48R, k88, cC.y);
      const unsigned T = rg.hT2[J], V = T + S8;          // 8x16 of this column pair
      const unsigned kV = imad(V, K, sp);
      const unsigned X = t.h ? T : S8;                   // what the other lane needs for its 16x8
      const unsigned Y = __shfl_xor_sync(0xffffffffu, X, 1);
      const unsigned Vo = __shfl_xor_sync(0xffffffffu, V, 1);
      eval2<10, 18, 19>(rg, t, kV, imad(V - X + Y, K, sp), cC.z);   // 8x16; 16x8 top (h = 0) / bottom (h = 1)
      eval1<11, 20>(rg, imad(Vo, K, kV), cC.w);                     // 16x16 (both lanes, same key)
    }
  }
}

// One window row r = r0 + U (r0 a multiple of 4).  Row r is row k = 4b + i of the macroblock for the displacement
// dy = r - k; with J = dy & 3 the phase is i = (U - J) & 3, so the 32 accumulators are addressed by compile-time (b, J)
// although the loop is unrolled over four rows only (the sweep body stays within the instruction caches).
// The last accumulate of bands 0 / 2 writes the hold register the next band reads four rows later (h0 / h2).
// MODE 0: interior rows 16 .. 2R-1, every displacement exists.  MODE 1: rows 4G .. 4G+3 of the first sixteen, displacements
// dy < 0 do not exist.  MODE 2: rows 2R+4G .. 2R+4G+3 of the last sixteen, displacements dy > 2R do not exist.  The first and
// last groups are specialised at compile time (no SAD and no evaluation is issued for a displacement outside [0, 2R]).
template <int MODE, int G, int U, int B, int I>
__host__ __device__ constexpr bool sad_valid()      // accumulate row k = 4B + I in this row?
{
  return MODE == 0 ? true : MODE == 1 ? (B < G || (B == G && I <= U)) : (B > G || (B == G && I >= U));
}
template <int MODE, int G, int U, int B>
__host__ __device__ constexpr bool band_valid()     // does band B of an existing displacement complete in this row?
{
  return MODE == 0 ? true : MODE == 1 ? (B < G || (B == G && U == 3)) : (B >= G);
}

template <int R, int MODE, int G, int U, int B, int J>
__device__ __forceinline__ void sad_step(Regs &rg, unsigned refL, unsigned refR)
{
  constexpr int i = (U - J) & 3, k = 4 * B + i;
  if constexpr (sad_valid<MODE, G, U, B, i>()) {
    const unsigned vL = vsad(refL, rg.curL[k], i ? rg.accL[B][J] : 0u), vR = vsad(refR, rg.curR[k], i ? rg.accR[B][J] : 0u);
    if (i == 3 && B == 0) { rg.h0L[J] = vL; rg.h0R[J] = vR; }
    else if (i == 3 && B == 2) { rg.h2L[J] = vL; rg.h2R[J] = vR; }
    else { rg.accL[B][J] = vL; rg.accR[B][J] = vR; }
  }
}

template <int R, int MODE, int G, int U>
__device__ __forceinline__ void sweep_row(Regs &rg, const Thr &t, const GroupBase &g)
{
  constexpr int NP = Geo<R>::NP;
  constexpr int J = (U + 1) & 3;      // dy & 3 of the displacements whose bands complete in this row
  const unsigned w0 = lds32<U * NP>(g.win), w1 = lds32<U * NP + 4>(g.win), w2 = lds32<U * NP + 8>(g.win);
  const unsigned refL = __funnelshift_r(w0, w1, t.shift), refR = __funnelshift_r(w1, w2, t.shift);
  // holds of the displacement four rows older, read before this row's completions replace them
  const unsigned h0L = rg.h0L[J], h0R = rg.h0R[J], h2L = rg.h2L[J], h2R = rg.h2R[J];
#define ME_SAD_BAND(B) sad_step<R, MODE, G, U, B, 0>(rg, refL, refR); sad_step<R, MODE, G, U, B, 1>(rg, refL, refR); \
                       sad_step<R, MODE, G, U, B, 2>(rg, refL, refR); sad_step<R, MODE, G, U, B, 3>(rg, refL, refR);
  ME_SAD_BAND(0) ME_SAD_BAND(1) ME_SAD_BAND(2) ME_SAD_BAND(3)
#undef ME_SAD_BAND
  // this row's twelve y-class words (already shifted by each band's lag when the table was built)
  constexpr int CYO = U * 2 * CYW * 4;
  uint4 cA = make_uint4(0, 0, 0, 0), cB = cA, cC = cA;
  if constexpr (band_valid<MODE, G, U, 0>() || band_valid<MODE, G, U, 1>()) cA = lds128<CYO>(g.cy);
  if constexpr (band_valid<MODE, G, U, 1>() || band_valid<MODE, G, U, 2>() || band_valid<MODE, G, U, 3>()) cB = lds128<CYO + 16>(g.cy);
  if constexpr (band_valid<MODE, G, U, 3>()) cC = lds128<CYO + 32>(g.cy);
  // bands in descending order (band 3 reads the 8x8 queue that band 1 of a later displacement advances)
  if constexpr (band_valid<MODE, G, U, 3>()) band_done<R, 3, U - 15, J>(rg, t, g, rg.accL[3][J], rg.accR[3][J], h2L, h2R, cA, cB, cC);
  if constexpr (band_valid<MODE, G, U, 2>()) band_done<R, 2, U - 11, J>(rg, t, g, rg.h2L[J], rg.h2R[J], 0u, 0u, cA, cB, cC);
  rg.hT2[J] = rg.hT1[J];
  if constexpr (band_valid<MODE, G, U, 1>()) band_done<R, 1, U - 7, J>(rg, t, g, rg.accL[1][J], rg.accR[1][J], h0L, h0R, cA, cB, cC);
  if constexpr (band_valid<MODE, G, U, 0>()) band_done<R, 0, U - 3, J>(rg, t, g, rg.h0L[J], rg.h0R[J], 0u, 0u, cA, cB, cC);
}

// four window rows starting at row r0 (a multiple of 4)
template <int R, int MODE, int G>
__device__ __forceinline__ void sweep4(Regs &rg, const Thr &t, unsigned row0_addr, int r0)
{
  GroupBase g;
  g.win = row0_addr + (unsigned)r0 * Geo<R>::NP;
  g.sp = t.sp_addr + (unsigned)r0 * (Geo<R>::SIDE * 2);
  g.cy = t.cy_addr + (unsigned)r0 * (2 * CYW * 4);
  sweep_row<R, MODE, G, 0>(rg, t, g);
  sweep_row<R, MODE, G, 1>(rg, t, g);
  sweep_row<R, MODE, G, 2>(rg, t, g);
  sweep_row<R, MODE, G, 3>(rg, t, g);
}

// unit U with its table address taken from shared memory (column 2R: the same for all threads of one h)
template <int U>
__device__ __forceinline__ void eval_unit_at(Regs &rg, const Thr &t, unsigned keyA, unsigned keyB, unsigned cyrow, unsigned ua_addr)
{
  const unsigned w = lds32(cyrow + unit_lag(U) * (2 * CYW * 4) + U * 4);
  const unsigned ua = lds32(ua_addr + U * 4);
  if (unit_b(U) >= 0) {
    const uint2 m = lds64(imad(w, t.k8, ua));
    rg.best[unit_a(U)] = __viaddmin_u32(keyA, m.x, rg.best[unit_a(U)]);
    rg.best[unit_b(U) >= 0 ? unit_b(U) : 0] = __viaddmin_u32(keyB, m.y, rg.best[unit_b(U) >= 0 ? unit_b(U) : 0]);
  } else {
    const unsigned m = lds32(imad(w, LUT_STRIDE_B, ua));
    rg.best[unit_a(U)] = __viaddmin_u32(keyA, m, rg.best[unit_a(U)]);
  }
}

// one whole position of the last window column (dx = 2R), all 21 slots of this lane.  col_addr: shared address of the first
// word of window row 0 that holds column 2R + 8h; shift: byte residue of the window start, as a funnel-shift count
template <int R>
__device__ __forceinline__ void last_column(Regs &rg, const Thr &t, unsigned col_addr, unsigned shift, unsigned cyt_h_addr,
                                            unsigned ua_addr, unsigned sptab_addr, int dy)
{
  constexpr unsigned K = Geo<R>::KMUL;
  constexpr int NP = Geo<R>::NP, SIDE = Geo<R>::SIDE;
  unsigned L[4] = { 0, 0, 0, 0 }, Q[4] = { 0, 0, 0, 0 };
  const unsigned a0 = col_addr + (unsigned)dy * NP;
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const unsigned w0 = lds32(a0 + k * NP), w1 = lds32(a0 + k * NP + 4), w2 = lds32(a0 + k * NP + 8);
    L[k >> 2] = vsad(__funnelshift_r(w0, w1, shift), rg.curL[k], L[k >> 2]);
    Q[k >> 2] = vsad(__funnelshift_r(w1, w2, shift), rg.curR[k], Q[k >> 2]);
  }
  const unsigned sp = lds16(sptab_addr + (unsigned)(dy * SIDE + 2 * R) * 2);
  const unsigned cyrow = cyt_h_addr + (unsigned)dy * (2 * CYW * 4);
#define KEY(v) imad((v), K, sp)
  const unsigned T = L[0] + L[1] + Q[0] + Q[1], S8 = L[2] + L[3] + Q[2] + Q[3], V = T + S8;
  const unsigned X = t.h ? T : S8;
  const unsigned Y = __shfl_xor_sync(0xffffffffu, X, 1);
  const unsigned Vo = __shfl_xor_sync(0xffffffffu, V, 1);
  eval_unit_at<0>(rg, t, KEY(L[0]), KEY(Q[0]), cyrow, ua_addr);
  eval_unit_at<1>(rg, t, KEY(L[0] + Q[0]), 0u, cyrow, ua_addr);
  eval_unit_at<2>(rg, t, KEY(L[1]), KEY(Q[1]), cyrow, ua_addr);
  eval_unit_at<3>(rg, t, KEY(L[1] + Q[1]), KEY(L[0] + L[1]), cyrow, ua_addr);
  eval_unit_at<4>(rg, t, KEY(Q[0] + Q[1]), KEY(T), cyrow, ua_addr);
  eval_unit_at<5>(rg, t, KEY(L[2]), KEY(Q[2]), cyrow, ua_addr);
  eval_unit_at<6>(rg, t, KEY(L[2] + Q[2]), 0u, cyrow, ua_addr);
  eval_unit_at<7>(rg, t, KEY(L[3]), KEY(Q[3]), cyrow, ua_addr);
  eval_unit_at<8>(rg, t, KEY(L[3] + Q[3]), KEY(L[2] + L[3]), cyrow, ua_addr);
  eval_unit_at<9>(rg, t, KEY(Q[2] + Q[3]), KEY(S8), cyrow, ua_addr);
  eval_unit_at<10>(rg, t, KEY(V), KEY(V - X + Y), cyrow, ua_addr);
  eval_unit_at<11>(rg, t, KEY(V + Vo), 0u, cyrow, ua_addr);
#undef KEY
}

// window tensor maps travel as a kernel parameter (constant bank): (list, ref) -> list * TUNED_MAX_REFS + ref
constexpr int TUNED_MAX_REFS = 8;
struct WinMaps { CUtensorMap m[2 * TUNED_MAX_REFS]; };

struct JobView {            // fields of me_mb_job read from its shared-memory word image
  int opix_x, opix_y, list, ref, pmv16_x, pmv16_y, lambda;
};
__device__ __forceinline__ JobView job_view(const uint32_t *w)
{
  JobView v;
  v.opix_x = (int)(int16_t)(w[0] & 0xffffu); v.opix_y = (int)(int16_t)(w[0] >> 16);
  v.list = (int)(int8_t)(w[1] & 0xffu); v.ref = (int)(int8_t)((w[1] >> 8) & 0xffu);
  v.pmv16_x = (int)(int16_t)(w[2] & 0xffffu); v.pmv16_y = (int)(int16_t)(w[2] >> 16);
  v.lambda = (int)w[44];                          // lambda[0]: byte offset 176
  return v;
}
// pmv[p] = word 3 + p
__device__ __forceinline__ int job_pmvx(const uint32_t *w, int p) { return (int)(int16_t)(w[3 + p] & 0xffffu); }
__device__ __forceinline__ int job_pmvy(const uint32_t *w, int p) { return (int)(int16_t)(w[3 + p] >> 16); }

// search centre (me_fullfast.c:550-566), RDO on
__device__ __forceinline__ void search_centre(const MeCfgDev &c, int R, const JobView &j, int &offx, int &offy)
{
  int x = j.pmv16_x / 4, y = j.pmv16_y / 4;
  x = me_clamp(x, -2047 + R, 2047 - R);
  y = me_clamp(y, c.mv_limit_min_y + R, c.mv_limit_max_y - R);
  offx = x; offy = y;
}

// 3 CTAs per SM at R = 32 (168 registers); 4 CTAs (128 registers, 43 spill stores) measured slower: 2.20 vs 2.01 ms
#ifndef ME_FFS_R32_CTAS
#define ME_FFS_R32_CTAS 3
#endif
template <int R>
__global__ void __launch_bounds__(Geo<R>::NT, (R == 32) ? ME_FFS_R32_CTAS : (R == 16) ? 6 : 1)
me_ffs_tuned_kernel(const uint32_t *__restrict__ jobs, int n_jobs, const MeRefDev *__restrict__ ref_table,
                    const __grid_constant__ WinMaps win_maps, const __grid_constant__ CUtensorMap cur_map,
                    const uint8_t *__restrict__ cur_pic, int cur_pitch, int use_tma,
                    MeCfgDev cfg, MeResultInt *__restrict__ out, int *__restrict__ deferred /* [0] count, [1..] job ids */,
                    int *__restrict__ ticket /* jobs handed out beyond the first two rounds (zeroed by the caller) */)
{
  using G = Geo<R>;
  constexpr int N = G::N, NP = G::NP, SIDE = G::SIDE, NT = G::NT;
  constexpr unsigned K = G::KMUL;
  constexpr int JW = sizeof(MeJob) / 4;   // 47 words
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem<R> &sm = *reinterpret_cast<Smem<R> *>(smem_raw);

  const int tid = threadIdx.x, dx = tid >> 1, lane = tid & 31, warp = tid >> 5;
  Thr t;
  t.h = tid & 1;
  t.k8 = (unsigned)cfg.search_range / (R / 8);     // = 8
  t.sp_addr = smem_u32(&sm.sptab[dx]);
  t.cy_addr = smem_u32(&sm.cyt[0][t.h][0]);
  const unsigned lut_addr = smem_u32(sm.lut);
  // Jobs are handed out dynamically: a CTA's first two jobs are blockIdx.x and blockIdx.x + gridDim.x, every further one is
  // the next ticket of a global counter, drawn two jobs ahead (the descriptor travels one job ahead of the window).  With a
  // static stride the CTAs of an SM finish at different times -- the warp schedulers favour one of the resident warps -- and
  // the last ones run alone at a third of the issue rate (ncu, round 1: 2.6 of 3 resident warps on average).
  if ((int)blockIdx.x >= n_jobs) return;
  if (cfg.abort_flag && *cfg.abort_flag) return;          // a job failed validation (me_check_jobs_kernel): nothing is searched

  // ---- once per CTA: spiral table, barriers, first two job descriptors, first window
  for (int i = tid; i < SIDE * SIDE; i += NT) {
    int y = i / SIDE, x = i - y * SIDE;
    sm.sptab[i] = (uint16_t)me_spiral_index(x - R, y - R);
  }
  if (tid == 0) {
    mbar_init(smem_u32(&sm.mbar[0]), 1);
    mbar_init(smem_u32(&sm.mbar[1]), 1);
    mbar_init_fence();
    sm.lut_valid = 0;
  }
  for (int i = tid; i < (int)(sizeof(sm.refs) / 4); i += NT)
    reinterpret_cast<uint32_t *>(sm.refs)[i] = reinterpret_cast<const uint32_t *>(ref_table)[i];
  if (tid < JW) {
    sm.job[0][tid] = jobs[(size_t)blockIdx.x * JW + tid];
    if ((int)(blockIdx.x + gridDim.x) < n_jobs) sm.job[1][tid] = jobs[((size_t)blockIdx.x + gridDim.x) * JW + tid];
  }
  if (tid == 0) { sm.jid[0] = (int)blockIdx.x; sm.jid[1] = (int)(blockIdx.x + gridDim.x); sm.jid[2] = n_jobs; }
  __syncthreads();
  // predictors in slot order (read with vector loads when a thread derives its table addresses)
  auto fill_pslot = [&](const uint32_t *jwn) {
    if (tid < 2 * NSLOT) { const int hh = tid / NSLOT, e = tid - hh * NSLOT; sm.pslot[hh][e] = jwn[3 + c_slot_part[hh][e]]; }
  };
  fill_pslot(sm.job[0]);
  __syncthreads();

  // window origin in search-plane coordinates; true when the whole box lies inside the plane (TMA path)
  auto window_origin = [&](const JobView &jv, const MeRefDev &rf, int &X0, int &Y0) -> bool {
    int offx, offy;
    search_centre(cfg, R, jv, offx, offy);
    X0 = jv.opix_x + offx - R + rf.spad;
    Y0 = jv.opix_y + offy - R + rf.spad;
    return use_tma && X0 >= 0 && Y0 >= 0 && X0 + N <= rf.sw && Y0 + N <= rf.sh;
  };
  auto issue_loads = [&](int slot, int buf) {      // thread 0 only
    const JobView jv = job_view(sm.job[slot]);
    const int ri = jv.list * ME_MAX_REFS + jv.ref;
    const MeRefDev &rf = sm.refs[ri];
    int X0, Y0;
    const bool tma_ok = window_origin(jv, rf, X0, Y0);
    const unsigned bar = smem_u32(&sm.mbar[buf]);
    proxy_fence();
    mbar_expect_tx(bar, 256u + (tma_ok ? (unsigned)(NP * N) : 0u));
    tma_load_2d(smem_u32(sm.cur[buf]), &cur_map, jv.opix_x, jv.opix_y, bar);
    if (tma_ok) tma_load_2d(smem_u32(sm.win[buf]), &win_maps.m[jv.list * TUNED_MAX_REFS + jv.ref], X0 & ~15, Y0, bar);
  };
  if (tid == 0 && use_tma) issue_loads(0, 0);

  for (int it = 0;; it++) {
    const int buf = it & 1, slot = it % 3, slot1 = (it + 1) % 3, slot2 = (it + 2) % 3;
    const int job_id = sm.jid[slot];
    if (job_id >= n_jobs) break;
    const bool has_next = sm.jid[slot1] < n_jobs;
    const uint32_t *jw = sm.job[slot];
    // prefetch: window + macroblock of the next job (TMA); ticket of the job after that (its descriptor is fetched behind the
    // barrier in front of the sweep and goes registers -> shared at the end)
    if (tid == 32 * (it & (G::NW - 1)) && use_tma && has_next) issue_loads(slot1, buf ^ 1);   // rotating warp
    int tk = 0;
    if (tid == 0) tk = has_next ? atomicAdd(ticket, 1) + 2 * (int)gridDim.x : n_jobs;
    uint32_t pf = 0;

    const JobView jv = job_view(jw);
    if ((unsigned)jv.lambda > (1u << 24)) {
      // lambda too large for 32-bit keys (never the case for H.264 quantisers): leave the job to the kernel of ffs_basic.cu
      if (tid == 0) { deferred[1 + atomicAdd(&deferred[0], 1)] = job_id; sm.jid[slot2] = tk; }
      if (use_tma) mbar_wait(smem_u32(&sm.mbar[buf]), (unsigned)((it >> 1) & 1));
      __syncthreads();
      const int nid2 = sm.jid[slot2];
      if (tid < JW && nid2 < n_jobs) sm.job[slot2][tid] = jobs[(size_t)nid2 * JW + tid];
      if (has_next) fill_pslot(sm.job[slot1]);
      __syncthreads();
      continue;
    }
    const MeRefDev &rf = sm.refs[jv.list * ME_MAX_REFS + jv.ref];
    int offx, offy, X0, Y0;
    search_centre(cfg, R, jv, offx, offy);
    const bool tma_ok = window_origin(jv, rf, X0, Y0);
    const int mis = X0 & 15;                                 // window column 0 sits at byte `mis` of every shared row
    t.shift = ((dx + mis) & 3) * 8;

    // ---- per-job tables
    if (!sm.lut_valid || sm.lut_lambda != jv.lambda) {       // uniform: K * ((lambda * bits) >> 16), bits = 2 * class sum + 2
      __syncthreads();
      if (tid < LUT_ENTRIES) sm.lut[tid * 3] = K * (unsigned)((jv.lambda * (2 * tid + 2)) >> 16);
      for (int i = tid; i < LUT_ENTRIES * LUT_ENTRIES; i += NT) {
        const int s0 = i / LUT_ENTRIES, s1 = i - s0 * LUT_ENTRIES;
        sm.lut2[i] = make_uint2(K * (unsigned)((jv.lambda * (2 * s0 + 2)) >> 16), K * (unsigned)((jv.lambda * (2 * s1 + 2)) >> 16));
      }
      if (tid == 0) { sm.lut_lambda = jv.lambda; sm.lut_valid = 1; }
    }
    {
      // y classes: PER threads per (h, unit) column, interleaved over dy
      constexpr int PER = NT / (2 * NUNIT);
      if (tid < 2 * NUNIT * PER) {
        const int col = tid / PER, first = tid - col * PER, hh = col / NUNIT, u = col - hh * NUNIT;
        const int ea = c_unit_a[u], eb = c_unit_b[u];
        const int pya = job_pmvy(jw, c_slot_part[hh][ea]), pyb = job_pmvy(jw, c_slot_part[hh][eb < 0 ? ea : eb]);
        uint32_t *dst = &sm.cyt[first + c_unit_lag[u]][hh][u];          // row r = dy + lag
        int d = 4 * (offy + first - R);
        for (int dy = first; dy < SIDE; dy += PER) {
          const int ca = len_class(d - pya), cb = len_class(d - pyb);
          *dst = (uint32_t)(eb < 0 ? ca : LUT_ENTRIES * ca + cb);
          d += 4 * PER; dst += PER * 2 * CYW;
        }
      }
    }
    const unsigned lut2_addr = smem_u32(sm.lut2);
    if (tid < 2 * NUNIT) {
      const int hh = tid / NUNIT, u = tid - hh * NUNIT, ea = c_unit_a[u], eb = c_unit_b[u];
      const int ca = len_class(4 * (offx + R) - job_pmvx(jw, c_slot_part[hh][ea]));
      const int cb = len_class(4 * (offx + R) - job_pmvx(jw, c_slot_part[hh][eb < 0 ? ea : eb]));
      sm.ua_last[hh][u] = eb < 0 ? lut_addr + LUT_STRIDE_B * ca : lut2_addr + 8 * (LUT_ENTRIES * ca + cb);
    }
    Regs rg;
#pragma unroll
    for (int e = 0; e < NSLOT; e++) rg.best[e] = 0xffffffffu;
    {
      const uint4 *ps = reinterpret_cast<const uint4 *>(sm.pslot[t.h]);
      uint32_t pw[24];
#pragma unroll
      for (int q = 0; q < 6; q++) { const uint4 v = ps[q]; pw[4 * q] = v.x; pw[4 * q + 1] = v.y; pw[4 * q + 2] = v.z; pw[4 * q + 3] = v.w; }
      const int xq = 4 * (offx + dx - R);
#pragma unroll
      for (int u = 0; u < NUNIT; u++) {
        const int ca = len_class(xq - (int)(int16_t)(pw[unit_a(u)] & 0xffffu));
        if (unit_b(u) >= 0) {
          const int cb = len_class(xq - (int)(int16_t)(pw[unit_b(u) >= 0 ? unit_b(u) : 0] & 0xffffu));
          rg.ua[u] = lut2_addr + 8 * (LUT_ENTRIES * ca + cb);
        } else {
          rg.ua[u] = lut_addr + LUT_STRIDE_B * ca;
        }
      }
    }

    // ---- wait for this job's macroblock + window
    if (use_tma) mbar_wait(smem_u32(&sm.mbar[buf]), (unsigned)((it >> 1) & 1));
    else if (tid < 64)       // diagnostic mode without TMA: plain loads of the current macroblock
      sm.cur[buf][tid] = *reinterpret_cast<const uint32_t *>(cur_pic + (size_t)(jv.opix_y + (tid >> 2)) * cur_pitch + jv.opix_x + 4 * (tid & 3));
    if (!tma_ok) {
      for (int i = tid; i < NP * N; i += NT) {
        int y = i / NP, x = i - y * NP;
        int sy = me_clamp(Y0 + y, 0, rf.sh - 1), sx = me_clamp(X0 - mis + x, 0, rf.sw - 1);
        sm.win[buf][i] = rf.search[(size_t)sy * rf.spitch + sx];
      }
    }
    if (tid == 0) sm.jid[slot2] = tk;
    __syncthreads();          // tables + manual window + next ticket visible
    const int nid2 = sm.jid[slot2];
    if (tid < JW && nid2 < n_jobs) pf = jobs[(size_t)nid2 * JW + tid];
#pragma unroll
    for (int k = 0; k < 16; k++) { rg.curL[k] = sm.cur[buf][4 * k + 2 * t.h]; rg.curR[k] = sm.cur[buf][4 * k + 2 * t.h + 1]; }

    // ---- the sweep
    const unsigned wbase = smem_u32(sm.win[buf]);
    const unsigned row0 = wbase + (unsigned)((dx + mis) >> 2) * 4 + 8 * t.h;
    sweep4<R, 1, 0>(rg, t, row0, 0);
    sweep4<R, 1, 1>(rg, t, row0, 4);
    sweep4<R, 1, 2>(rg, t, row0, 8);
    sweep4<R, 1, 3>(rg, t, row0, 12);
#ifdef ME_FFS_INTERIOR_REPEAT   // timing experiment only (results invalid): the interior rows N times, to read the steady-state rate of the sweep
#pragma unroll 1
    for (int rep = 1; rep < ME_FFS_INTERIOR_REPEAT; rep++)
#pragma unroll 1
      for (int r0 = 16; r0 < 2 * R; r0 += 4) sweep4<R, 0, 0>(rg, t, row0, r0);
#endif
#pragma unroll 1        // four rows per iteration: eight measured slower (2.14 vs 2.02 ms per 1080p picture, instruction cache)
    for (int r0 = 16; r0 < 2 * R; r0 += 4) sweep4<R, 0, 0>(rg, t, row0, r0);
    sweep4<R, 2, 0>(rg, t, row0, 2 * R);
    sweep4<R, 2, 1>(rg, t, row0, 2 * R + 4);
    sweep4<R, 2, 2>(rg, t, row0, 2 * R + 8);
    sweep4<R, 2, 3>(rg, t, row0, 2 * R + 12);

    // ---- last window column: pair i scores (2R, i); (2R, 2R) is scored by one warp (all its pairs, same result)
    {
      const unsigned cyh = smem_u32(&sm.cyt[0][t.h][0]), la = smem_u32(&sm.ua_last[t.h][0]), spa = smem_u32(sm.sptab);
      const unsigned col = wbase + (unsigned)((2 * R + mis) >> 2) * 4 + 8 * t.h, sh_last = (mis & 3) * 8;
      const int reps = (warp == (it & (G::NW - 1))) ? 2 : 1;      // the corner rotates over the sub-partitions
#pragma unroll 1                                                 // one copy of the code (the kernel is instruction-fetch sensitive)
      for (int rep = 0; rep < reps; rep++) last_column<R>(rg, t, col, sh_last, cyh, la, spa, rep ? 2 * R : dx);
    }

    // ---- minima over the CTA: even lanes hold the slots of column pair 0, odd lanes those of column pair 1 -- each half
    // reduces among itself (redux.sync with its own member mask)
    if (t.h) {
#pragma unroll
      for (int e = 0; e < NSLOT - 1; e++) {               // (h = 1, slot 20) duplicates (h = 0, slot 20)
        const unsigned m = __reduce_min_sync(0xaaaaaaaau, rg.best[e]);
        if (lane == 1) sm.red[warp][1][e] = m;
      }
    } else {
#pragma unroll
      for (int e = 0; e < NSLOT; e++) {
        const unsigned m = __reduce_min_sync(0x55555555u, rg.best[e]);
        if (lane == 0) sm.red[warp][0][e] = m;
      }
    }
    __syncthreads();
    if (tid < 2 * NSLOT - 1) {            // (h = 1, slot 20) duplicates (h = 0, slot 20)
      const int hh = tid / NSLOT, e = tid - hh * NSLOT;
      unsigned kmin = sm.red[0][hh][e];
#pragma unroll
      for (int w = 1; w < G::NW; w++) kmin = min(kmin, sm.red[w][hh][e]);
      const unsigned cost = kmin / K, sp = kmin - cost * K;
      int sx, sy;
      me_spiral_xy((int)sp, sx, sy);
      const int p = c_slot_part[hh][e];
      const unsigned mvw = ((unsigned)(offx + sx) & 0xffffu) | ((unsigned)(offy + sy) << 16);
      reinterpret_cast<uint32_t *>(out[job_id].mv)[p] = mvw;
      out[job_id].cost[p] = (int)cost;
    }
    if (tid < JW && nid2 < n_jobs) sm.job[slot2][tid] = pf;
    if (has_next) fill_pslot(sm.job[slot1]);
    __syncthreads();          // window / tables / red free for the next job
  }
}

template <int R>
int launch(const MeJob *jobs, int n_jobs, const MeRefDev *ref_table, const void *win_maps, const void *cur_map,
           const uint8_t *cur_pic, int cur_pitch, const MeCfgDev *cfg, MeResultInt *out, int *deferred, int *ticket, cudaStream_t st)
{
  // SM count and the shared-memory opt-in are per device: cached per device ordinal (the library may be shut down and
  // initialised again on another GPU of the box)
  static int sms_of[64];
  static bool attr_done[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (!attr_done[dev]) {
    if (cudaDeviceGetAttribute(&sms_of[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (cudaFuncSetAttribute(me_ffs_tuned_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<R>)) != cudaSuccess) return 0;
    attr_done[dev] = true;
  }
  const int sms = sms_of[dev];
  int per_sm = (R == 32) ? ME_FFS_R32_CTAS : (R == 16) ? 6 : 1;
  // timing experiments (profiles/r02_ffs_occupancy.md): fewer resident CTAs per SM, enforced with a larger shared-memory request
  static int cta_limit = -1;
  if (cta_limit < 0) { const char *e = getenv("ME_CUDA_FFS_CTAS"); cta_limit = e ? atoi(e) : 0; }
  size_t smem_bytes = sizeof(Smem<R>);
  if (cta_limit > 0 && cta_limit < per_sm) {
    per_sm = cta_limit;
    smem_bytes = (size_t)(220 * 1024 / cta_limit) & ~(size_t)1023;
    if (cudaFuncSetAttribute(me_ffs_tuned_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) return 0;
  }
  int grid = sms * per_sm;
  if (grid > n_jobs) grid = n_jobs;
  CUtensorMap cm;
  memcpy(&cm, cur_map, sizeof cm);
  // host table: [2][ME_MAX_REFS] maps of 128 bytes -> kernel parameter [2][TUNED_MAX_REFS]
  WinMaps wm;
  for (int l = 0; l < 2; l++)
    memcpy(&wm.m[l * TUNED_MAX_REFS], (const unsigned char *)win_maps + (size_t)l * ME_MAX_REFS * 128, sizeof(CUtensorMap) * TUNED_MAX_REFS);
  static int use_tma = -1;
  if (use_tma < 0) { const char *e = getenv("ME_CUDA_FFS_TMA"); use_tma = !(e && e[0] == '0'); }
  me_ffs_tuned_kernel<R><<<grid, Geo<R>::NT, smem_bytes, st>>>(reinterpret_cast<const uint32_t *>(jobs), n_jobs, ref_table, wm, cm,
                                                                    cur_pic, cur_pitch, use_tma, *cfg, out, deferred, ticket);
  return 1;
}

}  // namespace

// A 2-D byte tensor map (row pitch `pitch`, box bw x bh, no swizzle) for cp.async.bulk.tensor; out = 128 bytes, 64-aligned.
extern "C" int me_make_tensor_map_2d(void *out, const void *base, int w, int h, int pitch, int bw, int bh)
{
  const cuuint64_t dims[2] = { (cuuint64_t)w, (cuuint64_t)h }, strides[1] = { (cuuint64_t)pitch };
  const cuuint32_t box[2] = { (cuuint32_t)bw, (cuuint32_t)bh };
  return me_encode_tensor_map(out, base, 2, dims, strides, box);
}

// window height of the tuned kernel for a search range (the TMA box is that many rows of height + 16 bytes), 0 when the
// range has no tuned kernel
extern "C" int me_ffs_tuned_window(int search_range)
{
  return (search_range == 16 || search_range == 32 || search_range == 64) ? 2 * search_range + 16 : 0;
}

// Returns the number of kernels launched; 0 = configuration not covered (the caller runs the kernel of ffs_basic.cu on
// every job).  win_maps: host array [2][ME_MAX_REFS] of 128-byte tensor maps of the listed references' search planes;
// n_refs0 / n_refs1: list sizes.  Jobs the kernel cannot take (lambda factor above 2^24) are appended to deferred[1..],
// count in deferred[0] (zeroed by the caller); the caller then runs the kernel of ffs_basic.cu on that list.  ticket: the
// counter the persistent CTAs draw their jobs from (zeroed by the caller).
extern "C" int me_launch_ffs_batch_tuned(const MeJob *jobs, int n_jobs, const MeRefDev *ref_table, const void *win_maps,
                                         int n_refs0, int n_refs1, const void *cur_map, const uint8_t *cur_pic, int cur_pitch,
                                         const MeCfgDev *cfg, MeResultInt *out, int *deferred, int *ticket, cudaStream_t st)
{
  if (!win_maps || !cur_map || !deferred || !ticket) return 0;
  if (n_refs0 > TUNED_MAX_REFS || n_refs1 > TUNED_MAX_REFS) return 0;
  if (cfg->metric[0] != 0 || !cfg->rdopt || cfg->full_search != 2) return 0;
  if (cfg->chroma.on) return 0;                 // chroma terms (ChromaMEEnable): the general kernel
  switch (cfg->search_range) {
    case 16: return launch<16>(jobs, n_jobs, ref_table, win_maps, cur_map, cur_pic, cur_pitch, cfg, out, deferred, ticket, st);
    case 32: return launch<32>(jobs, n_jobs, ref_table, win_maps, cur_map, cur_pic, cur_pitch, cfg, out, deferred, ticket, st);
    case 64: return launch<64>(jobs, n_jobs, ref_table, win_maps, cur_map, cur_pic, cur_pitch, cfg, out, deferred, ticket, st);
    default: return 0;
  }
}
