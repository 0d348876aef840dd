// subpel_tuned.cu -- half / quarter-pel refinement, the throughput kernel (Hadamard SATD 4x4 / 8x8, SAD, SSE).
//
// Same results as me_subpel_batch_kernel (subpel.cu), i.e. as SubPelBlockMotionSearch (reference
// lencod/src/me_fullsearch.c:341-511) with computeSATD / HadamardSAD4x4 (lencod/src/me_distortion.c:657, :182), for
// MEDistortionHPel = MEDistortionQPel = 2; with Transform8x8Mode the block types 1..4 use HadamardSAD8x8 (me_distortion.c:272,
// test8x8transform, mv-search.c:640).  MEDistortionHPel / QPel = 0 / 1 (computeSAD me_distortion.c:351, computeSSE :500)
// run in a second instantiation (HAD = false) on the same lanes, so that the Hadamard instantiation keeps its compile-time
// metric; the sub-pel chroma terms of those two metrics stay in subpel.cu.
//
//  * Persistent CTAs of 128 threads (7 per SM), one (macroblock, list, ref) job at a time.
//  * The reference samples of a job come from ONE TMA box: a 48 x 33 pixel region of all 16 quarter-pel planes
//    (3-D tensor map x, y, plane; cp.async.bulk.tensor.3d, 25 KB) around the macroblock displaced by the integer vector of
//    its 16x16 partition.  It covers every candidate of every partition whose integer vector lies within +-8 pixels of
//    that vector and whose 4x4 tiles are not touched by the origin clamp of UMVLine4X (refbuf.c:37); a partition for
//    which this does not hold (checked exactly) reads its tiles from global memory with the clamp instead.  The
//    region of the next job is requested as soon as the last tile of the current one has been read.
//  * A thread owns one transform tile (4x4, or 8x8 for block types 1..4 with Transform8x8Mode) of one block type for the
//    whole job -- its current-picture lanes stay in registers -- and scores it candidate after candidate (9 half-pel, 8
//    quarter-pel).  Lanes are ordered (type, tile) with the tiles of one partition adjacent, so a
//    partition's SATD is a segmented butterfly sum over 1..16 lanes -- no atomics -- and the lane that ends up with the sum
//    adds the motion-vector rate and stores the cost; one thread per partition then replays the reference's sequential
//    first-strict-minimum scan (me_fullsearch.c:423-449, :482-502).
//  * The Hadamard transform runs on pairs of 16-bit lanes: pixels (0,2) / (1,3) of a row share a register, every value
//    carries a bias of 4096 so that lanes never go negative and plain 32-bit IADD3 (with a negated operand and the bias
//    as immediate) is an exact two-lane add / subtract.  Rows first (lane-parallel), the last, lane-crossing stage is
//    folded into the absolute sum: |u+v| + |u-v| with two VABSDIFF-accumulate per register.
#include <cstdlib>
#include <cstring>
#include "me_common.cuh"
#include "tma.cuh"

namespace {

constexpr int RD = 8;                       // integer vectors within +-RD of the 16x16 vector use the staged region
constexpr int RW = 48;                      // region row pitch (bytes): 2*RD + 17 pixels + up to 15 bytes of alignment
constexpr int RH = 2 * RD + 17;             // region rows
constexpr int REGION_BYTES = RW * RH * 16;  // all 16 planes
constexpr int NT = 128;
constexpr int T4_CTAS = 7;                  // CTAs per SM of the 4x4-transform variant (72 registers); 8 CTAs at 64 registers (no spills) measured 0.685 ms, 6 CTAs at 80 registers 0.607 ms, instead of 0.584
constexpr int T8_CTAS = 7;                  // CTAs per SM of the 8x8-transform variant (72 registers, no spills; measured 4: 1.11, 5: 1.01, 6: 0.95 ms)
constexpr unsigned BIAS = 0x1000u, BIAS2 = 0x10001000u;       // 4x4 Hadamard: |coefficient| <= 16 * 255 < 4096
constexpr unsigned BIAS82 = 0x40004000u;                      // 8x8 Hadamard: bias 0x4000 per lane, |coefficient before the last stage| <= 32 * 255 < 16384

// slot q of a (type, candidate) group of 16 lanes -> raster tile 4*ty+tx / partition number (ABI order); the tiles of a
// partition are adjacent and aligned to the partition's tile count
__constant__ uint8_t c_tile[7][16] = {
  {  0,  1,  2,  3,  4,  5,  6,  7,  8,  9, 10, 11, 12, 13, 14, 15 },
  {  0,  1,  2,  3,  4,  5,  6,  7,  8,  9, 10, 11, 12, 13, 14, 15 },
  {  0,  1,  4,  5,  8,  9, 12, 13,  2,  3,  6,  7, 10, 11, 14, 15 },
  {  0,  1,  4,  5,  2,  3,  6,  7,  8,  9, 12, 13, 10, 11, 14, 15 },
  {  0,  1,  2,  3,  4,  5,  6,  7,  8,  9, 10, 11, 12, 13, 14, 15 },
  {  0,  4,  1,  5,  2,  6,  3,  7,  8, 12,  9, 13, 10, 14, 11, 15 },
  {  0,  1,  2,  3,  4,  5,  6,  7,  8,  9, 10, 11, 12, 13, 14, 15 } };
__constant__ uint8_t c_part[7][16] = {
  {  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0 },
  {  1,  1,  1,  1,  1,  1,  1,  1,  2,  2,  2,  2,  2,  2,  2,  2 },
  {  3,  3,  3,  3,  3,  3,  3,  3,  4,  4,  4,  4,  4,  4,  4,  4 },
  {  5,  5,  5,  5,  6,  6,  6,  6,  7,  7,  7,  7,  8,  8,  8,  8 },
  {  9,  9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16 },
  { 17, 17, 18, 18, 19, 19, 20, 20, 21, 21, 22, 22, 23, 23, 24, 24 },
  { 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40 } };
__constant__ uint8_t c_group[7] = { 16, 8, 8, 4, 2, 2, 1 };     // tiles per partition
// the same for the 8x8 tiles (quadrants 2*qy+qx) of block types 1..4
__constant__ uint8_t c_quad[4][4] = { { 0, 1, 2, 3 }, { 0, 1, 2, 3 }, { 0, 2, 1, 3 }, { 0, 1, 2, 3 } };
__constant__ uint8_t c_part8[4][4] = { { 0, 0, 0, 0 }, { 1, 1, 2, 2 }, { 3, 3, 4, 4 }, { 5, 6, 7, 8 } };
__constant__ uint8_t c_group8[4] = { 4, 2, 2, 1 };
// spiral entries 0..8 (mv-search.c:366-393 with range 1)
__constant__ int8_t c_spx[9] = { 0, 0, 0, -1, 1, -1, 1, -1, 1 };
__constant__ int8_t c_spy[9] = { 0, -1, 1, -1, -1, 0, 0, 1, 1 };

template <bool T8> struct __align__(128) Smem {
  uint8_t  region[REGION_BYTES + 128];      // [plane][RH][RW]
  uint32_t curpk[16][8];                    // per tile: rows 0..3 x {(c0 + c1 | c2 + c3), (c0 - c1 | c2 - c3)} as biased 16-bit lanes
  uint32_t cur8[T8 ? 4 : 1][8][4];          // per 8x8 quadrant and row: (c0+c1 | c2+c3), (c0-c1 | c2-c3), (c4+c5 | c6+c7), (c4-c5 | c6-c7), bias BIAS8
  uint32_t job[48];                         // me_mb_job
  int      cost[T8 ? 9 : 1][9];             // 8x8-transform partitions (block types 1..4): cost per candidate
  int      mv[ME_NUM_PARTITIONS][2];        // current best vector, quarter pel
  int      mn[ME_NUM_PARTITIONS];           // its cost
  int      okp[ME_NUM_PARTITIONS];          // partition served by the staged region (else: global reads with the clamp)
  int      mvc[T8 ? 9 : 1][9];              // their motion-vector rate per candidate
  int      next_id;                         // the job after the next one (>= n_jobs: none)
  unsigned long long mbar;
  // the NEXT job, copied in with cp.async while the current one is scored (no global-memory latency between two jobs)
  uint32_t stg_job[48];                     // its descriptor
  uint32_t stg_mv[ME_NUM_PARTITIONS];       // integer-pel vectors of its partitions
  int      stg_cost[ME_NUM_PARTITIONS];     // and their costs
  uint32_t stg_cur[NT];                     // its macroblock, one word per (tile, row, even / odd) thread
  uint32_t stg_cur8[T8 ? NT : 1];           // and per (quadrant, row, lane pair) thread
  MeRefDev refs[2 * 8];                     // copy of the reference table, [list][ref < TUNED_MAX_REFS]
};

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// two-lane (16-bit, biased) add / subtract: one IADD3 each
__device__ __forceinline__ unsigned padd(unsigned a, unsigned b) { return a + b - BIAS2; }
__device__ __forceinline__ unsigned psub(unsigned a, unsigned b) { return a - b + BIAS2; }

// SATD of one 4x4 tile (HadamardSAD4x4, me_distortion.c:182-264): ref rows r[y] (4 bytes each); the current rows arrive with
// the first horizontal stage already applied (job-invariant): cs = (c0 + c1 | c2 + c3), ct = (c0 - c1 | c2 - c3), biased lanes
__device__ __forceinline__ int satd4x4(const unsigned r[4], const uint4 c01, const uint4 c23)
{
  const unsigned cs[4] = { c01.x, c01.z, c23.x, c23.z }, ct[4] = { c01.y, c01.w, c23.y, c23.w };
  unsigned s[4], t[4];
#pragma unroll
  for (int y = 0; y < 4; y++) {
    const unsigned re = __byte_perm(r[y], 0, 0x4240), ro = __byte_perm(r[y], 0, 0x4341);      // (r0 | r2), (r1 | r3)
    s[y] = cs[y] - re - ro;                                        // (d0 + d1 | d2 + d3), d = c - r, bias B
    t[y] = ct[y] - re + ro;                                        // (d0 - d1 | d2 - d3)
  }
  // after the vertical stages a register holds (u + B | v + B) and the last stage would give u + v and u - v:
  // |u + v| + |u - v| = 2 max(|u|, |v|), so the SATD ((sum + 1) >> 1) is the sum of max(|u|, |v|) over the eight registers --
  // packed 16-bit maxima (VIMNMX.U16x2): B + |x| = max(x + B, B - x), then the two lanes of a register against each other
  unsigned accp = 0;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const unsigned *x = k ? t : s;
    const unsigned a = padd(x[0], x[1]), b = psub(x[0], x[1]), c = padd(x[2], x[3]), d = psub(x[2], x[3]);
    const unsigned y4[4] = { padd(a, c), psub(a, c), padd(b, d), psub(b, d) };
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
      const unsigned mi = __vmaxu2(y4[i], 2 * BIAS2 - y4[i]), mj = __vmaxu2(y4[i + 1], 2 * BIAS2 - y4[i + 1]);
      accp += __vmaxu2(__byte_perm(mi, mj, 0x5410), __byte_perm(mi, mj, 0x7632));       // (B + max_i | B + max_j)
    }
  }
  return (int)((accp & 0xffffu) + (accp >> 16)) - 8 * (int)BIAS;   // four values per lane: <= 4 * (4096 + 4080) < 65536
}

// SAD / squared error of one 4x4 tile (the tile's share of computeSAD / computeSSE, me_distortion.c:351, :1042): the current rows
// come back from cs / ct as plain byte words (cs + ct = twice the even pixels, cs - ct = twice the odd ones)
__device__ __forceinline__ int sadsse4x4(const unsigned r[4], const uint4 c01, const uint4 c23, int sse)
{
  const unsigned cs[4] = { c01.x, c01.z, c23.x, c23.z }, ct[4] = { c01.y, c01.w, c23.y, c23.w };
  int acc = 0;
#pragma unroll
  for (int y = 0; y < 4; y++) {
    const unsigned ce = (cs[y] + ct[y] - 2 * BIAS2) >> 1, co = (cs[y] - ct[y]) >> 1;      // (c0 | c2), (c1 | c3): the doubled lanes are even
    const unsigned cw = __byte_perm(ce, co, 0x6240);                            // c0 | c1 << 8 | c2 << 16 | c3 << 24
    if (!sse) acc = (int)__vsadu4(r[y], cw) + acc;
    else {
#pragma unroll
      for (int k = 0; k < 4; k++) { const int d = (int)((r[y] >> (8 * k)) & 255u) - (int)((cw >> (8 * k)) & 255u); acc += d * d; }
    }
  }
  return acc;
}

__device__ __forceinline__ unsigned padd8(unsigned a, unsigned b) { return a + b - BIAS82; }
__device__ __forceinline__ unsigned psub8(unsigned a, unsigned b) { return a - b + BIAS82; }

// SATD of one 8x8 tile (HadamardSAD8x8, me_distortion.c:272-343): ref rows as two words (pixels 0..3, 4..7); the current rows
// come from shared memory with the first horizontal stage applied, four biased lane pairs per row: (c0 + c1 | c2 + c3),
// (c0 - c1 | c2 - c3), (c4 + c5 | c6 + c7), (c4 - c5 | c6 - c7).  Index bits 0 and 2 of the row transform pair whole registers,
// bit 1 pairs the two lanes of a register and is folded into the absolute sum like in the 4x4 case (packed maxima).
__device__ __forceinline__ int satd8x8(const unsigned rl[8], const unsigned rr[8], const uint32_t *curp)
{
  unsigned x[8][4];
#pragma unroll
  for (int y = 0; y < 8; y++) {
    const uint4 c = *reinterpret_cast<const uint4 *>(curp + 4 * y);
    const unsigned le = __byte_perm(rl[y], 0, 0x4240), lo = __byte_perm(rl[y], 0, 0x4341);
    const unsigned re = __byte_perm(rr[y], 0, 0x4240), ro = __byte_perm(rr[y], 0, 0x4341);
    const unsigned a0 = c.x - le - lo, a1 = c.y - le + lo, a2 = c.z - re - ro, a3 = c.w - re + ro;
    x[y][0] = padd8(a0, a2); x[y][1] = psub8(a0, a2); x[y][2] = padd8(a1, a3); x[y][3] = psub8(a1, a3);
  }
  unsigned acc = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    unsigned v[8];
#pragma unroll
    for (int y = 0; y < 8; y++) v[y] = x[y][j];
#pragma unroll
    for (int st = 1; st < 8; st <<= 1)
#pragma unroll
      for (int y = 0; y < 8; y++)
        if (!(y & st)) { const unsigned a = v[y], b = v[y + st]; v[y] = padd8(a, b); v[y + st] = psub8(a, b); }
    // |u + v| + |u - v| = 2 max(|u|, |v|) per register (u + B8 | v + B8): B8 + |x| = max(x + B8, B8 - x), then lane against lane
    unsigned accp = 0;                                             // four values <= 16320 per lane
#pragma unroll
    for (int y = 0; y < 8; y += 2) {
      const unsigned mi = __vmaxu2(v[y], 2 * BIAS82 - v[y]), mj = __vmaxu2(v[y + 1], 2 * BIAS82 - v[y + 1]);
      accp += __vmaxu2(__byte_perm(mi, mj, 0x5410), __byte_perm(mi, mj, 0x7632)) - BIAS82;
    }
    acc += (accp & 0xffffu) + (accp >> 16);
  }
  return (int)((2 * acc + 2) >> 2);                                // acc = half the reference's sum of absolute values
}

struct JobView { int opix_x, opix_y, list, ref, slice_type; };
__device__ __forceinline__ JobView job_view(const uint32_t *w)
{
  JobView v;
  v.opix_x = (int)(int16_t)(w[0] & 0xffffu); v.opix_y = (int)(int16_t)(w[0] >> 16);
  v.list = (int)(int8_t)(w[1] & 0xffu); v.ref = (int)(int8_t)((w[1] >> 8) & 0xffu); v.slice_type = (int)(int8_t)((w[1] >> 16) & 0xffu);
  return v;
}

constexpr int TUNED_MAX_REFS = 8;
struct PlaneMaps { CUtensorMap m[2 * TUNED_MAX_REFS]; };

// T8: Transform8x8Mode (block types 1..4 scored with the 8x8 Hadamard transform).  HAD: both sub-pel metrics are the Hadamard SAD
// (the shipped configurations; the metric is then a compile-time constant and the SAD / SSE code does not exist)
template <bool T8, bool HAD>
__global__ void __launch_bounds__(NT, T8 ? T8_CTAS : T4_CTAS)      // measured (4x4 transform): 6 CTAs 0.775 ms, 7 CTAs 0.736 ms, 8 CTAs (64 registers) 0.777 ms
me_subpel_tuned_kernel(const uint32_t *__restrict__ jobs, int n_jobs, const MeRefDev *__restrict__ ref_table,
                       const __grid_constant__ PlaneMaps maps, const uint8_t *__restrict__ cur, int cur_pitch, MeCfgDev cfg,
                       const MeResultInt *__restrict__ in_int, MeResult *__restrict__ out,
                       int *__restrict__ ticket /* jobs handed out beyond the first round (zeroed by the caller) */,
                       unsigned long long *__restrict__ stats /* [0] partitions served by the staged region, [1] all partitions; may be null */)
{
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem<T8> &sm = *reinterpret_cast<Smem<T8> *>(smem_raw);
  constexpr int JW = sizeof(MeJob) / 4;
  const int tid = threadIdx.x;
  // Jobs are handed out dynamically (first job blockIdx.x, then tickets of a global counter, see ffs_tuned.cu), drawn TWO jobs
  // ahead: while job n is scored, the descriptor, integer-pel results and macroblock of job n + 1 travel into a staging area
  // with cp.async, so that between two jobs there is no global-memory round trip -- one barrier, the region request (TMA),
  // the staging area unpacked into the working arrays, a second barrier.
  if ((int)blockIdx.x >= n_jobs) return;
  if (cfg.abort_flag && *cfg.abort_flag) return;          // a job failed validation (me_check_jobs_kernel): nothing is searched
  const unsigned bar = smem_u32(&sm.mbar);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init_fence();
    sm.next_id = atomicAdd(ticket, 1) + (int)gridDim.x;
  }
  for (int i = tid; i < 2 * TUNED_MAX_REFS * (int)(sizeof(MeRefDev) / 4); i += NT) {
    const int e = i / (int)(sizeof(MeRefDev) / 4), w = i - e * (int)(sizeof(MeRefDev) / 4);
    reinterpret_cast<uint32_t *>(&sm.refs[e])[w] =
        reinterpret_cast<const uint32_t *>(&ref_table[(e / TUNED_MAX_REFS) * ME_MAX_REFS + (e % TUNED_MAX_REFS)])[w];
  }

  // region origin of a job: its 16x16 partition's integer vector, minus the margin, x aligned down to 16 bytes (TMA)
  auto region_origin = [&](uint32_t job_w0, uint32_t mv0_w, int &Xr, int &Yr) {
    const int ox = (int)(int16_t)(job_w0 & 0xffffu), oy = (int)(int16_t)(job_w0 >> 16);
    const int mx = (int)(int16_t)(mv0_w & 0xffffu), my = (int)(int16_t)(mv0_w >> 16);
    Xr = (ox + ME_PAD + mx - RD - 1) & ~15;
    Yr = oy + ME_PAD + my - RD - 1;
  };
  // asynchronous copies of a job into the staging area.  Part A needs the job's number only; part B (the macroblock's
  // pixels) its origin, word 0 of the descriptor
  auto stage_a = [&](int id) {
    if (tid < JW) cp_async4(&sm.stg_job[tid], jobs + (size_t)id * JW + tid);
    if (tid < ME_NUM_PARTITIONS) {
      cp_async4(&sm.stg_mv[tid], reinterpret_cast<const uint32_t *>(in_int[id].mv) + tid);
      cp_async4(&sm.stg_cost[tid], in_int[id].cost + tid);
    }
  };
  auto stage_b = [&](uint32_t w0) {
    const int ox = (int)(int16_t)(w0 & 0xffffu), oy = (int)(int16_t)(w0 >> 16);
    const int tile = tid >> 3, y = (tid & 7) >> 1;          // thread -> (tile, row, even / odd): both threads of a row copy its word
    cp_async4(&sm.stg_cur[tid], cur + (size_t)(oy + 4 * (tile >> 2) + y) * cur_pitch + ox + 4 * (tile & 3));
    if (T8) {
      const int quad = tid >> 5, y8 = (tid >> 2) & 7, j = tid & 3;      // thread -> (quadrant, row, lane pair)
      cp_async4(&sm.stg_cur8[tid], cur + (size_t)(oy + 8 * (quad >> 1) + y8) * cur_pitch + ox + 8 * (quad & 1) + 4 * (j >> 1));
    }
  };
  // thread 0, behind the barrier that ends a job: the staged job's region
  auto issue_region = [&]() {
    const uint32_t w1 = sm.stg_job[1];
    int Xr, Yr;
    region_origin(sm.stg_job[0], sm.stg_mv[0], Xr, Yr);
    const int list = (int)(int8_t)(w1 & 0xffu), ref = (int)(int8_t)((w1 >> 8) & 0xffu);
    proxy_fence();
    mbar_expect_tx(bar, REGION_BYTES);
    tma_load_3d(smem_u32(sm.region), &maps.m[list * TUNED_MAX_REFS + ref], Xr, Yr, 0, bar);
  };
  // staging area -> working arrays (every thread unpacks what the job set-up used to fetch from global memory)
  auto commit = [&]() {
    if (tid < JW) sm.job[tid] = sm.stg_job[tid];
    const int ox = (int)(int16_t)(sm.stg_job[0] & 0xffffu), oy = (int)(int16_t)(sm.stg_job[0] >> 16);
    if (tid < ME_NUM_PARTITIONS) {
      const uint32_t w = sm.stg_mv[tid], w0 = sm.stg_mv[0], w1 = sm.stg_job[1];
      const int mix = (int)(int16_t)(w & 0xffffu), miy = (int)(int16_t)(w >> 16);
      const int m0x = (int)(int16_t)(w0 & 0xffffu), m0y = (int)(int16_t)(w0 >> 16);
      sm.mv[tid][0] = mix * 4;                                                 // mv <<= 2, mv-search.c:770-774
      sm.mv[tid][1] = miy * 4;
      sm.mn[tid] = cfg.start_hp ? sm.stg_cost[tid] : ME_INT_MAX;               // mv-search.c:785-788
      // does the staged region serve every tile of every candidate of the partition?  Its integer vector must lie within
      // +-RD of the 16x16 one, and no tile origin outside [0, size_pad] (the clamp of UMVLine4X would move it)
      const MeRefDev &rf = sm.refs[(int)(int8_t)(w1 & 0xffu) * TUNED_MAX_REFS + (int)(int8_t)((w1 >> 8) & 0xffu)];
      const int t = me_part_type(tid), blk = me_part_blk(tid);
      const int x0 = ox + ME_PAD + 4 * (blk & 3) + mix, y0 = oy + ME_PAD + 4 * (blk >> 2) + miy;   // first tile, centre
      const int x1 = x0 + me_bsx(t) - 4, y1 = y0 + me_bsy(t) - 4;                                  // last tile
      const int ok = abs(mix - m0x) <= RD && abs(miy - m0y) <= RD &&
                     x0 - 1 >= 0 && y0 - 1 >= 0 && x1 <= rf.w + 2 * ME_PAD - 1 - 16 && y1 <= rf.h + 2 * ME_PAD - 1 - 16;
      sm.okp[tid] = ok;
      if (stats) {                                   // share of the partitions the staged region serves (bench.py reports it)
        const unsigned b = __ballot_sync(__activemask(), ok);
        if ((tid & 31) == 0) { atomicAdd(&stats[0], (unsigned long long)__popc(b)); if (tid == 0) atomicAdd(&stats[1], (unsigned long long)ME_NUM_PARTITIONS); }
      }
    }
    {
      // current macroblock as biased lane pairs: thread -> (tile, row, even/odd)
      const int tile = tid >> 3, k = tid & 7, odd = k & 1;
      const uint32_t w = sm.stg_cur[tid];
      const unsigned ce = w & 0x00ff00ffu, co = (w >> 8) & 0x00ff00ffu;                   // (c0 | c2 << 16), (c1 | c3 << 16)
      sm.curpk[tile][k] = (odd ? ce - co : ce + co) + BIAS2;       // cs / ct of satd4x4
      if (T8) {
        const int quad = tid >> 5, y8 = (tid >> 2) & 7, j = tid & 3;
        const uint32_t w8 = sm.stg_cur8[T8 ? tid : 0];
        const unsigned e8 = w8 & 0x00ff00ffu, o8 = (w8 >> 8) & 0x00ff00ffu;
        sm.cur8[quad][y8][j] = ((j & 1) ? e8 - o8 : e8 + o8) + BIAS82;      // first horizontal stage, see satd8x8
      }
    }
  };

  // ---- the first job arrives the same way, synchronously
  int job_id = (int)blockIdx.x;
  stage_a(job_id);
  stage_b(jobs[(size_t)job_id * JW]);
  cp_async_wait_all();
  __syncthreads();
  if (tid == 0) issue_region();
  commit();
  __syncthreads();
  int next_id = sm.next_id;

  for (int it = 0; job_id < n_jobs; it++) {
    // the job after the next one (ticket), and the next one on its way into the staging area
    int tk = n_jobs;
    uint32_t w0n = 0;
    if (next_id < n_jobs) {
      if (tid == 0) tk = atomicAdd(ticket, 1) + (int)gridDim.x;
      stage_a(next_id);
      w0n = jobs[(size_t)next_id * JW];
    }
    const JobView jv = job_view(sm.job);
    const MeRefDev &ref = sm.refs[jv.list * TUNED_MAX_REFS + jv.ref];
    int Xr, Yr;
    region_origin(sm.job[0], ((unsigned)(sm.mv[0][0] >> 2) & 0xffffu) | ((unsigned)(sm.mv[0][1] >> 2) << 16), Xr, Yr);
    // ---- this thread's tile for the whole job.  4x4: slot = (block type, tile slot), 16 lanes per type, the tiles of a
    // partition adjacent (c_tile / c_part).  With the 8x8 transform only the types 5..7 have 4x4 tiles (threads 0..47 of the
    // first two warps) and the last two warps hold the 8x8 tiles of types 1..4 (c_quad / c_part8), four lane groups of 16.
    struct Slot4 { int p, g, x4, y4, rot, bx, by; bool ok, leader; uint4 c01, c23; } s4;
    struct Slot8 { int p, g, x4, y4, quad, grp; bool ok, leader; } s8;
    {
      const int nslot = T8 ? 3 * 16 : 7 * 16;
      const int slot = tid < nslot ? tid : 0, ti = (T8 ? 4 : 0) + (slot >> 4), q = slot & 15, tile = c_tile[ti][q];
      s4.p = c_part[ti][q]; s4.g = c_group[ti];
      s4.leader = tid < nslot && (q & (s4.g - 1)) == 0;
      s4.ok = sm.okp[s4.p] != 0;
      s4.x4 = (jv.opix_x + 4 * (tile & 3) + ME_PAD) << 2; s4.y4 = (jv.opix_y + 4 * (tile >> 2) + ME_PAD) << 2;
      // tiles of the lower macroblock half take their rows in the order 2, 3, 0, 1 (the Hadamard sum does not change under a
      // cyclic row shift): the lanes of a partition then hit four different bank groups instead of two
      // offset of the tile inside its partition: SAD / SSE blocks are clamped at the block origin, not per tile (me_distortion.c:365)
      s4.bx = 4 * ((tile & 3) - (me_part_blk(s4.p) & 3)); s4.by = 4 * ((tile >> 2) - (me_part_blk(s4.p) >> 2));
      s4.rot = (tile >> 2) & 2;
      s4.c01 = *reinterpret_cast<const uint4 *>(&sm.curpk[tile][s4.rot ? 4 : 0]);
      s4.c23 = *reinterpret_cast<const uint4 *>(&sm.curpk[tile][s4.rot ? 0 : 4]);
      const int l8 = (tid - 64) & 63, q16 = l8 & 15, t8 = q16 >> 2, qs = q16 & 3;
      s8.grp = l8 >> 4; s8.p = c_part8[t8][qs]; s8.g = c_group8[t8]; s8.quad = c_quad[t8][qs];
      s8.leader = (qs & (s8.g - 1)) == 0;
      s8.ok = sm.okp[s8.p] != 0;
      s8.x4 = (jv.opix_x + 8 * (s8.quad & 1) + ME_PAD) << 2; s8.y4 = (jv.opix_y + 8 * (s8.quad >> 1) + ME_PAD) << 2;
    }
    const unsigned region_addr = smem_u32(sm.region);
    mbar_wait(bar, (unsigned)(it & 1));

    for (int stage = 0; stage < 2; stage++) {                // 0: half pel (step 2), 1: quarter pel (step 1)
      const int step = stage == 0 ? 2 : 1;
      const int lam = (int)sm.job[44 + 1 + stage];          // lambda[H_PEL], lambda[Q_PEL]
      const int metric = HAD ? 2 : (stage == 0 ? cfg.metric[1] : cfg.metric[2]);      // 2 Hadamard SAD, 0 SAD, 1 squared error (uniform)
      const int first = stage == 0 ? cfg.start_hp : cfg.start_qp;
      if (T8) {
        // motion-vector rate of every candidate of the 8x8-transform partitions (block types 1..4), scored across lane groups
        for (int i = tid; i < 9 * 9; i += NT) {
          const int p = i / 9, c = i - 9 * p;
          sm.mvc[p][c] = me_mv_cost(lam, sm.mv[p][0] + step * c_spx[c], sm.mv[p][1] + step * c_spy[c],
                                    (int)(int16_t)(sm.job[3 + p] & 0xffffu), (int)(int16_t)(sm.job[3 + p] >> 16));
        }
        __syncthreads();
      }
      if (!T8 || tid < 64) {
        // ---- 4x4 tiles: this thread's tile, candidate after candidate
        // The lane that ends up with a partition's SATD also replays the reference's sequential scan over the candidates
        // (first strict minimum, me_fullsearch.c:423-449, :482-502) as they come, in order.
        const int mvpx = sm.mv[s4.p][0], mvpy = sm.mv[s4.p][1];
        const int pmx = (int)(int16_t)(sm.job[3 + s4.p] & 0xffffu), pmy = (int)(int16_t)(sm.job[3 + s4.p] >> 16);
        // The nine candidates are a 3 x 3 grid (offsets -step, 0, +step per component): everything that depends on one
        // component only is computed three times per stage, not nine -- lambda * code length (MV_COST_SMP, defines.h:126),
        // and the two halves of the region address.  Address of sample (x4, y4) of the staged region:
        //   region + [(x4 & 3) * RW*RH + (x4 >> 2) - Xr] + [(y4 & 3) * 4 * RW*RH + ((y4 >> 2) - Yr) * RW]
        // and the second bracket is a multiple of 4, so the word address and the funnel-shift count come from x alone.
        int lbx[3], lby[3];
        unsigned axw[3], shx[3], ayr[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const int ox = (k - 1) * step;
          lbx[k] = lam * me_mvbits(mvpx + ox - pmx);
          lby[k] = lam * me_mvbits(mvpy + ox - pmy);
          const int x4 = s4.x4 + mvpx + ox, y4 = s4.y4 + mvpy + ox;
          const unsigned ax = (unsigned)((x4 & 3) * (RW * RH) + ((x4 >> 2) - Xr));
          axw[k] = region_addr + (ax & ~3u); shx[k] = (ax & 3u) * 8;
          ayr[k] = (unsigned)((y4 & 3) * (4 * RW * RH) + ((y4 >> 2) - Yr + s4.rot) * RW);
        }
        const int half_delta = (2 - 2 * s4.rot) * RW;       // rows 2, 3 of the (rotated) tile relative to rows 0, 1
        int mn = (stage == 1 && !cfg.start_qp) ? ME_INT_MAX : sm.mn[s4.p];       // me_fullsearch.c:456-457
        int best = 0;
        const bool check0 = stage == 0 && !cfg.rdopt && jv.slice_type != ME_B_SLICE && jv.ref == 0 && s4.p == 0 && mvpx == 0 && mvpy == 0;
        const int gmax = __shfl_sync(0xffffffffu, s4.g, 0);   // largest partition (in tiles) among this warp's two block types
        // does this lane's partition take part in the butterfly stage with distance 1, 2, 4, 8?  As 0 / 1 factors: the add becomes
        // one multiply-add on the FMA pipe instead of a compare and a select on the (busier) ALU pipe
        const int gk[4] = { 1 < s4.g, 2 < s4.g, 4 < s4.g, 8 < s4.g };
#pragma unroll
        for (int c = 0; c < 9; c++) {
          if (c == 0 && first) continue;                      // start_me_refinement_hp / qp = 1: the centre keeps its cost
          // spiral entries 0..8 (c_spx / c_spy) as compile-time grid indices
          constexpr int SX[9] = { 1, 1, 1, 0, 2, 0, 2, 0, 2 }, SY[9] = { 1, 0, 2, 0, 0, 1, 1, 2, 2 };
          const int kx = SX[c], ky = SY[c];
          unsigned r[4];
          if (s4.ok) {
            const unsigned bA = axw[kx] + ayr[ky], bB = bA + half_delta, sh = shx[kx];
            unsigned w0, w1;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(bA));
            asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(bA));
            r[0] = __funnelshift_r(w0, w1, sh);
            asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(w0) : "r"(bA), "n"(RW));
            asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(w1) : "r"(bA), "n"(RW + 4));
            r[1] = __funnelshift_r(w0, w1, sh);
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(bB));
            asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(bB));
            r[2] = __funnelshift_r(w0, w1, sh);
            asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(w0) : "r"(bB), "n"(RW));
            asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(w1) : "r"(bB), "n"(RW + 4));
            r[3] = __funnelshift_r(w0, w1, sh);
          } else {
            const int x4 = s4.x4 + mvpx + (kx - 1) * step, y4 = s4.y4 + mvpy + (ky - 1) * step;
            // origin clamp per tile as computeSATD does (me_distortion.c:657), per block for SAD / SSE (:365, :1056)
            const uint8_t *rp = metric == 2 ? me_umv_ptr(ref, y4, x4)
                                            : me_umv_ptr(ref, y4 - (s4.by << 2), x4 - (s4.bx << 2)) + (size_t)s4.by * ref.pitch + s4.bx;
#pragma unroll
            for (int y = 0; y < 4; y++) {
              const uint8_t *b = rp + (size_t)((y + s4.rot) & 3) * ref.pitch;
              r[y] = (unsigned)b[0] | ((unsigned)b[1] << 8) | ((unsigned)b[2] << 16) | ((unsigned)b[3] << 24);
            }
          }
          int v = metric == 2 ? satd4x4(r, s4.c01, s4.c23) : sadsse4x4(r, s4.c01, s4.c23, metric);
          // segmented sum over the g lanes of the partition (stages no partition of this warp needs are skipped: uniform)
#pragma unroll
          for (int st = 0; st < 4; st++) {
            if ((1 << st) < gmax) {
              const int o = __shfl_xor_sync(0xffffffffu, v, 1 << st);
              asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(v) : "r"(o), "r"(gk[st]));
            }
          }
          int mc = v + ((lbx[kx] + lby[ky]) >> 16);
          if (c == 0 && check0) mc -= (lam * 16) >> 16;    // check_position0, me_fullsearch.c:361,439-442
          if (mc < mn) { mn = mc; best = c; }
        }
        if (s4.leader) {
          const int fx = mvpx + step * c_spx[best], fy = mvpy + step * c_spy[best];
          sm.mn[s4.p] = mn; sm.mv[s4.p][0] = fx; sm.mv[s4.p][1] = fy;
        }
        if (!T8) __syncwarp();       // the half-pel result is read by the lanes of the same warp only: no CTA-wide barrier
      } else {
        // ---- 8x8 tiles (block types 1..4): 16 lanes per (type, quadrant), four lane groups share the candidates
        const int mvpx = sm.mv[s8.p][0], mvpy = sm.mv[s8.p][1];
        // the same trip count for every lane (the shuffles below need the whole warp); a lane group without a candidate
        // left re-scores candidate 8 and does not store
#pragma unroll 1
        for (int k = 0; k < (9 - first + 3) / 4; k++) {
          const int cc = first + s8.grp + 4 * k, c = cc < 9 ? cc : 8;
          const int x4 = s8.x4 + mvpx + step * c_spx[c], y4 = s8.y4 + mvpy + step * c_spy[c];
          unsigned rl[8], rr[8];
          if (s8.ok) {
            const unsigned a = region_addr + (unsigned)((((y4 & 3) << 2) | (x4 & 3)) * (RW * RH) + ((y4 >> 2) - Yr) * RW + ((x4 >> 2) - Xr));
            const unsigned a4 = a & ~3u, sh = (a & 3u) * 8;
#pragma unroll
            for (int y = 0; y < 8; y++) {
              unsigned w0, w1, w2;
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(a4 + y * RW));
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"(a4 + y * RW + 4));
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w2) : "r"(a4 + y * RW + 8));
              rl[y] = __funnelshift_r(w0, w1, sh); rr[y] = __funnelshift_r(w1, w2, sh);
            }
          } else {
            const uint8_t *rp = me_umv_ptr(ref, y4, x4);   // per-tile (8x8) origin clamp
#pragma unroll
            for (int y = 0; y < 8; y++) {
              const uint8_t *b = rp + (size_t)y * ref.pitch;
              rl[y] = (unsigned)b[0] | ((unsigned)b[1] << 8) | ((unsigned)b[2] << 16) | ((unsigned)b[3] << 24);
              rr[y] = (unsigned)b[4] | ((unsigned)b[5] << 8) | ((unsigned)b[6] << 16) | ((unsigned)b[7] << 24);
            }
          }
          int v = satd8x8(rl, rr, &sm.cur8[s8.quad][0][0]);
#pragma unroll
          for (int sft = 1; sft < 4; sft <<= 1) {
            const int o = __shfl_xor_sync(0xffffffffu, v, sft);
            if (sft < s8.g) v += o;
          }
          if (s8.leader && cc < 9) sm.cost[s8.p][c] = v + sm.mvc[s8.p][c];
        }
      }
      if (stage == 0 && next_id < n_jobs) stage_b(w0n);   // the next macroblock's pixels: word 0 of its descriptor has arrived by now
      if (!T8) continue;
      __syncthreads();
      {
        // partitions of block types 1..4: their candidates were scored by four lane groups, scan them here
        if (tid < 9) {
          const int p = tid;
          int best = 0, mn = (stage == 1 && !cfg.start_qp) ? ME_INT_MAX : sm.mn[p];
          for (int c = first; c < 9; c++) {
            int mc = sm.cost[p][c];
            if (stage == 0 && c == 0 && !cfg.rdopt && jv.slice_type != ME_B_SLICE && jv.ref == 0 && p == 0 &&
                sm.mv[p][0] == 0 && sm.mv[p][1] == 0)
              mc -= (lam * 16) >> 16;                        // check_position0, me_fullsearch.c:361,439-442
            if (mc < mn) { mn = mc; best = c; }
          }
          sm.mn[p] = mn;
          sm.mv[p][0] += step * c_spx[best];
          sm.mv[p][1] += step * c_spy[best];
        }
        __syncthreads();
      }
    }
    // Without the 8x8 transform every partition lives inside one warp and the warps ran both stages without meeting.  They meet
    // here: the region and the working arrays are free, the staged copies of the next job have landed (every thread waited
    // for its own, the barrier publishes them)
    cp_async_wait_all();
    __syncthreads();
    // the job's results leave as two contiguous runs of 41 words (the result array may be pinned host memory the kernel writes
    // across PCIe: me_cuda_frame_search hands such buffers through without a copy)
    if (tid < ME_NUM_PARTITIONS) {
      reinterpret_cast<uint32_t *>(out[job_id].mv)[tid] = ((unsigned)sm.mv[tid][0] & 0xffffu) | ((unsigned)sm.mv[tid][1] << 16);
      out[job_id].cost[tid] = sm.mn[tid];
    }
    if (tid == 0) { if (next_id < n_jobs) issue_region(); sm.next_id = tk; }
    if (next_id < n_jobs) commit();
    __syncthreads();
    job_id = next_id;
    next_id = sm.next_id;
  }
}

}  // namespace

// 3-D byte tensor map over the 16 quarter-pel planes of a reference picture (x, y, plane), box = one sub-pel region
extern "C" int me_make_planes_tensor_map(void *out, const void *planes, int pitch, int hp, size_t plane_stride)
{
  const cuuint64_t dims[3] = { (cuuint64_t)pitch, (cuuint64_t)hp, 16 }, strides[2] = { (cuuint64_t)pitch, (cuuint64_t)plane_stride };
  const cuuint32_t box[3] = { RW, RH, 16 };
  return me_encode_tensor_map(out, planes, 3, dims, strides, box);
}

// Returns 1 when the kernel was launched, 0 when the configuration is not covered (the caller runs subpel.cu).
// plane_maps: host array [2][ME_MAX_REFS] of 128-byte tensor maps of the listed references' plane sets.
extern "C" int me_launch_subpel_batch_tuned(const MeJob *jobs, int n_jobs, const MeRefDev *ref_table, const void *plane_maps,
                                            int n_refs0, int n_refs1, const uint8_t *cur, int cur_pitch, const MeCfgDev *cfg,
                                            const MeResultInt *in_int, MeResult *out, int *ticket, unsigned long long *stats, cudaStream_t st)
{
  if (!plane_maps || !ticket || n_refs0 > TUNED_MAX_REFS || n_refs1 > TUNED_MAX_REFS) return 0;
  // 8x8 transform tiles exist for the Hadamard metric only: with Transform8x8Mode either both sub-pel metrics are Hadamard (8x8
  // variant) or neither (4x4 variant, the transform size does not matter); a mix goes to the general kernel
  const bool had1 = cfg->metric[1] == 2, had2 = cfg->metric[2] == 2;
  if (cfg->transform8x8 && had1 != had2) return 0;
  if (cfg->chroma.on && !(had1 && had2)) return 0;          // chroma terms of computeSAD / SSE at sub-pel positions: general kernel
  const bool use_t8 = cfg->transform8x8 && had1;
  static int general_sad = -1;                               // measurement switch: SAD / SSE sub-pel metrics through subpel.cu
  if (general_sad < 0) { const char *e = getenv("ME_CUDA_SUBPEL_SAD_GENERAL"); general_sad = e && atoi(e) ? 1 : 0; }
  if (general_sad && !(had1 && had2)) return 0;
  // per device (the library may be initialised again on another GPU of the box)
  static int sms_of[64];
  static bool attr_done[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (!attr_done[dev]) {
    if (cudaDeviceGetAttribute(&sms_of[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (cudaFuncSetAttribute(me_subpel_tuned_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<false>)) != cudaSuccess ||
        cudaFuncSetAttribute(me_subpel_tuned_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<false>)) != cudaSuccess ||
        cudaFuncSetAttribute(me_subpel_tuned_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<true>)) != cudaSuccess)
      return 0;
    attr_done[dev] = true;
  }
  const int sms = sms_of[dev];
  PlaneMaps pm;
  for (int l = 0; l < 2; l++)
    memcpy(&pm.m[l * TUNED_MAX_REFS], (const unsigned char *)plane_maps + (size_t)l * ME_MAX_REFS * 128, sizeof(CUtensorMap) * TUNED_MAX_REFS);
  static int t8_ctas = 0;
  if (!t8_ctas) { const char *e = getenv("ME_CUDA_SUBPEL_T8_CTAS"); t8_ctas = e ? atoi(e) : T8_CTAS; if (t8_ctas < 1 || t8_ctas > T8_CTAS) t8_ctas = T8_CTAS; }
  static int t4_ctas = 0;
  if (!t4_ctas) { const char *e = getenv("ME_CUDA_SUBPEL_CTAS"); t4_ctas = e ? atoi(e) : T4_CTAS; if (t4_ctas < 1 || t4_ctas > T4_CTAS) t4_ctas = T4_CTAS; }
  int grid = sms * (use_t8 ? t8_ctas : t4_ctas);
  if (grid > n_jobs) grid = n_jobs;
  if (use_t8)
    me_subpel_tuned_kernel<true, true><<<grid, NT, sizeof(Smem<true>), st>>>(reinterpret_cast<const uint32_t *>(jobs), n_jobs, ref_table, pm, cur,
                                                                       cur_pitch, *cfg, in_int, out, ticket, stats);
  else if (had1 && had2)
    me_subpel_tuned_kernel<false, true><<<grid, NT, sizeof(Smem<false>), st>>>(reinterpret_cast<const uint32_t *>(jobs), n_jobs, ref_table, pm, cur,
                                                                         cur_pitch, *cfg, in_int, out, ticket, stats);
  else
    me_subpel_tuned_kernel<false, false><<<grid, NT, sizeof(Smem<false>), st>>>(reinterpret_cast<const uint32_t *>(jobs), n_jobs, ref_table, pm, cur,
                                                                          cur_pitch, *cfg, in_int, out, ticket, stats);
  return 1;
}
