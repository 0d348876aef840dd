// me_common.cuh -- shared device/host definitions of libme_cuda (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/me_cuda.h"

#define ME_PAD ME_IMG_PAD
#define ME_MAX_REFS 32
#define ME_INT_MAX 2147483647

// One reference picture in HBM.
//  planes : the 16 quarter-pel planes of getSubImagesLuma (img_luma.c:45), u8, plane (fy,fx) at
//           planes + (fy*4+fx)*plane_stride, rows of `pitch` bytes (pitch % 128 == 0), hp = h+40 rows, wp = w+40 columns.
//  search : a second copy of the integer plane with a wider replicated border (spad pixels on every side) so that a
//           whole (2R+16)^2 search window around any macroblock whose centre drifts by up to `guard` pixels is a
//           plain in-bounds 2-D box -- the shape a TMA tile load wants.  Pixels of `search` equal plane (0,0) under
//           JM's origin clamp (refbuf.c:37) because both are pure edge replication outside the picture.
struct MeRefDev {
  const uint8_t *planes;
  const uint8_t *search;
  int w, h, wp, hp, pitch;
  size_t plane_stride;
  int spad, spitch, sw, sh;     // search plane: border, pitch, width (w+2*spad), height (h+2*spad)
};

struct MeCfgDev {
  int search_range, rdopt, full_search;
  int metric[3];
  int transform8x8;
  int mv_limit_min_y, mv_limit_max_y;
  int start_hp, start_qp;       // start_me_refinement_hp/qp (mv-search.c:396-397)
};

// partition tables (order documented in include/me_cuda.h)
__host__ __device__ __forceinline__ int me_part_type(int p)
{
  return p < 1 ? 1 : p < 3 ? 2 : p < 5 ? 3 : p < 9 ? 4 : p < 17 ? 5 : p < 25 ? 6 : 7;
}
__host__ __device__ __forceinline__ int me_part_blk(int p)
{
  // top-left 4x4 block index (4*row + col) of partition p
  if (p < 1) return 0;
  if (p < 3) return (p - 1) * 8;
  if (p < 5) return (p - 3) * 2;
  if (p < 9) { int i = p - 5; return (i >> 1) * 8 + (i & 1) * 2; }
  if (p < 17) return (p - 9) * 2;
  if (p < 25) { int i = p - 17; return (i >> 2) * 8 + (i & 3); }
  return p - 25;
}
__host__ __device__ __forceinline__ int me_first_part(int blocktype)
{
  // first partition number of a block type 1..7
  return blocktype == 1 ? 0 : blocktype == 2 ? 1 : blocktype == 3 ? 3 : blocktype == 4 ? 5 : blocktype == 5 ? 9 : blocktype == 6 ? 17 : 25;
}
__host__ __device__ __forceinline__ int me_bsx(int t) { return (t <= 2) ? 16 : (t <= 5) ? 8 : 4; }
__host__ __device__ __forceinline__ int me_bsy(int t) { return (t == 1 || t == 3) ? 16 : (t == 2 || t == 4 || t == 6) ? 8 : 4; }
// partition number of (block type, 4x4 block index of the block's top-left corner); -1 if not a partition origin
__host__ __device__ inline int me_part_index(int t, int blk)
{
  int first = me_first_part(t), n = (t == 1) ? 1 : (t <= 3) ? 2 : (t == 4) ? 4 : (t <= 6) ? 8 : 16;
  for (int i = 0; i < n; i++) if (me_part_blk(first + i) == blk) return first + i;
  return -1;
}

__host__ __device__ __forceinline__ int me_clamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// mvbits[d] (mv-search.c:333-341): 1 for 0, else 2*floor(log2|d|)+3
__device__ __forceinline__ int me_mvbits(int d)
{
  int a = d < 0 ? -d : d;
  return a == 0 ? 1 : 65 - 2 * __clz(a);
}
// MV_COST_SMP (defines.h:126-128), 32-bit wrap-around product like the C original
__device__ __forceinline__ int me_mv_cost(int f, int cx, int cy, int px, int py)
{
  return (f * (me_mvbits(cx - px) + me_mvbits(cy - py))) >> 16;
}

// Position of displacement (dx, dy) in JM's spiral order (mv-search.c:366-393).
// Ring l = max(|dx|,|dy|) starts at (2l-1)^2.  Inside the ring: first the rows y = -l / +l for x = -l+1..l-1
// (alternating top, bottom), then the columns x = -l / +l for y = -l..l (alternating left, right).
__host__ __device__ __forceinline__ int me_spiral_index(int dx, int dy)
{
  int ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
  int l = ax > ay ? ax : ay;
  if (l == 0) return 0;
  int base = (2 * l - 1) * (2 * l - 1);
  if (ay == l && ax < l) return base + 2 * (dx + l - 1) + (dy > 0 ? 1 : 0);
  return base + 2 * (2 * l - 1) + 2 * (dy + l) + (dx > 0 ? 1 : 0);
}

// inverse of me_spiral_index
__device__ __forceinline__ void me_spiral_xy(int s, int &dx, int &dy)
{
  if (s == 0) { dx = 0; dy = 0; return; }
  int l = (int)((sqrtf((float)s) + 1.0f) * 0.5f);
  while ((2 * l - 1) * (2 * l - 1) > s) l--;
  while ((2 * l + 1) * (2 * l + 1) <= s) l++;
  int k = s - (2 * l - 1) * (2 * l - 1);
  if (k < 2 * (2 * l - 1)) { dx = (k >> 1) - l + 1; dy = (k & 1) ? l : -l; }
  else { k -= 2 * (2 * l - 1); dy = (k >> 1) - l; dx = (k & 1) ? l : -l; }
}

// UMVLine4X (refbuf.c:37): byte offset into the plane set for padded quarter-pel coordinates
__device__ __forceinline__ const uint8_t *me_umv_ptr(const MeRefDev &r, int y4, int x4)
{
  int xp = me_clamp(x4 >> 2, 0, r.w + 2 * ME_PAD - 1 - 16);
  int yp = me_clamp(y4 >> 2, 0, r.h + 2 * ME_PAD - 1 - 16);
  return r.planes + (size_t)(((y4 & 3) << 2) | (x4 & 3)) * r.plane_stride + (size_t)yp * r.pitch + xp;
}

// device-side job/result layouts = the ABI structs
typedef me_mb_job MeJob;
typedef me_mb_result MeResult;
typedef me_mb_result_int MeResultInt;

// Stage one job descriptor in shared memory with whole-word loads (the caller synchronises afterwards).
// Reading the 16-bit fields of me_mb_job straight from global memory made ptxas 12.9 (sm_100a) pair the two halves of
// pmv16 into one 32-bit load and then take the high half from the wrong register -- observed as a search centre that
// depended on stale register contents; word loads + shared-memory field reads avoid that code path.
__device__ __forceinline__ void me_load_job(MeJob *dst, const MeJob *src)
{
  static_assert(sizeof(MeJob) % 4 == 0, "me_mb_job must be a whole number of words");
  const unsigned words = sizeof(MeJob) / 4;
  if (threadIdx.x < words)
    reinterpret_cast<volatile uint32_t *>(dst)[threadIdx.x] = reinterpret_cast<const uint32_t *>(src)[threadIdx.x];
}

// One block search over a stored surface (FastFullPelBlockMotionSearch, me_fullfast.c:833-903): first strict minimum of
// SAD + motion-vector rate in spiral order; with !rdopt the zero vector is scored first (:867-876).  Called by all threads
// of the CTA, returns the result to all of them.  red: 8 words of shared memory.
__device__ __forceinline__ void me_ffs_block_argmin(const int32_t *__restrict__ surf_p, int setup_range, int offx, int offy,
                                                    int pmv_x, int pmv_y, int search_range, int min_mcost, int lambda, int rdopt,
                                                    unsigned long long *red, int &out_mvx, int &out_mvy, int &out_cost)
{
  const int Pset = (2 * setup_range + 1) * (2 * setup_range + 1);
  int P = (2 * search_range + 1) * (2 * search_range + 1);
  if (P > Pset) P = Pset;
  unsigned long long best = ~0ull;
  int pos00 = -1;
  if (!rdopt && abs(offx) <= setup_range && abs(offy) <= setup_range) pos00 = me_spiral_index(-offx, -offy);
  for (int s = threadIdx.x; s < P; s += blockDim.x) {
    int dx, dy;
    me_spiral_xy(s, dx, dy);
    int mc = surf_p[s] + me_mv_cost(lambda, (offx + dx) << 2, (offy + dy) << 2, pmv_x, pmv_y);
    unsigned long long key = ((unsigned long long)(unsigned)mc << 32) | (unsigned)(s + 1);
    best = key < best ? key : best;
  }
  if (threadIdx.x == 0 && pos00 >= 0) {
    int mc = surf_p[pos00] + me_mv_cost(lambda, 0, 0, pmv_x, pmv_y);
    unsigned long long key = ((unsigned long long)(unsigned)mc << 32);
    best = key < best ? key : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other < best ? other : best;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
  __syncthreads();
  best = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); i++) best = red[i] < best ? red[i] : best;
  int cost = (int)(unsigned)(best >> 32);
  unsigned tie = (unsigned)(best & 0xffffffffu);
  int bestpos = 0;
  // min_mcost from the caller: a candidate only wins when strictly below it (me_fullfast.c:871,891)
  if (cost < min_mcost) { bestpos = tie == 0 ? pos00 : (int)tie - 1; } else cost = min_mcost;
  int dx, dy;
  me_spiral_xy(bestpos, dx, dy);
  out_mvx = offx + dx; out_mvy = offy + dy; out_cost = cost;
}
