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
  // ChromaMEEnable: the picture's reconstructed chroma planes (StorablePicture::imgUV), u8, cw x ch samples, rows of cpitch
  // bytes; nullptr when the caller has not supplied them
  const uint8_t *chroma[2];
  int cw, ch, cpitch;
  // 4:4:4: chroma[k] is instead the 16-plane set built from the U / V plane by the luma interpolator (image.c:1647-1657),
  // cplane_stride bytes per plane (0 for 4:2:0 / 4:2:2)
  size_t cplane_stride;
};

// Chroma terms of the integer-pel searches (ChromaMEEnable; me_fullfast.c:776-814, me_distortion.c:376-402, :516-543).
struct MeChromaDev {
  int on;                       // the global ChromaMEEnable as the next call sees it (me_distortion.c:56; mv-search.c:612, :779)
  int weight;                   // input->ChromaMEWeight
  int shift_x, shift_y;         // chroma_shift_x / chroma_shift_y (lencod.c:2856-2879): 3,3 for 4:2:0; 3,2 for 4:2:2
  int cur_pitch;
  const uint8_t *cur[2];        // original chroma planes (imgUV_org)
};

struct MeCfgDev {
  int search_range, rdopt, full_search;
  int metric[3];
  int transform8x8;
  int mv_limit_min_y, mv_limit_max_y;
  int start_hp, start_qp;       // start_me_refinement_hp/qp (mv-search.c:396-397)
  MeChromaDev chroma;
  // frame batch on caller memory the host has not read (pinned arrays handed through): a validation kernel in front of the searches
  // stores 1 + the index of the first job with a field outside what the kernels index with; the batch kernels then return at once
  const int *abort_flag;        // nullptr: jobs validated on the host
  // more than 8 bits per sample (imgpel is u16 in the reference, global.h:50): planes and current picture hold 16-bit samples --
  // pitch / plane_stride of MeRefDev count samples either way; served by the drop-in kernels instantiated for uint16_t
  int pel16;                    // 0: 8-bit samples
  int max_pel;                  // img->max_imgpel_value: (1 << bit depth) - 1
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

// the same for planes of samples of type Pel (uint8_t / uint16_t): pitch and plane_stride count samples
template <typename Pel>
__device__ __forceinline__ const Pel *me_umv_ptr_p(const MeRefDev &r, int y4, int x4)
{
  int xp = me_clamp(x4 >> 2, 0, r.w + 2 * ME_PAD - 1 - 16);
  int yp = me_clamp(y4 >> 2, 0, r.h + 2 * ME_PAD - 1 - 16);
  return reinterpret_cast<const Pel *>(r.planes) + (size_t)(((y4 & 3) << 2) | (x4 & 3)) * r.plane_stride + (size_t)yp * r.pitch + xp;
}

// ---- chroma sub-images (getSubImagesChroma, img_chroma.c:374-440) evaluated on the fly: sub-image (sy, sx) at padded-plane
// position (ypos, xpos) is the blend of four neighbours of the edge-replicated reconstruction with eighth-sample weights,
// rshift_rnd_sf(., 6) (generateChroma00/01/10/XX, img_chroma.c:34-372)
__device__ __forceinline__ int me_cr_pel(const MeRefDev &r, const MeChromaDev &c, int k, int ypos, int xpos, int sy, int sx)
{
  if (r.cplane_stride)      // 4:4:4: a stored quarter-pel plane set, addressed like luma (refbuf.c:53 with shift 2, mask 3)
    return r.chroma[k][(size_t)((sy << 2) | sx) * r.cplane_stride + (size_t)ypos * r.cpitch + xpos];
  const int jy = ypos - ((4 * ME_PAD) >> c.shift_y), ix = xpos - ((4 * ME_PAD) >> c.shift_x);
  const int y0 = me_clamp(jy, 0, r.ch - 1), y1 = me_clamp(jy + 1, 0, r.ch - 1), x0 = me_clamp(ix, 0, r.cw - 1), x1 = me_clamp(ix + 1, 0, r.cw - 1);
  const int kk = sy << (3 - c.shift_y), ll = sx << (3 - c.shift_x);
  const int w01 = (8 - kk) * ll, w00 = ((8 - kk) << 3) - w01, w11 = kk * ll, w10 = (kk << 3) - w11;
  const uint8_t *p = r.chroma[k];
  return (w00 * p[(size_t)y0 * r.cpitch + x0] + w01 * p[(size_t)y0 * r.cpitch + x1] +
          w10 * p[(size_t)y1 * r.cpitch + x0] + w11 * p[(size_t)y1 * r.cpitch + x1] + 32) >> 6;
}
// UMVLine8X_chroma (refbuf.c:53-59): y, x = luma quarter-pel coordinates including the luma padding
__device__ __forceinline__ void me_cr_line(const MeRefDev &r, const MeChromaDev &c, int y, int x, int &ypos, int &xpos, int &sy, int &sx)
{
  xpos = me_clamp(x >> c.shift_x, 0, r.cw - 1 + 2 * ((4 * ME_PAD) >> c.shift_x) - (64 >> c.shift_x));    // size_x_cr_pad, mbuffer.c:425
  ypos = me_clamp(y >> c.shift_y, 0, r.ch - 1 + 2 * ((4 * ME_PAD) >> c.shift_y) - (64 >> c.shift_y));
  sy = y & ((1 << c.shift_y) - 1);
  sx = x & ((1 << c.shift_x) - 1);
}
// chroma part of computeSAD / SSE (me_distortion.c:376-402, :1069-1091) or of computeBiPredSAD1 (:516-543, r2 != nullptr) for the
// luma block bsx x bsy at (pic_pix_x, pic_pix_y): distortion of the co-located block of component k, NOT yet weighted
// bw_*: weighted bi-prediction of the chroma samples (computeBiPredSAD2, me_distortion.c:608-643: weight1_cr / weight2_cr /
// offsetBi_cr with the LUMA denominator and rounding of the function, :565-566 -- kept as it is)
__device__ __forceinline__ int me_cr_block(const MeChromaDev &c, int sse, int k, int pic_pix_x, int pic_pix_y, int bsx, int bsy,
                                           const MeRefDev &r1, int x41, int y41, const MeRefDev *r2, int x42, int y42,
                                           int bw_apply = 0, int bw_w1 = 0, int bw_w2 = 0, int bw_obi = 0, int bw_denom = 0)
{
  const int scx = c.shift_x - 2, scy = c.shift_y - 2;
  const int bx = bsx >> scx, by = bsy >> scy, cx0 = pic_pix_x >> scx, cy0 = pic_pix_y >> scy;
  int yp1, xp1, sy1, sx1, yp2 = 0, xp2 = 0, sy2 = 0, sx2 = 0, m = 0;
  me_cr_line(r1, c, y41, x41, yp1, xp1, sy1, sx1);
  if (r2) me_cr_line(*r2, c, y42, x42, yp2, xp2, sy2, sx2);
  for (int y = 0; y < by; y++)
    for (int x = 0; x < bx; x++) {
      int p = me_cr_pel(r1, c, k, yp1 + y, xp1 + x, sy1, sx1);
      if (r2) {
        const int q = me_cr_pel(*r2, c, k, yp2 + y, xp2 + x, sy2, sx2);
        p = bw_apply ? me_clamp(((bw_w1 * p + bw_w2 * q + (bw_denom ? 1 << bw_denom : 0)) >> (bw_denom + 1)) + bw_obi, 0, 255) : (p + q + 1) >> 1;
      }
      const int d = (int)c.cur[k][(size_t)(cy0 + y) * c.cur_pitch + cx0 + x] - p;
      m += sse ? d * d : abs(d);
    }
  return m;
}
// chroma term of the fast full search (me_fullfast.c:776-814): the co-located chroma block of every 4x4 luma block, both
// components, added to the 16 leaves as it is (no ChromaMEWeight).  abs_x / abs_y: the candidate's position in the picture.
__device__ __forceinline__ void me_leaf_chroma(const MeRefDev &r, const MeChromaDev &c, int sse, int opix_x, int opix_y, int abs_x, int abs_y,
                                               int leaf[16])
{
  const int mbx = 64 >> c.shift_x, mby = 64 >> c.shift_y;
  const int ocx = (mbx * opix_x) >> 4, ocy = (mby * opix_y) >> 4;      // img->opix_c_x/y, macroblock.c:206-207
  int yp, xp, sy, sx;
  me_cr_line(r, c, (abs_y + ME_PAD) << 2, (abs_x + ME_PAD) << 2, yp, xp, sy, sx);
  for (int k = 0; k < 2; k++)
    for (int cy = 0; cy < mby; cy++)
      for (int cx = 0; cx < mbx; cx++) {
        const int d = me_cr_pel(r, c, k, yp + cy, xp + cx, sy, sx) - (int)c.cur[k][(size_t)(ocy + cy) * c.cur_pitch + ocx + cx];
        const int b = (cy / (mby >> 2)) * 4 + cx / (mbx >> 2);
        // leaf[] is indexed with compile-time constants everywhere else: keep it in registers
#pragma unroll
        for (int q = 0; q < 16; q++) if (q == b) leaf[q] += sse ? d * d : abs(d);
      }
}

// search range of one block search (PartitionMotionSearch, mv-search.c:1411-1416)
__host__ __device__ __forceinline__ int me_block_range(const MeCfgDev &c, int ref, int blocktype)
{
  int r1 = (ref < 1 ? ref : 1) + 1;
  if (c.full_search == 2) return c.search_range;
  if (c.full_search == 1) return c.search_range / r1;
  return c.search_range / (r1 * (blocktype < 2 ? blocktype : 2));
}

// ---- one launch per PartitionMotionSearch (me_cuda_pms_begin, include/me_cuda.h): kernel parameters and result records
struct MePmsRef {               // one (list, ref) of the macroblock: planes the uni-predictive searches read, stored surfaces, centre
  MeRefDev dev;
  const int32_t *surf;          // [41][P] in spiral order (me_ffs_surfaces_kernel)
  int offx, offy, setup_range;  // search centre relative to the macroblock origin, range the surfaces were built for
  int list, ref;
};
struct MePmsParams {
  me_mb_motion st;
  MePmsRef refs[ME_PMS_MAX_REFS];
  int n_refs, n_items;
  int item_type[ME_PMS_MAX_ITEMS], item_blk[ME_PMS_MAX_ITEMS];
  int item_slot[ME_PMS_MAX_ITEMS];      // where the item's records go: result[(slot * ME_PMS_MAX_REFS + list-ref) * 4 + k]
  int lambda[3];
  int opix_x, opix_y, slice_type;
};
#define ME_PMS_SLOTS 8          // items whose results can wait at the same time (a later launch must not drop e.g. the 8x16 guess)
struct MePmsRecord {            // one block search of a group launch, as the drop-in calls will ask for it
  int pmv_x, pmv_y;             // the predictor the device derived (SetMotionVectorPredictor)
  int int_mvx, int_mvy, int_cost;       // FastFullPelBlockMotionSearch result (integer pel)
  int mvx, mvy, cost;           // SubPelBlockMotionSearch result (quarter pel)
};
// sub-blocks of PartitionMotionSearch(blocktype, block8x8) in call order (mv-search.c:1387-1401, :1424-1430): 4x4-unit origin
// (h, v) of the k-th block search; returns their number
__host__ __device__ inline int me_pms_blocks(int blocktype, int block8x8, int hs[4], int vs[4])
{
  const int parttype = blocktype < 4 ? blocktype : 4;
  const int psx[8] = { 4, 4, 4, 2, 2, 2, 1, 1 }, psy[8] = { 4, 4, 2, 4, 2, 1, 2, 1 };     // input->part_size (configfile.c:805-841)
  const int bx0[5][4] = { { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 2, 0, 0 }, { 0, 2, 0, 2 } };
  const int by0[5][4] = { { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 2, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 2, 2 } };
  const int by = by0[parttype][block8x8], bx = bx0[parttype][block8x8];
  int n = 0;
  for (int v = by; v < by + psy[parttype]; v += psy[blocktype])
    for (int h = bx; h < bx + psx[parttype]; h += psx[blocktype]) { hs[n] = h; vs[n] = v; n++; }
  return n;
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
