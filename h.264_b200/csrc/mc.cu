// mc.cu -- motion-compensated luma prediction and the mode-decision costs built on it.
//
// The host-side consumers of the quarter-pel planes other than the motion searches:
//   OneComponentLumaPrediction / LumaPrediction   reference lencod/src/macroblock.c:807-827, :832-930
//   distortion4x4 / distortion8x8                 lencod/src/me_distortion.c:76-103, :110-137 (HadamardSAD4x4 :182, HadamardSAD8x8 :272)
//   the residual part of BIDPartitionCost, GetSkipCostMB, GetDirectCost8x8   lencod/src/mv-search.c:1052-1128, :1136-1181, :1273-1325
// A block's samples come from the plane Pic[y & 3][x & 3] of the listed picture at UMVLine4X's clamped block origin
// (refbuf.c:37-43), rows following at the plane pitch -- never from the weighted copies the searches may use: prediction
// weights are applied to the fetched samples exactly as macroblock.c:876-909 does.
#include "me_common.cuh"

namespace {

// the same 24-byte layout as me_pred_block (include/me_cuda.h)
struct PredBlock {
  int16_t pix_x, pix_y;
  int8_t p_dir, ref[2], apply_weights;
  int16_t mv[2][2];
  int16_t weight[2], offset, log_denom;
};
static_assert(sizeof(PredBlock) == sizeof(me_pred_block), "me_pred_block layout");

// origin of the block in one list's plane set (OneComponentLumaPrediction, macroblock.c:816)
template <typename Pel = uint8_t>
__device__ __forceinline__ const Pel *block_origin(const MeRefDev *table, const PredBlock &b, int list)
{
  const MeRefDev &r = table[list * ME_MAX_REFS + b.ref[list]];
  return me_umv_ptr_p<Pel>(r, ((b.pix_y + ME_PAD) << 2) + b.mv[list][1], ((b.pix_x + ME_PAD) << 2) + b.mv[list][0]);
}

// sample of the prediction from the sample(s) of the listed picture(s) (LumaPrediction, macroblock.c:876-930)
// max_pel: img->max_imgpel_value (255 for 8-bit pictures)
__device__ __forceinline__ int pred_pel(const PredBlock &b, int l0, int l1, int max_pel = 255)
{
  if (b.apply_weights) {
    const int round = (1 << b.log_denom) >> 1;                        // wp_luma_round (weighted_prediction.c:50)
    if (b.p_dir == 2) return me_clamp(((b.weight[0] * l0 + b.weight[1] * l1 + 2 * round) >> (b.log_denom + 1)) + b.offset, 0, max_pel);
    return me_clamp(((b.weight[0] * (b.p_dir ? l1 : l0) + round) >> b.log_denom) + b.offset, 0, max_pel);
  }
  return b.p_dir == 2 ? (l0 + l1 + 1) >> 1 : b.p_dir ? l1 : l0;
}

// HadamardSAD4x4 (me_distortion.c:182-264)
__device__ int hadamard4x4(const int *d)
{
  int t[16], s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int a = d[4 * i], b = d[4 * i + 1], c = d[4 * i + 2], e = d[4 * i + 3];
    const int p = a + e, q = b + c, u = a - e, v = b - c;
    t[4 * i] = p + q; t[4 * i + 1] = p - q; t[4 * i + 2] = u + v; t[4 * i + 3] = u - v;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int a = t[i], b = t[4 + i], c = t[8 + i], e = t[12 + i];
    const int p = a + e, q = b + c, u = a - e, v = b - c;
    s += abs(p + q) + abs(p - q) + abs(u + v) + abs(u - v);
  }
  return (s + 1) >> 1;
}

// one thread per sample: pred[n][bsy][bsx]
__global__ void __launch_bounds__(256)
me_luma_prediction_kernel(const PredBlock *__restrict__ blocks, int n, int bsx, int bsy, const MeRefDev *__restrict__ table,
                          uint8_t *__restrict__ pred)
{
  const int per = bsx * bsy;
  const size_t total = (size_t)n * per;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / per), o = (int)(i - (size_t)k * per), y = o / bsx, x = o - y * bsx;
    const PredBlock b = blocks[k];
    int l0 = 0, l1 = 0;
    if (b.p_dir == 0 || b.p_dir == 2) l0 = block_origin(table, b, 0)[(size_t)y * table[b.ref[0]].pitch + x];
    if (b.p_dir == 1 || b.p_dir == 2) l1 = block_origin(table, b, 1)[(size_t)y * table[ME_MAX_REFS + b.ref[1]].pitch + x];
    pred[i] = (uint8_t)(b.p_dir < 0 ? 0 : pred_pel(b, l0, l1));
  }
}

// one warp per 8x8 region: lane -> row (lane >> 2), columns 2 * (lane & 3) and + 1.  The differences go through shared memory
// to the lanes that own a transform: lanes 0..3 one 4x4 block each, lanes 8..15 one row / column of the 8x8 transform each.
// Pel = uint16_t: pictures of more than 8 bits per sample (16-bit planes and current picture, weighted samples clip at max_pel)
template <typename Pel>
__global__ void __launch_bounds__(128)
me_pred_cost8x8_kernel(const PredBlock *__restrict__ blocks, int n_regions, int metric, const MeRefDev *__restrict__ table,
                       const Pel *__restrict__ cur, int cur_pitch, int max_pel, int32_t *__restrict__ cost4, int32_t *__restrict__ cost8)
{
  __shared__ int diff[4][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int reg = blockIdx.x * 4 + warp; reg < n_regions; reg += gridDim.x * 4) {
    const int row = lane >> 2, col = 2 * (lane & 3);
    const int k = (row >> 2) * 2 + (col >> 2);                        // 4x4 block of this lane's two samples, raster order
    const PredBlock b = blocks[(size_t)4 * reg + k];
    const unsigned allowed = __ballot_sync(0xffffffffu, b.p_dir >= 0);
    int *d = diff[warp];
    if (allowed == 0xffffffffu) {
      const int j = row & 3, i = col & 3;
      int l0[2] = { 0, 0 }, l1[2] = { 0, 0 };
      if (b.p_dir == 0 || b.p_dir == 2) {
        const Pel *p = block_origin<Pel>(table, b, 0) + (size_t)j * table[b.ref[0]].pitch + i;
        l0[0] = p[0]; l0[1] = p[1];
      }
      if (b.p_dir == 1 || b.p_dir == 2) {
        const Pel *p = block_origin<Pel>(table, b, 1) + (size_t)j * table[ME_MAX_REFS + b.ref[1]].pitch + i;
        l1[0] = p[0]; l1[1] = p[1];
      }
      // the block's position in the picture: its own pix_x / pix_y (the four blocks of a region are adjacent in the callers, but
      // nothing here relies on it)
      const Pel *c = cur + (size_t)(b.pix_y + j) * cur_pitch + b.pix_x + i;
      d[8 * row + col] = (int)c[0] - pred_pel(b, l0[0], l1[0], max_pel);
      d[8 * row + col + 1] = (int)c[1] - pred_pel(b, l0[1], l1[1], max_pel);
    }
    __syncwarp();
    if (allowed != 0xffffffffu) {                                    // GetDirectCost8x8: direct mode not allowed for a block of the region
      if (lane == 0) { cost4[reg] = ME_INT_MAX; cost8[reg] = ME_INT_MAX; }
      continue;
    }
    int c4 = 0, c8 = 0;
    if (metric == 2) {
      if (lane < 4) {                                                // distortion4x4 = HadamardSAD4x4 of block `lane`
        int v[16];
#pragma unroll
        for (int y = 0; y < 4; y++)
#pragma unroll
          for (int x = 0; x < 4; x++) v[4 * y + x] = d[8 * (4 * (lane >> 1) + y) + 4 * (lane & 1) + x];
        c4 = hadamard4x4(v);
      }
      c4 += __shfl_xor_sync(0xffffffffu, c4, 1);
      c4 += __shfl_xor_sync(0xffffffffu, c4, 2);
      // HadamardSAD8x8 (me_distortion.c:272-343): rows by lanes 8..15, then columns, sum of magnitudes, (s + 2) >> 2
      __syncwarp();
      if (lane >= 8 && lane < 16) {
        int *r = d + 8 * (lane - 8);
#pragma unroll
        for (int st = 4; st >= 1; st >>= 1)
#pragma unroll
          for (int x = 0; x < 8; x++)
            if (!(x & st)) { const int a = r[x], bb = r[x + st]; r[x] = a + bb; r[x + st] = a - bb; }
      }
      __syncwarp();
      int s = 0;
      if (lane >= 8 && lane < 16) {
        int v[8];
#pragma unroll
        for (int y = 0; y < 8; y++) v[y] = d[8 * y + (lane - 8)];
#pragma unroll
        for (int st = 4; st >= 1; st >>= 1)
#pragma unroll
          for (int y = 0; y < 8; y++)
            if (!(y & st)) { const int a = v[y], bb = v[y + st]; v[y] = a + bb; v[y + st] = a - bb; }
#pragma unroll
        for (int y = 0; y < 8; y++) s += abs(v[y]);
      }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      c8 = (s + 2) >> 2;
    } else {
      // SAD / squared error: the same sum whichever way the samples are grouped (byte_abs / img->quad, me_distortion.c:82-95)
      const int a = d[8 * row + col], bb = d[8 * row + col + 1];
      int s = metric == 0 ? abs(a) + abs(bb) : a * a + bb * bb;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      c4 = c8 = s;
    }
    if (lane == 0) { cost4[reg] = c4; cost8[reg] = c8; }
    __syncwarp();
  }

}

// a rectangle of one plane widened to imgpel (the binding pages planes into StorablePicture::imgY_sub with it)
__global__ void __launch_bounds__(256)
me_widen_region_kernel(const uint8_t *__restrict__ plane, int pitch, int x0, int y0, int w, int h, uint16_t *__restrict__ dst)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  const int y = i / w, x = i - y * w;
  dst[i] = plane[(size_t)(y0 + y) * pitch + x0 + x];
}

}  // namespace

extern "C" void me_launch_luma_prediction(const void *blocks, int n, int bsx, int bsy, const MeRefDev *table, uint8_t *pred, cudaStream_t st)
{
  const size_t total = (size_t)n * bsx * bsy;
  size_t grid = (total + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  me_luma_prediction_kernel<<<(unsigned)grid, 256, 0, st>>>((const PredBlock *)blocks, n, bsx, bsy, table, pred);
}

extern "C" void me_launch_pred_cost8x8(const void *blocks, int n_regions, int metric, const MeRefDev *table, const uint8_t *cur,
                                       int cur_pitch, int pel16, int max_pel, int32_t *cost4, int32_t *cost8, cudaStream_t st)
{
  int grid = (n_regions + 3) / 4;
  if (grid > 148 * 8) grid = 148 * 8;
  if (pel16)
    me_pred_cost8x8_kernel<uint16_t><<<grid, 128, 0, st>>>((const PredBlock *)blocks, n_regions, metric, table, reinterpret_cast<const uint16_t *>(cur),
                                                            cur_pitch, max_pel, cost4, cost8);
  else
    me_pred_cost8x8_kernel<uint8_t><<<grid, 128, 0, st>>>((const PredBlock *)blocks, n_regions, metric, table, cur, cur_pitch, 255, cost4, cost8);
}

extern "C" void me_launch_widen_region(const uint8_t *plane, int pitch, int x0, int y0, int w, int h, uint16_t *dst, cudaStream_t st)
{
  me_widen_region_kernel<<<(w * h + 255) / 256, 256, 0, st>>>(plane, pitch, x0, y0, w, h, dst);
}
