// planes.cu -- quarter-pel reference plane generation on the device.
//
// Replaces getSubImagesLuma and its helpers (reference lencod/src/img_luma.c:45-639, called from
// UnifiedOneForthPix, lencod/src/image.c:1601-1661): from one reconstructed luma picture S (w x h) build the 16
// planes P[fy][fx] of size (h+40) x (w+40):
//   G = P[0][0](y,x) = S(clamp(y-20,0,h-1), clamp(x-20,0,w-1))                                   img_luma.c:120
//   t(y,x)  = 20*(G(x)+G(x+1)) - 5*(G(x-1)+G(x+2)) + (G(x-2)+G(x+3))   columns clamped to [0,wp-1]  img_luma.c:181
//   b = P[0][2] = clip255((t+16)>>5)
//   h = P[2][0] = clip255((same 6 taps vertically on G, rows clamped to [0,hp-1], +16)>>5)        img_luma.c:293
//   j = P[2][2] = clip255((6 taps vertically on t, +512)>>10)                                      img_luma.c:390
//   twelve quarter positions = (p+q+1)>>1 of two of {G,b,h,j} at the same / right / lower sample    img_luma.c:483-639
// All arithmetic is 32-bit integer.  One kernel computes all 16 planes of a 64x16 tile from a shared-memory copy of
// the tile plus halo; a second tiny kernel writes the wide-border search copy of the integer plane (me_common.cuh).
//
// Roofline: HBM-bound.  Algorithmic bytes per reference picture = w*h (read) + 16*(w+40)*(h+40) (write).
#include "me_common.cuh"

#define TW 64
#define TH 16
#define GW (TW + 7)   // columns x0-2 .. x0+TW+4
#define GH (TH + 7)   // rows    y0-2 .. y0+TH+4
#define TMW (TW + 1)  // t / b / j for columns x0 .. x0+TW

__device__ __forceinline__ int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

__global__ void __launch_bounds__(256)
me_planes_kernel(const uint8_t *__restrict__ src, int src_stride, int w, int h,
                 uint8_t *__restrict__ planes, int pitch, size_t plane_stride)
{
  __shared__ uint8_t Gs[GH][GW + 1];
  __shared__ int     Ts[GH][TMW];        // unrounded horizontal sums t for rows y0-2..y0+TH+4, columns x0..x0+TW
  __shared__ uint8_t Bs[TH + 1][TMW];    // b for rows y0..y0+TH
  __shared__ uint8_t Hs[TH + 1][TMW];    // h
  __shared__ uint8_t Js[TH + 1][TMW];    // j
  const int wp = w + 2 * ME_PAD, hp = h + 2 * ME_PAD;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int tid = threadIdx.x;

  // integer samples with both clamps (padded-plane clamp of the filters, then picture clamp of the border)
  for (int i = tid; i < GH * GW; i += blockDim.x) {
    int ly = i / GW, lx = i - ly * GW;
    int y = me_clamp(y0 - 2 + ly, 0, hp - 1), x = me_clamp(x0 - 2 + lx, 0, wp - 1);
    Gs[ly][lx] = src[(size_t)me_clamp(y - ME_PAD, 0, h - 1) * src_stride + me_clamp(x - ME_PAD, 0, w - 1)];
  }
  __syncthreads();
  // horizontal six-tap sums.  Local column lx of Ts is plane column x0+lx; its taps are Gs columns lx .. lx+5.
  // The plane-edge clamp of the taps is already in Gs for every column except when x0+lx itself is past wp-1
  // (only the +1 neighbour column of the last tile), where the value is never used unclamped (see idx1 below).
  for (int i = tid; i < GH * TMW; i += blockDim.x) {
    int ly = i / TMW, lx = i - ly * TMW;
    const uint8_t *g = &Gs[ly][lx];
    Ts[ly][lx] = 20 * (g[2] + g[3]) - 5 * (g[1] + g[4]) + (g[0] + g[5]);
  }
  __syncthreads();
  for (int i = tid; i < (TH + 1) * TMW; i += blockDim.x) {
    int ly = i / TMW, lx = i - ly * TMW;        // plane row y0+ly, column x0+lx
    Bs[ly][lx] = (uint8_t)clip255((Ts[ly + 2][lx] + 16) >> 5);
    int v = 20 * (Gs[ly + 2][lx + 2] + Gs[ly + 3][lx + 2]) - 5 * (Gs[ly + 1][lx + 2] + Gs[ly + 4][lx + 2])
            + (Gs[ly][lx + 2] + Gs[ly + 5][lx + 2]);
    Hs[ly][lx] = (uint8_t)clip255((v + 16) >> 5);
    int u = 20 * (Ts[ly + 2][lx] + Ts[ly + 3][lx]) - 5 * (Ts[ly + 1][lx] + Ts[ly + 4][lx]) + (Ts[ly][lx] + Ts[ly + 5][lx]);
    Js[ly][lx] = (uint8_t)clip255((u + 512) >> 10);
  }
  __syncthreads();
  // write the 16 planes: one thread per 4 consecutive pixels, one 32-bit store per plane (wp is a multiple of 4 and the
  // plane pitch a multiple of 128, so a group never straddles the plane edge and every store is aligned)
  for (int i = tid; i < TH * (TW / 4); i += blockDim.x) {
    int ly = i / (TW / 4), lx0 = (i - ly * (TW / 4)) * 4;
    int y = y0 + ly, x = x0 + lx0;
    if (y >= hp || x >= wp) continue;
    int ly1 = (y + 1 > hp - 1) ? ly : ly + 1;   // lower neighbour, clamped at the plane edge
    uint32_t o[16];
#pragma unroll
    for (int k = 0; k < 16; k++) o[k] = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      int lx = lx0 + k;
      int lx1 = (x + k + 1 > wp - 1) ? lx : lx + 1;   // right neighbour, clamped at the plane edge
      int g = Gs[ly + 2][lx + 2], b = Bs[ly][lx], hv = Hs[ly][lx], j = Js[ly][lx];
      int g_r = Gs[ly + 2][lx1 + 2], h_r = Hs[ly][lx1];
      int g_d = Gs[ly1 + 2][lx + 2], b_d = Bs[ly1][lx];
      const int sh = 8 * k;
      o[0]  |= (uint32_t)g << sh;
      o[1]  |= (uint32_t)((g + b + 1) >> 1) << sh;
      o[2]  |= (uint32_t)b << sh;
      o[3]  |= (uint32_t)((b + g_r + 1) >> 1) << sh;
      o[4]  |= (uint32_t)((g + hv + 1) >> 1) << sh;
      o[5]  |= (uint32_t)((b + hv + 1) >> 1) << sh;
      o[6]  |= (uint32_t)((b + j + 1) >> 1) << sh;
      o[7]  |= (uint32_t)((b + h_r + 1) >> 1) << sh;
      o[8]  |= (uint32_t)hv << sh;
      o[9]  |= (uint32_t)((hv + j + 1) >> 1) << sh;
      o[10] |= (uint32_t)j << sh;
      o[11] |= (uint32_t)((j + h_r + 1) >> 1) << sh;
      o[12] |= (uint32_t)((hv + g_d + 1) >> 1) << sh;
      o[13] |= (uint32_t)((hv + b_d + 1) >> 1) << sh;
      o[14] |= (uint32_t)((j + b_d + 1) >> 1) << sh;
      o[15] |= (uint32_t)((b_d + h_r + 1) >> 1) << sh;
    }
    uint8_t *P = planes + (size_t)y * pitch + x;
#pragma unroll
    for (int k = 0; k < 16; k++) *reinterpret_cast<uint32_t *>(P + (size_t)k * plane_stride) = o[k];
  }
}

// wide-border integer plane for the search-window tiles: one thread per 4 consecutive bytes (sw is a multiple of 4)
__global__ void __launch_bounds__(256)
me_search_plane_kernel(const uint8_t *__restrict__ src, int src_stride, int w, int h,
                       uint8_t *__restrict__ dst, int spad, int spitch, int sw, int sh)
{
  int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  int y = blockIdx.y;
  if (x >= sw || y >= sh) return;
  const uint8_t *row = src + (size_t)me_clamp(y - spad, 0, h - 1) * src_stride;
  uint32_t v = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) v |= (uint32_t)row[me_clamp(x + k - spad, 0, w - 1)] << (8 * k);
  *reinterpret_cast<uint32_t *>(dst + (size_t)y * spitch + x) = v;
}

// widen one plane row block back to imgpel for the host copy of imgY_sub
__global__ void __launch_bounds__(256)
me_widen_planes_kernel(const uint8_t *__restrict__ planes, int pitch, size_t plane_stride, int wp, int hp,
                       uint16_t *__restrict__ dst)
{
  // dst: [16][hp][wp] contiguous
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = (size_t)16 * hp * wp;
  if (i >= n) return;
  int x = (int)(i % wp);
  size_t t = i / wp;
  int y = (int)(t % hp), pl = (int)(t / hp);
  dst[i] = planes[(size_t)pl * plane_stride + (size_t)y * pitch + x];
}

extern "C" void me_launch_planes(const uint8_t *src, int src_stride, int w, int h, uint8_t *planes, int pitch,
                                 size_t plane_stride, uint8_t *search, int spad, int spitch, cudaStream_t st)
{
  int wp = w + 2 * ME_PAD, hp = h + 2 * ME_PAD;
  dim3 grid((wp + TW - 1) / TW, (hp + TH - 1) / TH);
  me_planes_kernel<<<grid, 256, 0, st>>>(src, src_stride, w, h, planes, pitch, plane_stride);
  int sw = w + 2 * spad, sh = h + 2 * spad;
  dim3 g2((sw / 4 + 255) / 256, sh);
  me_search_plane_kernel<<<g2, 256, 0, st>>>(src, src_stride, w, h, search, spad, spitch, sw, sh);
}

extern "C" void me_launch_widen_planes(const uint8_t *planes, int pitch, size_t plane_stride, int wp, int hp,
                                       uint16_t *dst, cudaStream_t st)
{
  size_t n = (size_t)16 * hp * wp;
  me_widen_planes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(planes, pitch, plane_stride, wp, hp, dst);
}

// Weighted-prediction ME (UseWeightedReferenceME): every reference sample a uni-predictive distortion reads is replaced by
//   weighted_pel = iClip1(255, ((weight * p + round) >> log_denom) + offset),  round = 1 << (log_denom - 1)
// (me_fullfast.c:653, computeSADWP me_distortion.c:413, computeSATDWP :734, weighted_prediction.c:50).  That is a function of
// the sample alone, and JM's origin clamp moves coordinates, not values -- so the searches run unchanged on a second copy
// of the picture's planes with the 256-entry table applied once per (picture, weights).  HBM-bound: one read and one
// write of the plane set (2 x 40 MB at 1080p); 16 bytes per thread, table in shared memory.
__global__ void __launch_bounds__(256)
me_weight_planes_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n16, int weight, int offset, int log_denom)
{
  __shared__ uint8_t lut[256];
  {
    const int p = threadIdx.x, round = log_denom ? 1 << (log_denom - 1) : 0;
    lut[p] = (uint8_t)me_clamp(((weight * p + round) >> log_denom) + offset, 0, 255);
  }
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = src[i];
    uint4 o;
    const uint32_t in[4] = { v.x, v.y, v.z, v.w };
    uint32_t out[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
      out[k] = (uint32_t)lut[in[k] & 0xffu] | ((uint32_t)lut[(in[k] >> 8) & 0xffu] << 8) |
               ((uint32_t)lut[(in[k] >> 16) & 0xffu] << 16) | ((uint32_t)lut[in[k] >> 24] << 24);
    o.x = out[0]; o.y = out[1]; o.z = out[2]; o.w = out[3];
    dst[i] = o;
  }
}

// bytes: multiple of 16 (plane sets and search planes have 128-byte pitches)
extern "C" void me_launch_weight_planes(const uint8_t *src, uint8_t *dst, size_t bytes, int weight, int offset, int log_denom,
                                        cudaStream_t st)
{
  const size_t n16 = bytes / 16;
  size_t blocks = (n16 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks == 0) return;
  me_weight_planes_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint4 *>(dst), n16,
                                                            weight, offset, log_denom);
}
