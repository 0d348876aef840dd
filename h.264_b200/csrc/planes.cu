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
// All arithmetic is integer.  One kernel computes all 16 planes of a 128x16 tile from a shared-memory copy of the tile plus halo,
// four samples per 32-bit word throughout (byte-SIMD averages, 16-bit lane pairs for the six-tap sums); a second tiny kernel
// writes the wide-border search copy of the integer plane (me_common.cuh).
//
// Roofline: HBM-bound.  Algorithmic bytes per reference picture = w*h (read) + 16*(w+40)*(h+40) (write).
#include "me_common.cuh"

#define TW 128                 // tile: 128 x 16 plane samples per CTA
#define TH 16
#define GROWS (TH + 5)         // plane rows y0-2 .. y0+TH+2: what the vertical filters and the lower neighbours of the tile read
#define GWORDS (TW / 4 + 2)    // plane columns x0-4 .. x0+TW+3 as 32-bit words (x0 is a multiple of 128: word loads stay aligned)
#define BIASP 0x0A000A00u      // 2560 in both 16-bit lanes: keeps the six-tap sums (-2550 .. 10710) non-negative; a multiple of 32

// Everything below works on words of four samples.  Two samples share a register as 16-bit lanes for the six-tap sums (the
// biased sums stay below 2^14, so plain 32-bit adds and multiplies never carry between the lanes); the vertical filter of the
// unrounded horizontal sums (plane j) needs 19 bits and runs per sample.  Rounding with the bias in place: the bias is a multiple
// of 32 and the tap weights sum to 32, so ((t + bias) + 16) >> 5 = ((t + 16) >> 5) + 80 and (sum of weighted (t + bias) + 512) >> 10
// = ((sum + 512) >> 10) + 80: clip255(v) = clamp(v + 80, 80, 335) - 80.
__device__ __forceinline__ unsigned pair_lo(unsigned w) { return __byte_perm(w, 0, 0x4140); }      // (b0 | b1 << 16)
__device__ __forceinline__ unsigned pair_hi(unsigned w) { return __byte_perm(w, 0, 0x4342); }      // (b2 | b3 << 16)
__device__ __forceinline__ unsigned pair_mid(unsigned w) { return __byte_perm(w, 0, 0x4241); }     // (b1 | b2 << 16)
// 20 * (a + b) - 5 * (c + d) + (e + f) + bias on lane pairs
__device__ __forceinline__ unsigned six_tap_pairs(unsigned a, unsigned b, unsigned c, unsigned d, unsigned e, unsigned f)
{
  return (a + b) * 20u + (e + f + BIASP) - (c + d) * 5u;
}
// two biased six-tap sums -> clip255((t + 16) >> 5) in both lanes
__device__ __forceinline__ unsigned round5_pairs(unsigned t)
{
  const unsigned v = ((t + 0x00100010u) >> 5) & 0x07ff07ffu;
  return __vmins2(__vmaxs2(v, 0x00500050u), 0x014f014fu) - 0x00500050u;
}
__device__ __forceinline__ unsigned pack_pairs(unsigned p01, unsigned p23) { return __byte_perm(p01, p23, 0x6420); }

__global__ void __launch_bounds__(256)
me_planes_kernel(const uint8_t *__restrict__ src, int src_stride, int w, int h,
                 uint8_t *__restrict__ planes, int pitch, size_t plane_stride, int src_words_ok)
{
  __shared__ uint32_t Gw[GROWS][GWORDS];        // integer samples, rows y0-2.., columns x0-4..
  __shared__ uint32_t Tw[GROWS][TW / 2];        // biased horizontal six-tap sums as lane pairs, rows y0-2.., columns x0..x0+TW-1
  __shared__ uint32_t Bw[TH + 1][TW / 4];       // b, rows y0..y0+TH
  __shared__ uint32_t Hw[TH][TW / 4 + 1];       // h, rows y0..y0+TH-1, columns x0..x0+TW+3
  __shared__ uint32_t Jw[TH][TW / 4];           // j
  const int wp = w + 2 * ME_PAD, hp = h + 2 * ME_PAD;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int tid = threadIdx.x;

  // ---- integer samples.  G(y, x) = S(clamp(y - 20, 0, h - 1), clamp(x - 20, 0, w - 1)) for every y, x: the plane-edge clamp of
  // the filters (img_luma.c:181, :293) lands in the replicated border, where it changes nothing
  for (int i = tid; i < GROWS * GWORDS; i += 256) {
    const int ly = i / GWORDS, wi = i - ly * GWORDS;
    const int sy = me_clamp(y0 - 2 + ly - ME_PAD, 0, h - 1), sx = x0 - 4 + 4 * wi - ME_PAD;
    const uint8_t *row = src + (size_t)sy * src_stride;
    uint32_t v;
    if (src_words_ok && sx >= 0 && sx + 3 <= w - 1) v = *reinterpret_cast<const uint32_t *>(row + sx);
    else
      v = (uint32_t)row[me_clamp(sx, 0, w - 1)] | ((uint32_t)row[me_clamp(sx + 1, 0, w - 1)] << 8) |
          ((uint32_t)row[me_clamp(sx + 2, 0, w - 1)] << 16) | ((uint32_t)row[me_clamp(sx + 3, 0, w - 1)] << 24);
    Gw[ly][wi] = v;
  }
  __syncthreads();
  // ---- horizontal sums t (img_luma.c:181-190), four columns per thread: samples g0..g11 = three words starting at column c - 4
  for (int i = tid; i < GROWS * (TW / 4); i += 256) {
    const int ly = i / (TW / 4), wq = i - ly * (TW / 4);
    const unsigned W0 = Gw[ly][wq], W1 = Gw[ly][wq + 1], W2 = Gw[ly][wq + 2];
    const unsigned V1 = __funnelshift_r(W0, W1, 24), V2 = __funnelshift_r(W1, W2, 24);    // (g3 g4 g5 g6), (g7 g8 g9 g10)
    const unsigned g23 = pair_hi(W0), g45 = pair_lo(W1), g67 = pair_hi(W1), g89 = pair_lo(W2);
    const unsigned g34 = pair_lo(V1), g56 = pair_mid(W1), g78 = pair_lo(V2), g9a = pair_hi(V2);
    // columns c, c + 1 (samples g4, g5) and c + 2, c + 3 (g6, g7)
    uint2 t;
    t.x = six_tap_pairs(g45, g56, g34, g67, g23, g78);
    t.y = six_tap_pairs(g67, g78, g56, g89, g45, g9a);
    *reinterpret_cast<uint2 *>(&Tw[ly][2 * wq]) = t;
  }
  __syncthreads();
  // ---- b (rows y0..y0+TH), h (columns x0..x0+TW+3) and j
  for (int i = tid; i < (TH + 1) * (TW / 4); i += 256) {
    const int r = i / (TW / 4), wq = i - r * (TW / 4);
    const uint2 t = *reinterpret_cast<const uint2 *>(&Tw[r + 2][2 * wq]);
    Bw[r][wq] = pack_pairs(round5_pairs(t.x), round5_pairs(t.y));                         // img_luma.c:206
  }
  for (int i = tid; i < TH * (TW / 4 + 1); i += 256) {
    const int r = i / (TW / 4 + 1), wq = i - r * (TW / 4 + 1);
    unsigned lo[6], hi[6];
#pragma unroll
    for (int k = 0; k < 6; k++) { const unsigned W = Gw[r + k][wq + 1]; lo[k] = pair_lo(W); hi[k] = pair_hi(W); }
    Hw[r][wq] = pack_pairs(round5_pairs(six_tap_pairs(lo[2], lo[3], lo[1], lo[4], lo[0], lo[5])),
                           round5_pairs(six_tap_pairs(hi[2], hi[3], hi[1], hi[4], hi[0], hi[5])));      // img_luma.c:293-330
  }
  for (int i = tid; i < TH * (TW / 4); i += 256) {
    const int r = i / (TW / 4), wq = i - r * (TW / 4);
    uint2 t[6];
#pragma unroll
    for (int k = 0; k < 6; k++) t[k] = *reinterpret_cast<const uint2 *>(&Tw[r + k][2 * wq]);
    // vertical six taps on the biased sums: the pair sums fit 16 bits, the weighted sum does not
    unsigned out[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const unsigned s1 = q ? t[2].y + t[3].y : t[2].x + t[3].x, s2 = q ? t[1].y + t[4].y : t[1].x + t[4].x,
                     s3 = q ? t[0].y + t[5].y : t[0].x + t[5].x;
      int v0 = (int)((s1 & 0xffffu) * 20u + (s3 & 0xffffu)) - (int)((s2 & 0xffffu) * 5u);
      int v1 = (int)((s1 >> 16) * 20u + (s3 >> 16)) - (int)((s2 >> 16) * 5u);
      v0 = me_clamp((v0 + 512) >> 10, 80, 335) - 80;                                     // img_luma.c:390-420 with the bias (see above)
      v1 = me_clamp((v1 + 512) >> 10, 80, 335) - 80;
      out[q] = (unsigned)v0 | ((unsigned)v1 << 16);
    }
    Jw[r][wq] = pack_pairs(out[0], out[1]);
  }
  __syncthreads();
  // ---- the 16 planes (img_luma.c:483-639): one word of four samples per thread and plane, a warp writes one 128-byte line
  for (int i = tid; i < TH * (TW / 4); i += 256) {
    const int r = i / (TW / 4), wq = i - r * (TW / 4);
    const int y = y0 + r, x = x0 + 4 * wq;
    if (y >= hp || x >= wp) continue;
    const unsigned g = Gw[r + 2][wq + 1], gn = Gw[r + 2][wq + 2], g_d = Gw[r + 3][wq + 1];
    const unsigned b = Bw[r][wq], b_d = Bw[r + 1][wq];
    const unsigned hv = Hw[r][wq], hn = Hw[r][wq + 1];
    const unsigned j = Jw[r][wq];
    const unsigned g_r = __funnelshift_r(g, gn, 8), h_r = __funnelshift_r(hv, hn, 8);     // right neighbours
    uint8_t *P = planes + (size_t)y * pitch + x;
#define ME_PUT(k, v) *reinterpret_cast<uint32_t *>(P + (size_t)(k) * plane_stride) = (v)
    ME_PUT(0, g);                    ME_PUT(1, __vavgu4(g, b));       ME_PUT(2, b);                    ME_PUT(3, __vavgu4(b, g_r));
    ME_PUT(4, __vavgu4(g, hv));      ME_PUT(5, __vavgu4(b, hv));      ME_PUT(6, __vavgu4(b, j));       ME_PUT(7, __vavgu4(b, h_r));
    ME_PUT(8, hv);                   ME_PUT(9, __vavgu4(hv, j));      ME_PUT(10, j);                   ME_PUT(11, __vavgu4(j, h_r));
    ME_PUT(12, __vavgu4(hv, g_d));   ME_PUT(13, __vavgu4(hv, b_d));   ME_PUT(14, __vavgu4(j, b_d));    ME_PUT(15, __vavgu4(b_d, h_r));
#undef ME_PUT
  }
}

// wide-border integer plane for the search-window tiles: one thread per 4 consecutive bytes (sw is a multiple of 4)
__global__ void __launch_bounds__(256)
me_search_plane_kernel(const uint8_t *__restrict__ src, int src_stride, int w, int h,
                       uint8_t *__restrict__ dst, int spad, int spitch, int sw, int sh, int src_words_ok)
{
  int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  int y = blockIdx.y;
  if (x >= sw || y >= sh) return;
  const uint8_t *row = src + (size_t)me_clamp(y - spad, 0, h - 1) * src_stride;
  const int sx = x - spad;
  uint32_t v = 0;
  if (src_words_ok && sx >= 0 && sx + 3 <= w - 1) v = *reinterpret_cast<const uint32_t *>(row + sx);      // inside the picture: one aligned word
  else {
#pragma unroll
    for (int k = 0; k < 4; k++) v |= (uint32_t)row[me_clamp(sx + k, 0, w - 1)] << (8 * k);
  }
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
  const int words_ok = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)src_stride) & 3) == 0;      // aligned word loads of the source rows
  me_planes_kernel<<<grid, 256, 0, st>>>(src, src_stride, w, h, planes, pitch, plane_stride, words_ok);
  int sw = w + 2 * spad, sh = h + 2 * spad;
  dim3 g2((sw / 4 + 255) / 256, sh);
  me_search_plane_kernel<<<g2, 256, 0, st>>>(src, src_stride, w, h, search, spad, spitch, sw, sh, words_ok && (spad & 3) == 0);
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

// ------------------------------------------------------------------------------------------------------------------
// The 16 planes of a picture with more than 8 bits per sample (imgpel is u16 in the reference, global.h:50): 16-bit samples
// in and out, the clips of getSubImagesLuma at img->max_imgpel_value (img_luma.c:45-639).  One thread per position of the padded
// plane computes all sixteen values from the picture with clamped coordinates -- the padded integer plane is pure edge
// replication, so every clamp the reference applies to padded-plane indices equals a clamp of picture coordinates.
// 16-bit pictures are not on the throughput path: no byte tricks, no search-plane copy (the searches read plane (0,0)).
__global__ void __launch_bounds__(256)
me_planes16_kernel(const uint16_t *__restrict__ src, int src_stride /* samples */, int w, int h, uint16_t *__restrict__ planes,
                   int pitch /* samples */, size_t plane_stride /* samples */, int max_pel)
{
  const int wp = w + 2 * ME_PAD, hp = h + 2 * ME_PAD;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= wp || y >= hp) return;
  auto L = [&](int yy, int xx) -> int {
    return (int)src[(size_t)me_clamp(yy - ME_PAD, 0, h - 1) * src_stride + me_clamp(xx - ME_PAD, 0, w - 1)];
  };
  auto clip = [&](int v) -> int { return me_clamp(v, 0, max_pel); };
  auto hsum = [&](int yy, int xx) -> int {       // unrounded horizontal six-tap sum (getHorSubImageSixTap, img_luma.c:181-280)
    return 20 * (L(yy, xx) + L(yy, xx + 1)) - 5 * (L(yy, xx - 1) + L(yy, xx + 2)) + (L(yy, xx - 2) + L(yy, xx + 3));
  };
  auto vhalf = [&](int yy, int xx) -> int {      // getVerSubImageSixTap, img_luma.c:293-380
    return clip((20 * (L(yy, xx) + L(yy + 1, xx)) - 5 * (L(yy - 1, xx) + L(yy + 2, xx)) + (L(yy - 2, xx) + L(yy + 3, xx)) + 16) >> 5);
  };
  int hs[6];
#pragma unroll
  for (int k = 0; k < 6; k++) hs[k] = hsum(y - 2 + k, x);
  const int g = L(y, x), gx1 = L(y, x + 1), gy1 = L(y + 1, x);
  const int b = clip((hs[2] + 16) >> 5), by1 = clip((hs[3] + 16) >> 5);
  const int hv = vhalf(y, x), hx1 = vhalf(y, x + 1);
  const int j = clip((20 * (hs[2] + hs[3]) - 5 * (hs[1] + hs[4]) + (hs[0] + hs[5]) + 512) >> 10);      // getVerSubImageSixTapTmp :390-470
  const size_t o = (size_t)y * pitch + x;
  auto put = [&](int fy, int fx, int v) { planes[(size_t)(fy * 4 + fx) * plane_stride + o] = (uint16_t)v; };
  put(0, 0, g); put(0, 2, b); put(2, 0, hv); put(2, 2, j);
  put(0, 1, (g + b + 1) >> 1);   put(1, 0, (g + hv + 1) >> 1);  put(1, 1, (b + hv + 1) >> 1);    // quarter-pel positions, img_luma.c:483-639
  put(1, 2, (b + j + 1) >> 1);   put(2, 1, (hv + j + 1) >> 1);
  put(0, 3, (b + gx1 + 1) >> 1); put(1, 3, (b + hx1 + 1) >> 1); put(2, 3, (j + hx1 + 1) >> 1);
  put(3, 0, (hv + gy1 + 1) >> 1); put(3, 1, (hv + by1 + 1) >> 1); put(3, 2, (j + by1 + 1) >> 1);
  put(3, 3, (by1 + hx1 + 1) >> 1);
}

extern "C" void me_launch_planes16(const uint16_t *src, int src_stride, int w, int h, uint16_t *planes, int pitch, size_t plane_stride,
                                   int max_pel, cudaStream_t st)
{
  const int wp = w + 2 * ME_PAD, hp = h + 2 * ME_PAD;
  dim3 grid((wp + 31) / 32, (hp + 7) / 8);
  me_planes16_kernel<<<grid, 256, 0, st>>>(src, src_stride, w, h, planes, pitch, plane_stride, max_pel);
}

// the same for 16-bit samples (more than 8 bits per sample): computed per sample, clip at img->max_imgpel_value
__global__ void __launch_bounds__(256)
me_weight_planes16_kernel(const uint16_t *__restrict__ src, uint16_t *__restrict__ dst, size_t n, int weight, int offset, int log_denom, int max_pel)
{
  const int round = log_denom ? 1 << (log_denom - 1) : 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = (uint16_t)me_clamp(((weight * (int)src[i] + round) >> log_denom) + offset, 0, max_pel);
}

extern "C" void me_launch_weight_planes16(const uint16_t *src, uint16_t *dst, size_t n, int weight, int offset, int log_denom, int max_pel,
                                          cudaStream_t st)
{
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks == 0) return;
  me_weight_planes16_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, dst, n, weight, offset, log_denom, max_pel);
}
