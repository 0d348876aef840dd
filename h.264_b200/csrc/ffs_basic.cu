// ffs_basic.cu -- integer-pel fast full search, straightforward kernels.
//
// Replaces SetupFastFullPelSearch (reference lencod/src/me_fullfast.c:491-823), SetupLargerBlocks (:210-273) and
// FastFullPelBlockMotionSearch (:833-903).
//
//  * me_ffs_batch_basic_kernel : one CTA per (macroblock, list, ref) job; every thread owns whole search positions:
//    16 4x4 SADs with __vsadu4 (VABSDIFF4.U8.ACC) from a shared-memory copy of the (2R+16)^2 search window, the 25
//    larger-block sums, then for all 41 partitions SAD + lambda*R(mv - pmv) folded into a (cost, spiral index) key
//    whose minimum is exactly JM's "first strict minimum in spiral order" (me_fullfast.c:879-897).  General in the
//    metric (SAD or, as the reference does for any non-zero MEDistortionFPel, squared error :512) and in the search
//    range; it is the functional reference on the device and the fallback of the tuned kernel in ffs_tuned.cu.
//  * me_ffs_surfaces_kernel / me_ffs_block_kernel : the two halves as separate launches for the drop-in call
//    sequence of the encoder (setup once per (macroblock, list, ref), then one argmin per block search).
#include "me_common.cuh"

// search centre (me_fullfast.c:550-566)
__device__ __forceinline__ void me_search_centre(const MeCfgDev &c, int range, int opix_x, int opix_y,
                                                 int pmv16_x, int pmv16_y, int &cx, int &cy)
{
  int x = pmv16_x / 4, y = pmv16_y / 4;             // C division: truncates toward zero
  if (!c.rdopt) { x = me_clamp(x, -range, range); y = me_clamp(y, -range, range); }
  x = me_clamp(x, -2047 + range, 2047 - range);
  y = me_clamp(y, c.mv_limit_min_y + range, c.mv_limit_max_y - range);
  cx = x + opix_x; cy = y + opix_y;
}
__device__ __forceinline__ int me_setup_range(const MeCfgDev &c, int ref)
{
  return (c.full_search == 2 || ref == 0) ? c.search_range : c.search_range / 2;   // me_fullfast.c:127-140
}

// Fill the shared-memory window: tile(y,x) = picture sample at (cy-R+y, cx-R+x) with edge replication.  Outside the
// picture JM clamps the 16x16 block ORIGIN into the 20-pixel border (refbuf.c:37); as the border is pure replication
// this equals per-sample replication.  The wide-border search plane makes the clamp a no-op for almost all windows.
__device__ __forceinline__ void me_load_window(uint8_t *tile, int tp, int n, const MeRefDev &r, int cx, int cy, int R)
{
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
    int y = i / n, x = i - y * n;
    int sy = me_clamp(cy - R + y + r.spad, 0, r.sh - 1), sx = me_clamp(cx - R + x + r.spad, 0, r.sw - 1);
    tile[y * tp + x] = r.search[(size_t)sy * r.spitch + sx];
  }
}

// 16 leaf distortions of one position from the window tile.  SSE selects squared error.
template <bool SSE>
__device__ __forceinline__ void me_leaf16(const uint8_t *tile, int tp, const uint32_t *curw, int dx, int dy, int leaf[16])
{
#pragma unroll
  for (int b = 0; b < 16; b++) leaf[b] = 0;
  const int sh = (dx & 3) * 8;
#pragma unroll
  for (int row = 0; row < 16; row++) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(tile + (dy + row) * tp + (dx & ~3));
    uint32_t a0 = w[0], a1 = w[1], a2 = w[2], a3 = w[3], a4 = w[4];
    uint32_t r0 = __funnelshift_r(a0, a1, sh), r1 = __funnelshift_r(a1, a2, sh);
    uint32_t r2 = __funnelshift_r(a2, a3, sh), r3 = __funnelshift_r(a3, a4, sh);
    uint32_t rr[4] = { r0, r1, r2, r3 };
#pragma unroll
    for (int c = 0; c < 4; c++) {
      uint32_t cw = curw[row * 4 + c];
      if (!SSE) {
        leaf[(row >> 2) * 4 + c] = __vsadu4(rr[c], cw) + leaf[(row >> 2) * 4 + c];
      } else {
        int s = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          int d = (int)((rr[c] >> (8 * k)) & 255) - (int)((cw >> (8 * k)) & 255);
          s += d * d;
        }
        leaf[(row >> 2) * 4 + c] += s;
      }
    }
  }
}

// SetupLargerBlocks: the 41 partition sums in ABI order from the 16 leaves
__device__ __forceinline__ void me_aggregate(const int leaf[16], int part[41])
{
  int t6[8], t5[8], t4[4];
#pragma unroll
  for (int i = 0; i < 4; i++) { t6[i] = leaf[i] + leaf[i + 4]; t6[4 + i] = leaf[8 + i] + leaf[12 + i]; }   // 4w x 8h at blocks 0..3, 8..11
#pragma unroll
  for (int i = 0; i < 8; i++) t5[i] = leaf[2 * i] + leaf[2 * i + 1];                                        // 8w x 4h at blocks 0,2,..,14
  t4[0] = t6[0] + t6[1]; t4[1] = t6[2] + t6[3]; t4[2] = t6[4] + t6[5]; t4[3] = t6[6] + t6[7];               // 8x8 at 0,2,8,10
  int t3a = t4[0] + t4[2], t3b = t4[1] + t4[3];                                                             // 8w x 16h at 0,2
  int t2a = t4[0] + t4[1], t2b = t4[2] + t4[3];                                                             // 16w x 8h at 0,8
  part[0] = t3a + t3b;
  part[1] = t2a; part[2] = t2b; part[3] = t3a; part[4] = t3b;
#pragma unroll
  for (int i = 0; i < 4; i++) part[5 + i] = t4[i];
#pragma unroll
  for (int i = 0; i < 8; i++) part[9 + i] = t5[i];
#pragma unroll
  for (int i = 0; i < 8; i++) part[17 + i] = t6[i];
#pragma unroll
  for (int i = 0; i < 16; i++) part[25 + i] = leaf[i];
}

// Pel = uint16_t: pictures of more than 8 bits per sample -- window and macroblock as 16-bit samples, scalar differences, the window
// read from plane (0,0) with clamped coordinates (no search-plane copy for such pictures), no chroma terms
template <bool SSE, typename Pel>
__global__ void __launch_bounds__(256)
me_ffs_batch_basic_kernel(const MeJob *__restrict__ jobs, int n_jobs, const MeRefDev *__restrict__ ref_table,
                          const Pel *__restrict__ cur, int cur_pitch, MeCfgDev cfg,
                          MeResultInt *__restrict__ out, const int *__restrict__ job_list)
{
  if (cfg.abort_flag && *cfg.abort_flag) return;
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ uint32_t curw[sizeof(Pel) == 1 ? 64 : 128];    // the macroblock: 64 words of four 8-bit samples, or 256 16-bit samples
  __shared__ unsigned long long red[8][ME_NUM_PARTITIONS];
  __shared__ int16_t pmv_s[ME_NUM_PARTITIONS][2];
  __shared__ MeJob sjob;
  // job_list == nullptr: CTA b takes job b.  Otherwise job_list[0] jobs, ids in job_list[1..] (the jobs the tuned
  // kernel of ffs_tuned.cu handed back), walked with a grid stride.
  const int total = job_list ? job_list[0] : n_jobs;
  for (int ji = blockIdx.x; ji < total; ji += gridDim.x) {
  const int job_id = job_list ? job_list[1 + ji] : ji;
  me_load_job(&sjob, jobs + job_id);
  __syncthreads();
  const MeJob &job = sjob;
  const MeRefDev ref = ref_table[job.list * ME_MAX_REFS + job.ref];
  const int R = me_setup_range(cfg, job.ref);
  const int n = 2 * R + 16, tp = sizeof(Pel) == 1 ? (n + 3) & ~3 : n + 1, side = 2 * R + 1;
  int cx, cy;
  me_search_centre(cfg, R, job.opix_x, job.opix_y, job.pmv16[0], job.pmv16[1], cx, cy);
  const int offx = cx - job.opix_x, offy = cy - job.opix_y;
  uint8_t *tile = smem;   // (n+1) rows of tp bytes (+ one spare row: the 5th word of the last window word may be read)
  Pel *tile16 = reinterpret_cast<Pel *>(smem);              // 16-bit samples: n rows of tp samples
  Pel *cur16 = reinterpret_cast<Pel *>(curw);               // ... and the macroblock as 256 samples (curw: 64 words; see below)

  if constexpr (sizeof(Pel) == 1) {
  me_load_window(tile, tp, n, ref, cx, cy, R);
  if (threadIdx.x < 64) {
    int row = threadIdx.x >> 2, c = threadIdx.x & 3;
    curw[threadIdx.x] = *reinterpret_cast<const uint32_t *>(cur + (size_t)(job.opix_y + row) * cur_pitch + job.opix_x + 4 * c);
  }
  } else {
    const Pel *plane = reinterpret_cast<const Pel *>(ref.planes);
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
      const int y = i / n, x = i - y * n;
      const int sy = me_clamp(cy - R + y + ME_PAD, 0, ref.hp - 1), sx = me_clamp(cx - R + x + ME_PAD, 0, ref.wp - 1);
      tile16[y * tp + x] = plane[(size_t)sy * ref.pitch + sx];
    }
    cur16[threadIdx.x] = cur[(size_t)(job.opix_y + (threadIdx.x >> 4)) * cur_pitch + job.opix_x + (threadIdx.x & 15)];
  }
  if (threadIdx.x < ME_NUM_PARTITIONS) { pmv_s[threadIdx.x][0] = job.pmv[threadIdx.x][0]; pmv_s[threadIdx.x][1] = job.pmv[threadIdx.x][1]; }
  __syncthreads();

  unsigned long long best[ME_NUM_PARTITIONS];
#pragma unroll
  for (int p = 0; p < ME_NUM_PARTITIONS; p++) best[p] = ~0ull;
  const int lam = job.lambda[0];
  // position of the zero vector inside the window (only used when !rdopt)
  const int dx00 = -offx + R, dy00 = -offy + R;

  for (int idx = threadIdx.x; idx < side * side; idx += blockDim.x) {
    int dy = idx / side, dx = idx - dy * side;
    int leaf[16], part[ME_NUM_PARTITIONS];
    if constexpr (sizeof(Pel) == 1) {
      me_leaf16<SSE>(tile, tp, curw, dx, dy, leaf);
      if (cfg.chroma.on && ref.chroma[0]) me_leaf_chroma(ref, cfg.chroma, SSE, job.opix_x, job.opix_y, cx - R + dx, cy - R + dy, leaf);
    } else {
#pragma unroll
      for (int b = 0; b < 16; b++) leaf[b] = 0;
      for (int row = 0; row < 16; row++)
#pragma unroll
        for (int col = 0; col < 16; col++) {
          const int d = (int)tile16[(dy + row) * tp + dx + col] - (int)cur16[row * 16 + col];
          leaf[(row >> 2) * 4 + (col >> 2)] += SSE ? d * d : abs(d);
        }
    }
    me_aggregate(leaf, part);
    int sidx = me_spiral_index(dx - R, dy - R);
    bool is00 = (!cfg.rdopt) && dx == dx00 && dy == dy00;
    unsigned tie = is00 ? 0u : (unsigned)(sidx + 1);
    int candx = (offx + dx - R) << 2, candy = (offy + dy - R) << 2;
#pragma unroll
    for (int p = 0; p < ME_NUM_PARTITIONS; p++) {
      int t = me_part_type(p);
      int rp = me_block_range(cfg, job.ref, t);
      int maxpos = (2 * rp + 1) * (2 * rp + 1);
      if (sidx < maxpos || is00) {
        // the zero vector is scored with candidate (0,0) whatever its position (me_fullfast.c:869)
        int mc = part[p] + me_mv_cost(lam, is00 ? 0 : candx, is00 ? 0 : candy, pmv_s[p][0], pmv_s[p][1]);
        unsigned long long key = ((unsigned long long)(unsigned)mc << 32) | tie;
        best[p] = key < best[p] ? key : best[p];
      }
    }
  }
  // reduce over the CTA
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int p = 0; p < ME_NUM_PARTITIONS; p++) {
    unsigned long long k = best[p];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
      k = other < k ? other : k;
    }
    if (lane == 0) red[warp][p] = k;
  }
  __syncthreads();
  if (threadIdx.x < ME_NUM_PARTITIONS) {
    int p = threadIdx.x;
    unsigned long long k = red[0][p];
    for (int wdx = 1; wdx < (int)(blockDim.x >> 5); wdx++) k = red[wdx][p] < k ? red[wdx][p] : k;
    int cost = (int)(unsigned)(k >> 32);
    unsigned tie = (unsigned)(k & 0xffffffffu);
    int dx, dy;
    if (tie == 0) { dx = -offx; dy = -offy; } else me_spiral_xy((int)tie - 1, dx, dy);
    out[job_id].mv[p][0] = (int16_t)(offx + dx);
    out[job_id].mv[p][1] = (int16_t)(offy + dy);
    out[job_id].cost[p] = cost;
  }
  __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------
// drop-in call sequence: surfaces once per (macroblock, list, ref) ...
template <bool SSE>
__global__ void __launch_bounds__(128)
me_ffs_surfaces_kernel(MeRefDev ref, const uint8_t *__restrict__ cur, int cur_pitch, int opix_x, int opix_y,
                       int cx, int cy, int R, MeChromaDev chroma, int32_t *__restrict__ surf /* [41][P] in spiral order */)
{
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ uint32_t curw[64];
  const int n = 2 * R + 16, tp = (n + 3) & ~3, side = 2 * R + 1, P = side * side;
  // each CTA needs only the window rows of its positions, but the window is small: load all of it
  me_load_window(smem, tp, n, ref, cx, cy, R);
  if (threadIdx.x < 64) {
    int row = threadIdx.x >> 2, c = threadIdx.x & 3;
    curw[threadIdx.x] = *reinterpret_cast<const uint32_t *>(cur + (size_t)(opix_y + row) * cur_pitch + opix_x + 4 * c);
  }
  __syncthreads();
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  int dy = idx / side, dx = idx - dy * side;
  int leaf[16], part[ME_NUM_PARTITIONS];
  me_leaf16<SSE>(smem, tp, curw, dx, dy, leaf);
  if (chroma.on && ref.chroma[0]) me_leaf_chroma(ref, chroma, SSE, opix_x, opix_y, cx - R + dx, cy - R + dy, leaf);
  me_aggregate(leaf, part);
  int s = me_spiral_index(dx - R, dy - R);
#pragma unroll
  for (int p = 0; p < ME_NUM_PARTITIONS; p++) surf[(size_t)p * P + s] = part[p];
}

// The same surfaces for pictures of more than 8 bits per sample (16-bit planes and current picture): plain scalar arithmetic, window
// read from plane (0,0) of the picture's plane set with clamped coordinates (its 20-sample border and the clamp are both pure edge
// replication, i.e. JM's origin clamp, refbuf.c:37).  No chroma terms (refused at init for such pictures).
template <bool SSE>
__global__ void __launch_bounds__(128)
me_ffs_surfaces16_kernel(MeRefDev ref, const uint16_t *__restrict__ cur, int cur_pitch, int opix_x, int opix_y,
                         int cx, int cy, int R, int32_t *__restrict__ surf /* [41][P] in spiral order */)
{
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ uint16_t curs[256];
  uint16_t *win = reinterpret_cast<uint16_t *>(smem);
  const int n = 2 * R + 16, tp = n + 1, side = 2 * R + 1, P = side * side;
  const uint16_t *plane = reinterpret_cast<const uint16_t *>(ref.planes);
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
    const int y = i / n, x = i - y * n;
    const int sy = me_clamp(cy - R + y + ME_PAD, 0, ref.hp - 1), sx = me_clamp(cx - R + x + ME_PAD, 0, ref.wp - 1);
    win[y * tp + x] = plane[(size_t)sy * ref.pitch + sx];
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) curs[i] = cur[(size_t)(opix_y + (i >> 4)) * cur_pitch + opix_x + (i & 15)];
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const int dy = idx / side, dx = idx - dy * side;
  int leaf[16], part[ME_NUM_PARTITIONS];
#pragma unroll
  for (int b = 0; b < 16; b++) leaf[b] = 0;
  for (int row = 0; row < 16; row++)
#pragma unroll
    for (int col = 0; col < 16; col++) {
      const int d = (int)win[(dy + row) * tp + dx + col] - (int)curs[row * 16 + col];
      leaf[(row >> 2) * 4 + (col >> 2)] += SSE ? d * d : abs(d);
    }
  me_aggregate(leaf, part);
  const int s = me_spiral_index(dx - R, dy - R);
#pragma unroll
  for (int p = 0; p < ME_NUM_PARTITIONS; p++) surf[(size_t)p * P + s] = part[p];
}

// ... then one argmin per block search (FastFullPelBlockMotionSearch, me_fullfast.c:833-903)
__global__ void __launch_bounds__(256)
me_ffs_block_kernel(const int32_t *__restrict__ surf_p, int setup_range, int offx, int offy, int pmv_x, int pmv_y,
                    int search_range, int min_mcost, int lambda, int rdopt, int *result /* mvx, mvy, cost, ticket */, int ticket)
{
  __shared__ unsigned long long red[8];
  int mvx, mvy, cost;
  me_ffs_block_argmin(surf_p, setup_range, offx, offy, pmv_x, pmv_y, search_range, min_mcost, lambda, rdopt, red, mvx, mvy, cost);
  if (threadIdx.x == 0) {
    result[0] = mvx; result[1] = mvy; result[2] = cost;
    __threadfence_system();                              // `result` is mapped host memory: the host spins on the ticket
    *(volatile int *)&result[3] = ticket;
  }
}

extern "C" void me_launch_ffs_batch_basic(const MeJob *jobs, int n_jobs, const MeRefDev *ref_table, const uint8_t *cur,
                                          int cur_pitch, const MeCfgDev *cfg, MeResultInt *out, const int *job_list,
                                          cudaStream_t st)
{
  int R = cfg->search_range, n = 2 * R + 16, tp = (n + 3) & ~3;
  size_t smem = (size_t)(n + 1) * tp + 16;
  int grid = job_list ? (n_jobs < 296 ? n_jobs : 296) : n_jobs;
  if (cfg->pel16) {
    const uint16_t *cur16 = reinterpret_cast<const uint16_t *>(cur);
    smem = (size_t)n * (n + 1) * sizeof(uint16_t);          // 41.8 KB at R = 64: inside the default limit
    if (cfg->metric[0])
      me_ffs_batch_basic_kernel<true, uint16_t><<<grid, 256, smem, st>>>(jobs, n_jobs, ref_table, cur16, cur_pitch, *cfg, out, job_list);
    else
      me_ffs_batch_basic_kernel<false, uint16_t><<<grid, 256, smem, st>>>(jobs, n_jobs, ref_table, cur16, cur_pitch, *cfg, out, job_list);
    return;
  }
  if (cfg->metric[0])
    me_ffs_batch_basic_kernel<true, uint8_t><<<grid, 256, smem, st>>>(jobs, n_jobs, ref_table, cur, cur_pitch, *cfg, out, job_list);
  else
    me_ffs_batch_basic_kernel<false, uint8_t><<<grid, 256, smem, st>>>(jobs, n_jobs, ref_table, cur, cur_pitch, *cfg, out, job_list);
}

extern "C" void me_launch_ffs_surfaces(const MeRefDev *ref, const uint8_t *cur, int cur_pitch, int opix_x, int opix_y,
                                       int cx, int cy, int R, int sse, const MeChromaDev *chroma, int32_t *surf, cudaStream_t st)
{
  int n = 2 * R + 16, tp = (n + 3) & ~3, P = (2 * R + 1) * (2 * R + 1);
  size_t smem = (size_t)(n + 1) * tp + 16;
  if (sse) me_ffs_surfaces_kernel<true><<<(P + 127) / 128, 128, smem, st>>>(*ref, cur, cur_pitch, opix_x, opix_y, cx, cy, R, *chroma, surf);
  else     me_ffs_surfaces_kernel<false><<<(P + 127) / 128, 128, smem, st>>>(*ref, cur, cur_pitch, opix_x, opix_y, cx, cy, R, *chroma, surf);
}

extern "C" void me_launch_ffs_surfaces16(const MeRefDev *ref, const uint16_t *cur, int cur_pitch, int opix_x, int opix_y,
                                         int cx, int cy, int R, int sse, int32_t *surf, cudaStream_t st)
{
  const int n = 2 * R + 16, P = (2 * R + 1) * (2 * R + 1);
  const size_t smem = (size_t)n * (n + 1) * sizeof(uint16_t);      // 41.8 KB at R = 64: inside the default limit
  if (sse) me_ffs_surfaces16_kernel<true><<<(P + 127) / 128, 128, smem, st>>>(*ref, cur, cur_pitch, opix_x, opix_y, cx, cy, R, surf);
  else     me_ffs_surfaces16_kernel<false><<<(P + 127) / 128, 128, smem, st>>>(*ref, cur, cur_pitch, opix_x, opix_y, cx, cy, R, surf);
}

extern "C" void me_launch_ffs_block(const int32_t *surf_p, int setup_range, int offx, int offy, int pmv_x, int pmv_y,
                                    int search_range, int min_mcost, int lambda, int rdopt, int *result, int ticket, cudaStream_t st)
{
  me_ffs_block_kernel<<<1, 256, 0, st>>>(surf_p, setup_range, offx, offy, pmv_x, pmv_y, search_range, min_mcost, lambda,
                                         rdopt, result, ticket);
}


// Validation of a job array the host has not read (pinned caller memory handed straight to the kernels): the fields the search
// kernels index with.  *bad = 1 + index of the first offending job (0: all fine).
__global__ void __launch_bounds__(256)
me_check_jobs_kernel(const MeJob *__restrict__ jobs, int n_jobs, int n_refs0, int n_refs1, int width, int height, int *bad)
{
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_jobs; i += gridDim.x * blockDim.x) {
    const uint32_t w0 = reinterpret_cast<const uint32_t *>(&jobs[i])[0], w1 = reinterpret_cast<const uint32_t *>(&jobs[i])[1];
    const int ox = (int)(int16_t)(w0 & 0xffffu), oy = (int)(int16_t)(w0 >> 16);
    const int list = (int)(int8_t)(w1 & 0xffu), ref = (int)(int8_t)((w1 >> 8) & 0xffu);
    const bool ok = list >= 0 && list <= 1 && ref >= 0 && ref < (list ? n_refs1 : n_refs0) &&
                    ox >= 0 && oy >= 0 && ox + 16 <= width && oy + 16 <= height && !(ox & 15) && !(oy & 15);
    if (!ok) atomicCAS(bad, 0, i + 1);
  }
}

extern "C" void me_launch_check_jobs(const MeJob *jobs, int n_jobs, int n_refs0, int n_refs1, int width, int height, int *bad, cudaStream_t st)
{
  int grid = (n_jobs + 255) / 256;
  if (grid > 148 * 4) grid = 148 * 4;
  me_check_jobs_kernel<<<grid, 256, 0, st>>>(jobs, n_jobs, n_refs0, n_refs1, width, height, bad);
}
