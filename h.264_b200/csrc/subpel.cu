// subpel.cu -- half / quarter-pel refinement (uni- and bi-predictive) and the bi-predictive integer search.
//
// Replaces SubPelBlockMotionSearch (reference lencod/src/me_fullsearch.c:341-511), SubPelBlockSearchBiPred (:520-754),
// FullPelBlockMotionBiPred (:164-332) and the distortion functions they call: computeSAD / computeSATD /
// computeBiPredSAD1 / computeBiPredSATD1 with HadamardSAD4x4 / HadamardSAD8x8 (lencod/src/me_distortion.c:351,657,482,
// 824,182,272).  Reference samples come from the 16 quarter-pel planes with JM's origin clamp (refbuf.c:37), applied per
// transform tile for SATD and per block for SAD exactly as the reference does.
//
// Work decomposition: a refinement stage scores 9 (or 8) candidates of every partition; the unit of work is one
// transform tile (4x4, or 8x8 when Transform8x8Mode && blocktype <= 4, mv-search.c:640) of one candidate.  Tiles
// accumulate into a per-(partition, candidate) cost in shared memory; one thread per partition then replays the
// reference's sequential "first strict minimum" scan (me_fullsearch.c:423-449, 482-502).  The reference's early exits
// (mcost >= min_mcost -> continue; partial sums > min_mcost -> return) never change the outcome and are dropped.
#include "me_common.cuh"

// spiral entries 0..8 (mv-search.c:366-393 with range 1)
__constant__ int8_t c_sp_x[9] = { 0, 0, 0, -1, 1, -1, 1, -1, 1 };
__constant__ int8_t c_sp_y[9] = { 0, -1, 1, -1, -1, 0, 0, 1, 1 };

__device__ __forceinline__ int hadamard4(const int d[16])
{
  int t[16], s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int a = d[4 * i], b = d[4 * i + 1], c = d[4 * i + 2], e = d[4 * i + 3];
    int p = a + e, q = b + c, u = a - e, v = b - c;
    t[4 * i] = p + q; t[4 * i + 1] = p - q; t[4 * i + 2] = u + v; t[4 * i + 3] = u - v;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int a = t[i], b = t[4 + i], c = t[8 + i], e = t[12 + i];
    int p = a + e, q = b + c, u = a - e, v = b - c;
    s += abs(p + q) + abs(p - q) + abs(u + v) + abs(u - v);
  }
  return (s + 1) >> 1;                                   // me_distortion.c:263
}

__device__ __forceinline__ void wht8(int v[8])
{
#pragma unroll
  for (int stage = 4; stage >= 1; stage >>= 1)
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (!(k & stage)) { int a = v[k], b = v[k + stage]; v[k] = a + b; v[k + stage] = a - b; }
}

__device__ __noinline__ int hadamard8(int d[64])
{
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) wht8(&d[8 * i]);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    int col[8];
#pragma unroll
    for (int k = 0; k < 8; k++) col[k] = d[8 * k + i];
    wht8(col);
#pragma unroll
    for (int k = 0; k < 8; k++) s += abs(col[k]);
  }
  return (s + 2) >> 2;                                   // me_distortion.c:342
}

// Weighted bi-prediction (computeBiPred*2, me_distortion.c:556, :907, :1254): what FullPelBlockMotionBiPred /
// SubPelBlockSearchBiPred derive from their `list` argument (me_fullsearch.c:202-234, :538-590).  Same layout in me_cuda_host.cu.
struct MeBiW { int apply, w1, w2, obi, denom; int cw1[2], cw2[2], cobi[2]; int maxv; };   // c*: the same for the two chroma components; maxv: img->max_imgpel_value (set with apply)

// prediction sample of a bi-predictive distortion: (p1 + p2 + 1) >> 1 (me_distortion.c:502), or with weights
// iClip1(255, ((weight1 * p1 + weight2 * p2 + 2 * wp_luma_round) >> (denom + 1)) + offsetBi)
__device__ __forceinline__ int bipred_pel(int p1, int p2, const MeBiW &w)
{
  if (!w.apply) return (p1 + p2 + 1) >> 1;
  return me_clamp(((w.w1 * p1 + w.w2 * p2 + (w.denom ? 1 << w.denom : 0)) >> (w.denom + 1)) + w.obi, 0, w.maxv);
}

// distortion of one N x N tile.  metric: 0 SAD, 1 SSE, 2 SATD.  p2 != nullptr: bi-prediction.
// blk / bsx / bx / by: origin of the block in the current picture, its width, the tile's offset inside it.  They matter for one
// case only: the 8x8 branch of computeBiPredSATD2 reads the eighth sample of every row without advancing its source pointer
// (me_distortion.c:1015), so each row of the ORIGINAL block starts one sample further left than the row above -- linear
// index by*bsx + bx + y*(bsx-1) + x into the bsx-wide copy of the block (mv-search.c:607-611).  Reproduced as is.
// Pel: uint8_t, or uint16_t for pictures of more than 8 bits per sample (the drop-in kernels are instantiated for both)
template <int N, typename Pel>
__device__ __forceinline__ int tile_dist(int metric, const Pel *__restrict__ blk, int src_pitch, int bsx, int bx, int by,
                                         const Pel *__restrict__ p1, int pitch1,
                                         const Pel *__restrict__ p2, int pitch2, const MeBiW &bw)
{
  int d[N * N];
  const Pel *src = blk + by * src_pitch + bx;
  const bool skew = N == 8 && metric == 2 && p2 && bw.apply;
#pragma unroll
  for (int y = 0; y < N; y++)
#pragma unroll
    for (int x = 0; x < N; x++) {
      int pred = p1[y * pitch1 + x];
      if (p2) pred = bipred_pel(pred, p2[y * pitch2 + x], bw);
      int s;
      if (skew) { const int L = by * bsx + bx + y * (bsx - 1) + x; s = blk[(L / bsx) * src_pitch + (L % bsx)]; }
      else s = src[y * src_pitch + x];
      d[y * N + x] = s - pred;
    }
  if (metric == 2) {
    if (N == 4) return hadamard4(d);
    return hadamard8(d);
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < N * N; i++) s += metric ? d[i] * d[i] : abs(d[i]);
  return s;
}

// The reference fetches SAD blocks with one origin clamp per block and SATD tiles with one per tile; both are served by
// this helper: (bx,by) = tile offset inside the block, `per_tile` selects where the clamp applies.
template <typename Pel>
__device__ __forceinline__ const Pel *tile_ptr(const MeRefDev &r, int x4, int y4, int bx, int by, bool per_tile)
{
  if (per_tile) return me_umv_ptr_p<Pel>(r, y4 + (by << 2), x4 + (bx << 2));
  return me_umv_ptr_p<Pel>(r, y4, x4) + (size_t)by * r.pitch + bx;
}

// ------------------------------------------------------------------------------------------------------------------
// batch: one CTA per (macroblock, list, ref) job, all 41 partitions
template <typename Pel>
__global__ void __launch_bounds__(256)
me_subpel_batch_kernel(const MeJob *__restrict__ jobs, int n_jobs, const MeRefDev *__restrict__ ref_table,
                       const Pel *__restrict__ cur, int cur_pitch, MeCfgDev cfg,
                       const MeResultInt *__restrict__ in_int, MeResult *__restrict__ out)
{
  if (cfg.abort_flag && *cfg.abort_flag) return;
  __shared__ int cost_s[ME_NUM_PARTITIONS][9];
  __shared__ int mv_s[ME_NUM_PARTITIONS][2];
  __shared__ int min_s[ME_NUM_PARTITIONS];
  __shared__ Pel cur_s[16][16];
  __shared__ MeJob sjob;
  const int job_id = blockIdx.x;
  if (job_id >= n_jobs) return;
  me_load_job(&sjob, jobs + job_id);
  __syncthreads();
  const MeJob &job = sjob;
  const MeRefDev ref = ref_table[job.list * ME_MAX_REFS + job.ref];
  const int tid = threadIdx.x;

  cur_s[tid >> 4][tid & 15] = cur[(size_t)(job.opix_y + (tid >> 4)) * cur_pitch + job.opix_x + (tid & 15)];
  if (tid < ME_NUM_PARTITIONS) {
    // one word = (mv_x, mv_y); explicit word load + split (see me_load_job for why 16-bit global loads are avoided)
    uint32_t w = reinterpret_cast<const uint32_t *>(in_int[job_id].mv)[tid];
    mv_s[tid][0] = (int)(int16_t)(w & 0xffffu) * 4;        // mv <<= 2, mv-search.c:770-774
    mv_s[tid][1] = (int)(int16_t)(w >> 16) * 4;
    min_s[tid] = cfg.start_hp ? in_int[job_id].cost[tid] : ME_INT_MAX;   // mv-search.c:785-788
  }
  __syncthreads();

  for (int stage = 0; stage < 2; stage++) {               // 0: half-pel (step 2), 1: quarter-pel (step 1)
    const int step = stage == 0 ? 2 : 1;
    const int level = stage == 0 ? 1 : 2;
    const int metric = cfg.metric[level];
    const int lam = job.lambda[level];
    const int first = stage == 0 ? cfg.start_hp : cfg.start_qp;
    if (stage == 1 && !cfg.start_qp && tid < ME_NUM_PARTITIONS) min_s[tid] = ME_INT_MAX;   // me_fullsearch.c:456-457
    for (int i = tid; i < ME_NUM_PARTITIONS * 9; i += blockDim.x) {
      int p = i / 9, c = i - 9 * p;
      cost_s[p][c] = me_mv_cost(lam, mv_s[p][0] + step * c_sp_x[c], mv_s[p][1] + step * c_sp_y[c], job.pmv[p][0], job.pmv[p][1]);
    }
    __syncthreads();
    // tiles: for every block type the macroblock is covered by 16 4x4 tiles (or 4 8x8 tiles)
    for (int i = tid; i < 7 * 16 * 9; i += blockDim.x) {
      int c = i % 9, rest = i / 9;
      int tile = rest & 15, t = (rest >> 4) + 1;
      if (c < first) continue;
      bool use8 = cfg.transform8x8 && t <= 4 && metric == 2;
      int tx = tile & 3, ty = tile >> 2;                  // tile position in 4-pel units inside the macroblock
      if (use8 && ((tx | ty) & 1)) continue;
      int bw = me_bsx(t) >> 2, bh = me_bsy(t) >> 2;       // block size in 4-pel units
      int px = tx / bw * bw, py = ty / bh * bh;           // origin of the partition that contains the tile
      int p = me_part_index(t, py * 4 + px);
      int x4 = ((job.opix_x + 4 * px + ME_PAD) << 2) + mv_s[p][0] + step * c_sp_x[c];
      int y4 = ((job.opix_y + 4 * py + ME_PAD) << 2) + mv_s[p][1] + step * c_sp_y[c];
      int bx = 4 * (tx - px), by = 4 * (ty - py);
      const Pel *rp = tile_ptr<Pel>(ref, x4, y4, bx, by, metric == 2);
      const Pel *sp = &cur_s[4 * ty][4 * tx];
      const MeBiW nobw = { 0, 0, 0, 0, 0, { 0, 0 }, { 0, 0 }, { 0, 0 } };
      int d = use8 ? tile_dist<8, Pel>(metric, sp, 16, 16, 0, 0, rp, ref.pitch, nullptr, 0, nobw)
                   : tile_dist<4, Pel>(metric, sp, 16, 16, 0, 0, rp, ref.pitch, nullptr, 0, nobw);
      // chroma terms of computeSAD / computeSSE (me_distortion.c:376-402, :1069-1091): once per block, by the lane of its first tile
      if constexpr (sizeof(Pel) == 1)
      if (cfg.chroma.on && metric != 2 && bx == 0 && by == 0 && ref.chroma[0])
        for (int k = 0; k < 2; k++)
          d += cfg.chroma.weight * me_cr_block(cfg.chroma, metric, k, job.opix_x + 4 * px, job.opix_y + 4 * py, me_bsx(t), me_bsy(t),
                                               ref, x4, y4, nullptr, 0, 0);
      atomicAdd(&cost_s[p][c], d);
    }
    __syncthreads();
    if (tid < ME_NUM_PARTITIONS) {
      int p = tid, best = 0, mn = min_s[p];
      for (int c = first; c < 9; c++) {
        int mc = cost_s[p][c];
        if (stage == 0 && c == 0 && !cfg.rdopt && job.slice_type != ME_B_SLICE && job.ref == 0 && p == 0 &&
            mv_s[p][0] == 0 && mv_s[p][1] == 0)
          mc -= (lam * 16) >> 16;                         // check_position0, me_fullsearch.c:361,439-442
        if (mc < mn) { mn = mc; best = c; }
      }
      min_s[p] = mn;
      mv_s[p][0] += step * c_sp_x[best];
      mv_s[p][1] += step * c_sp_y[best];
    }
    __syncthreads();
  }
  if (tid < ME_NUM_PARTITIONS) {
    out[job_id].mv[tid][0] = (int16_t)mv_s[tid][0];
    out[job_id].mv[tid][1] = (int16_t)mv_s[tid][1];
    out[job_id].cost[tid] = min_s[tid];
  }
}

// ------------------------------------------------------------------------------------------------------------------
// single block, uni- or bi-predictive (drop-in calls).  One CTA.
//   bi == 0: SubPelBlockMotionSearch.  bi == 1: SubPelBlockSearchBiPred with r1 read at the fixed vector (s_mv).
struct MeSubpelArgs {
  int pic_pix_x, pic_pix_y, blocktype;
  int pmv_x, pmv_y, mv_x, mv_y, s_mv_x, s_mv_y;
  int search_pos2, search_pos4, min_mcost;
  int lambda[3];
  int check0;           // check_position0 already evaluated by the host side (needs slice type / ref)
  int bi;
  MeBiW bw;             // bi == 1: weights of this call's `list`
};

// scratch of the block-level device functions below (one per CTA)
struct BlockScratch {
  unsigned long long red[8];
  int cost_s[9];
  int mv_s[2], min_sh;
  int grid[49], done;      // the pattern searches' candidate costs ((2E+1)^2 positions, E <= 3) and their early-return flag
};

// One block, both sub-pel levels.  bi == 0: SubPelBlockMotionSearch (r1 searched).  bi == 1: SubPelBlockSearchBiPred, r1 read at
// the fixed vector (s_mv), r2 at the candidate.  Called by all threads of the CTA; returns the result to all of them.
template <typename Pel>
__device__ void subpel_block(const MeRefDev &r1, const MeRefDev &r2, const Pel *__restrict__ cur, int cur_pitch,
                             const MeCfgDev &cfg, const MeSubpelArgs &a, BlockScratch &sc, int &out_mvx, int &out_mvy, int &out_cost)
{
  const int tid = threadIdx.x;
  const int t = a.blocktype, bsx = me_bsx(t), bsy = me_bsy(t);
  __syncthreads();                      // the scratch may still be read from a previous call
  if (tid == 0) { sc.mv_s[0] = a.mv_x; sc.mv_s[1] = a.mv_y; sc.min_sh = a.min_mcost; }
  __syncthreads();
  for (int stage = 0; stage < 2; stage++) {
    const int step = stage == 0 ? 2 : 1, level = stage == 0 ? 1 : 2;
    const int metric = cfg.metric[level], lam = a.lambda[level];
    const int first = stage == 0 ? cfg.start_hp : cfg.start_qp;
    int npos = stage == 0 ? a.search_pos2 : a.search_pos4;
    if (stage == 0 && !cfg.start_hp && npos < 1) npos = 1;   // max_pos2, me_fullsearch.c:366
    if (npos > 9) npos = 9;
    if (tid == 0 && stage == 1 && !cfg.start_qp) sc.min_sh = ME_INT_MAX;
    if (tid < 9) sc.cost_s[tid] = me_mv_cost(lam, sc.mv_s[0] + step * c_sp_x[tid], sc.mv_s[1] + step * c_sp_y[tid], a.pmv_x, a.pmv_y);
    __syncthreads();
    bool use8 = cfg.transform8x8 && t <= 4 && metric == 2;
    int n = use8 ? 8 : 4, tw = bsx / n, th = bsy / n;
    for (int i = tid; i < tw * th * 9; i += blockDim.x) {
      int c = i % 9, tile = i / 9;
      if (c < first || c >= npos) continue;
      int bx = (tile % tw) * n, by = (tile / tw) * n;
      int x4 = ((a.pic_pix_x + ME_PAD) << 2) + sc.mv_s[0] + step * c_sp_x[c];
      int y4 = ((a.pic_pix_y + ME_PAD) << 2) + sc.mv_s[1] + step * c_sp_y[c];
      const Pel *blk = cur + (size_t)a.pic_pix_y * cur_pitch + a.pic_pix_x;
      int d;
      if (!a.bi) {
        const Pel *rp = tile_ptr<Pel>(r1, x4, y4, bx, by, metric == 2);
        d = use8 ? tile_dist<8, Pel>(metric, blk, cur_pitch, bsx, bx, by, rp, r1.pitch, nullptr, 0, a.bw)
                 : tile_dist<4, Pel>(metric, blk, cur_pitch, bsx, bx, by, rp, r1.pitch, nullptr, 0, a.bw);
      } else {
        int fx4 = ((a.pic_pix_x + ME_PAD) << 2) + a.s_mv_x, fy4 = ((a.pic_pix_y + ME_PAD) << 2) + a.s_mv_y;
        const Pel *p1 = tile_ptr<Pel>(r1, fx4, fy4, bx, by, metric == 2);
        const Pel *p2 = tile_ptr<Pel>(r2, x4, y4, bx, by, metric == 2);
        d = use8 ? tile_dist<8, Pel>(metric, blk, cur_pitch, bsx, bx, by, p1, r1.pitch, p2, r2.pitch, a.bw)
                 : tile_dist<4, Pel>(metric, blk, cur_pitch, bsx, bx, by, p1, r1.pitch, p2, r2.pitch, a.bw);
        if constexpr (sizeof(Pel) == 1)                     // chroma terms: 8-bit pictures only (refused at init otherwise)
        if (cfg.chroma.on && metric != 2 && bx == 0 && by == 0 && r1.chroma[0] && r2.chroma[0])      // computeBiPredSAD1 :516-543
          for (int k = 0; k < 2; k++)
            d += cfg.chroma.weight * me_cr_block(cfg.chroma, metric, k, a.pic_pix_x, a.pic_pix_y, bsx, bsy, r1, fx4, fy4, &r2, x4, y4,
                                                 a.bw.apply, a.bw.cw1[k], a.bw.cw2[k], a.bw.cobi[k], a.bw.denom);
      }
      if constexpr (sizeof(Pel) == 1)
      if (!a.bi && cfg.chroma.on && metric != 2 && bx == 0 && by == 0 && r1.chroma[0])                // computeSAD / SSE :376-402
        for (int k = 0; k < 2; k++)
          d += cfg.chroma.weight * me_cr_block(cfg.chroma, metric, k, a.pic_pix_x, a.pic_pix_y, bsx, bsy, r1, x4, y4, nullptr, 0, 0);
      atomicAdd(&sc.cost_s[c], d);
    }
    __syncthreads();
    if (tid == 0) {
      int best = 0, mn = sc.min_sh;
      for (int c = first; c < npos; c++) {
        int mc = sc.cost_s[c];
        if (stage == 0 && c == 0 && a.check0) mc -= (lam * 16) >> 16;
        if (mc < mn) { mn = mc; best = c; }
      }
      sc.min_sh = mn;
      sc.mv_s[0] += step * c_sp_x[best];
      sc.mv_s[1] += step * c_sp_y[best];
    }
    __syncthreads();
  }
  out_mvx = sc.mv_s[0]; out_mvy = sc.mv_s[1]; out_cost = sc.min_sh;
}

// ------------------------------------------------------------------------------------------------------------------
// Sub-pel stages of the pattern searches (SearchMode 1 UMHexagonS, 2 simplified UMHexagonS, 3 EPZS):
// EPZSSubPelBlockMotionSearch (me_epzs.c:2390-2720), EPZSSubPelBlockSearchBiPred (:2728-3125), UMHEXSubPelBlockMotionSearch
// (me_umhex.c:562-730), smpUMHEXSubPelBlockMotionSearch / smpUMHEXFullSubPelBlockMotionSearch (me_umhexsmp.c:616-812, :422-608).
// The reference walks a few data-dependent positions one after the other; here every position the walk can reach is scored at
// once (the 3 x 3 grid of a half- or quarter-pel level, the 7 x 7 quarter-pel grid of the UMHexagonS walks: one transform tile of
// one candidate per thread, as above) and one thread then replays the walk over the table of costs.  Same outcome: the thresholds
// the reference hands to its distortion functions on these paths only cut candidates that the comparison after the call rejects.
struct MePatArgs {
  int pattern;              // 1 EPZS, 2 UMHexagonS, 3 simplified UMHexagonS, 4 simplified UMHexagonS "full"
  int subthres;             // EPZS: subthres[blocktype] (me_epzs.c:354)
  int uplayer_x, uplayer_y; // simplified UMHexagonS: smpUMHEX_pred_MV_uplayer_X / _Y (me_umhexsmp.c:1203-1214)
  int thr1, thr3;           // ... SubPelThreshold1 / 3 >> block_type_shift_factor[blocktype] (:106-107, :46)
  int pmv2_x, pmv2_y;       // EPZS bi-predictive: predictor of the fixed vector (its rate is part of every candidate's cost)
};

// search_point_qp (me_epzs.c:44; search_point_hp, :43, is twice that); entry 9 repeats entry 5
__constant__ int8_t c_epzs_x[10] = { 0, -1, 0, 1, 0, -1, 1, 1, -1, -1 };
__constant__ int8_t c_epzs_y[10] = { 0, 0, 1, 0, -1, 1, 1, -1, -1, 1 };

// costs (vector rate + distortion) of the (2E+1)^2 candidates centre + step * (dx, dy), |dx|, |dy| <= E, into sc.grid[(dy+E)*(2E+1) + dx+E].
// Called by all threads; ends with a barrier.
template <typename Pel>
__device__ void pattern_grid(const MeRefDev &r1, const MeRefDev &r2, const Pel *__restrict__ cur, int cur_pitch, const MeCfgDev &cfg,
                             const MeSubpelArgs &a, const MePatArgs &pa, BlockScratch &sc, int level, int lam, int step, int E)
{
  const int tid = threadIdx.x, W = 2 * E + 1, ncand = W * W;
  const int t = a.blocktype, bsx = me_bsx(t), bsy = me_bsy(t);
  const int metric = cfg.metric[level];
  const int cx0 = sc.mv_s[0], cy0 = sc.mv_s[1];
  __syncthreads();                      // the grid may still be read by the walk of the previous level
  for (int c = tid; c < ncand; c += blockDim.x) {
    const int mx = cx0 + step * (c % W - E), my = cy0 + step * (c / W - E);
    int m = me_mv_cost(lam, mx, my, a.pmv_x, a.pmv_y);
    if (a.bi) m += me_mv_cost(lam, a.s_mv_x, a.s_mv_y, pa.pmv2_x, pa.pmv2_y);     // me_epzs.c:2867, :3021
    sc.grid[c] = m;
  }
  __syncthreads();
  const bool use8 = cfg.transform8x8 && t <= 4 && metric == 2;
  const int n = use8 ? 8 : 4, tw = bsx / n, th = bsy / n;
  const Pel *blk = cur + (size_t)a.pic_pix_y * cur_pitch + a.pic_pix_x;
  for (int i = tid; i < tw * th * ncand; i += blockDim.x) {
    const int c = i % ncand, tile = i / ncand;
    const int bx = (tile % tw) * n, by = (tile / tw) * n;
    const int x4 = ((a.pic_pix_x + ME_PAD) << 2) + cx0 + step * (c % W - E);
    const int y4 = ((a.pic_pix_y + ME_PAD) << 2) + cy0 + step * (c / W - E);
    int d;
    if (!a.bi) {
      const Pel *rp = tile_ptr<Pel>(r1, x4, y4, bx, by, metric == 2);
      d = use8 ? tile_dist<8, Pel>(metric, blk, cur_pitch, bsx, bx, by, rp, r1.pitch, nullptr, 0, a.bw)
               : tile_dist<4, Pel>(metric, blk, cur_pitch, bsx, bx, by, rp, r1.pitch, nullptr, 0, a.bw);
      if constexpr (sizeof(Pel) == 1)
      if (cfg.chroma.on && metric != 2 && bx == 0 && by == 0 && r1.chroma[0])                        // computeSAD / SSE, me_distortion.c:376-402
        for (int k = 0; k < 2; k++)
          d += cfg.chroma.weight * me_cr_block(cfg.chroma, metric, k, a.pic_pix_x, a.pic_pix_y, bsx, bsy, r1, x4, y4, nullptr, 0, 0);
    } else {
      const int fx4 = ((a.pic_pix_x + ME_PAD) << 2) + a.s_mv_x, fy4 = ((a.pic_pix_y + ME_PAD) << 2) + a.s_mv_y;
      const Pel *p1 = tile_ptr<Pel>(r1, fx4, fy4, bx, by, metric == 2);
      const Pel *p2 = tile_ptr<Pel>(r2, x4, y4, bx, by, metric == 2);
      d = use8 ? tile_dist<8, Pel>(metric, blk, cur_pitch, bsx, bx, by, p1, r1.pitch, p2, r2.pitch, a.bw)
               : tile_dist<4, Pel>(metric, blk, cur_pitch, bsx, bx, by, p1, r1.pitch, p2, r2.pitch, a.bw);
      if constexpr (sizeof(Pel) == 1)
      if (cfg.chroma.on && metric != 2 && bx == 0 && by == 0 && r1.chroma[0] && r2.chroma[0])         // computeBiPredSAD1 :516-543
        for (int k = 0; k < 2; k++)
          d += cfg.chroma.weight * me_cr_block(cfg.chroma, metric, k, a.pic_pix_x, a.pic_pix_y, bsx, bsy, r1, fx4, fy4, &r2, x4, y4,
                                               a.bw.apply, a.bw.cw1[k], a.bw.cw2[k], a.bw.cobi[k], a.bw.denom);
    }
    atomicAdd(&sc.grid[c], d);
  }
  __syncthreads();
}

// one level of the EPZS refinements over the 3 x 3 table (one thread): the diamond points (and the centre from `first`), then the
// corner(s) the best two point at (me_epzs.c:2470-2580, :2603-2712).  hp: the half-pel level, whose `case 0` evaluates nothing.
// early_thr >= 0: the uni-predictive early return (:2496, :2629).  Returns true when that return was taken.
__device__ bool epzs_level_walk(const int *grid, int step, bool hp, int first, int end_default, int early_thr, int pmv_x, int pmv_y,
                                int &mvx, int &mvy, int &min_mcost)
{
  int best = 0, second = 0, second_cost = ME_INT_MAX, start = 5, end = end_default;
  for (int pos = first; pos < 5; pos++) {
    const int mc = grid[(c_epzs_y[pos] + 1) * 3 + c_epzs_x[pos] + 1];
    if (mc < min_mcost) { second_cost = min_mcost; second = best; min_mcost = mc; best = pos; }
    else if (mc < second_cost) { second_cost = mc; second = pos; }
  }
  if (early_thr >= 0 && best == 0 && pmv_x == mvx && pmv_y == mvy && min_mcost < early_thr) return true;
  if (best != 0 && second != 0) {
    switch (best ^ second) {
      case 1: start = 6; end = 7; break;
      case 3: start = 5; end = 6; break;
      case 5: start = 8; end = 9; break;
      case 7: start = 7; end = 8; break;
      default: break;
    }
  } else {
    switch (best + second) {
      case 0: if (hp) { start = 5; end = 5; } break;
      case 1: start = 8; end = 10; break;
      case 2: start = 5; end = 7; break;
      case 5: start = 6; end = 8; break;
      case 7: start = 7; end = 9; break;
      default: break;
    }
  }
  if (end > 10) end = 10;
  if (best != 0 || pmv_x != mvx || pmv_y != mvy)
    for (int pos = start; pos < end; pos++) {
      const int mc = grid[(c_epzs_y[pos] + 1) * 3 + c_epzs_x[pos] + 1];
      if (mc < min_mcost) { min_mcost = mc; best = pos; }
    }
  mvx += step * c_epzs_x[best]; mvy += step * c_epzs_y[best];
  return false;
}

// the quarter-pel diamond walks of the UMHexagonS searches over the 7 x 7 table (one thread; me_umhex.c:640-727, me_umhexsmp.c:700-811)
__device__ void umhex_walk(const int *grid, bool smp, bool start_hp, const MePatArgs &pa, int pmv_x, int pmv_y, int &mvx, int &mvy, int &min_mcost)
{
  const int R = 3, mx0 = mvx, my0 = mvy;
  unsigned long long seen = 1ull << (R * 7 + R);
  int curx = 0, cury = 0;
  const int fx = (pmv_x - mx0) % 4, fy = (pmv_y - my0) % 4;                      // C remainder: the sign of the dividend
  const int ux = (pa.uplayer_x - mx0) % 4, uy = (pa.uplayer_y - my0) % 4;
  if (!start_hp) {
    const int mc = grid[R * 7 + R];
    if (mc < min_mcost) { min_mcost = mc; curx = mx0; cury = my0; }               // otherwise the walk starts from the vector (0, 0), as in the reference
  } else { curx = mx0; cury = my0; }
  if (smp && mx0 == 0 && my0 == 0 && fx == 0 && ux == 0 && fy == 0 && uy == 0 && min_mcost < pa.thr1) { mvx = curx; mvy = cury; return; }
  if (fx || fy) {
    const int g = (fy + R) * 7 + fx + R, mc = grid[g];
    seen |= 1ull << g;
    if (mc < min_mcost) { min_mcost = mc; curx = mx0 + fx; cury = my0 + fy; }
  }
  for (int i = 0; i < R; i++) {
    const int nx = curx, ny = cury;
    bool abort_search = true;
    for (int m = 0; m < 4; m++) {
      // diamond order: (-1,0) (0,1) (1,0) (0,-1) in me_umhex.c:578-579, (-1,0) (1,0) (0,-1) (0,1) in me_umhexsmp.c:39-40
      const int ddx = smp ? (m == 0 ? -1 : m == 1 ? 1 : 0) : (m == 0 ? -1 : m == 2 ? 1 : 0);
      const int ddy = smp ? (m == 2 ? -1 : m == 3 ? 1 : 0) : (m == 1 ? 1 : m == 3 ? -1 : 0);
      const int ox = nx + ddx - mx0, oy = ny + ddy - my0;
      if (ox < -R || ox > R || oy < -R || oy > R) continue;
      const int g = (oy + R) * 7 + ox + R;
      if ((seen >> g) & 1) continue;
      seen |= 1ull << g;
      const int mc = grid[g];
      if (mc < min_mcost) { min_mcost = mc; curx = nx + ddx; cury = ny + ddy; abort_search = false; }
      if (smp && min_mcost < pa.thr3) { mvx = curx; mvy = cury; return; }
    }
    if (abort_search) break;
  }
  mvx = curx; mvy = cury;
}

// One block through the sub-pel stage of a pattern search.  Called by all threads of the CTA; returns the result to all of them.
template <typename Pel>
__device__ void pattern_block(const MeRefDev &r1, const MeRefDev &r2, const Pel *__restrict__ cur, int cur_pitch, const MeCfgDev &cfg,
                              const MeSubpelArgs &a, const MePatArgs &pa, BlockScratch &sc, int &out_mvx, int &out_mvy, int &out_cost)
{
  const int tid = threadIdx.x;
  __syncthreads();
  if (tid == 0) { sc.mv_s[0] = a.mv_x; sc.mv_s[1] = a.mv_y; sc.min_sh = a.min_mcost; sc.done = 0; }
  __syncthreads();
  if (pa.pattern == 2 || pa.pattern == 3) {
    pattern_grid(r1, r2, cur, cur_pitch, cfg, a, pa, sc, 2, a.lambda[2], 1, 3);
    if (tid == 0) {
      int mvx = sc.mv_s[0], mvy = sc.mv_s[1], mn = sc.min_sh;
      umhex_walk(sc.grid, pa.pattern == 3, cfg.start_hp != 0, pa, a.pmv_x, a.pmv_y, mvx, mvy, mn);
      sc.mv_s[0] = (int16_t)mvx; sc.mv_s[1] = (int16_t)mvy; sc.min_sh = mn;       // *mv_x = (short) currmv_x
    }
  } else {
    for (int stage = 0; stage < 2; stage++) {
      const int level = pa.pattern == 4 ? 2 : stage + 1, lam = a.lambda[level], step = stage == 0 ? 2 : 1;
      pattern_grid(r1, r2, cur, cur_pitch, cfg, a, pa, sc, level, lam, step, 1);
      if (tid == 0) {
        int mvx = sc.mv_s[0], mvy = sc.mv_s[1], mn = sc.min_sh;
        if (pa.pattern == 1) {
          // EPZSSubPelBlockMotionSearch / EPZSSubPelBlockSearchBiPred
          int first, end_default;
          if (stage == 0) {
            first = a.bi ? (a.min_mcost == ME_INT_MAX ? 0 : cfg.start_hp) : cfg.start_hp;             // me_epzs.c:2760
            end_default = ((!cfg.start_hp || !cfg.start_qp) && a.search_pos2 < 1) ? 1 : a.search_pos2; // :2411, :2761
          } else {
            if (!cfg.start_qp) mn = ME_INT_MAX;                                                        // :2581, :3008
            first = cfg.start_qp; end_default = a.search_pos4;
          }
          if (epzs_level_walk(sc.grid, step, stage == 0, first, end_default, a.bi ? -1 : pa.subthres, a.pmv_x, a.pmv_y, mvx, mvy, mn)) sc.done = 1;
          mvx = (int16_t)mvx; mvy = (int16_t)mvy;                                                      // the vector lives in shorts
        } else {
          // smpUMHEXFullSubPelBlockMotionSearch: the plain nine positions per level, quarter-pel metric and factor for both, each
          // loop left as soon as the minimum is below SubPelThreshold3 (me_umhexsmp.c:504-540, :563-600)
          const int first = stage == 0 ? cfg.start_hp : cfg.start_qp;
          int npos = stage == 0 ? a.search_pos2 : a.search_pos4;
          if (stage == 0 && !cfg.start_hp && npos < 1) npos = 1;
          if (npos > 9) npos = 9;
          if (stage == 1 && !cfg.start_qp) mn = ME_INT_MAX;
          int best = 0;
          for (int c = first; c < npos; c++) {
            if (me_mv_cost(lam, mvx + step * c_sp_x[c], mvy + step * c_sp_y[c], a.pmv_x, a.pmv_y) >= mn) continue;   // skips the threshold test too (:519)
            int mc = sc.grid[(c_sp_y[c] + 1) * 3 + c_sp_x[c] + 1];
            if (stage == 0 && c == 0 && a.check0) mc -= (lam * 16) >> 16;
            if (mc < mn) { mn = mc; best = c; }
            if (mn < pa.thr3) break;
          }
          mvx = (int16_t)(mvx + step * c_sp_x[best]); mvy = (int16_t)(mvy + step * c_sp_y[best]);
          if (stage == 0 && mvx == 0 && mvy == 0 && a.pmv_x == 0 && a.pmv_y == 0 && mn < pa.thr1) sc.done = 1;       // :546-551
        }
        sc.mv_s[0] = mvx; sc.mv_s[1] = mvy; sc.min_sh = mn;
      }
      __syncthreads();
      if (sc.done) break;
    }
  }
  __syncthreads();
  out_mvx = sc.mv_s[0]; out_mvy = sc.mv_s[1]; out_cost = sc.min_sh;
}

template <typename Pel>
__global__ void __launch_bounds__(256)
me_subpel_pattern_kernel(MeRefDev r1, MeRefDev r2, const Pel *__restrict__ cur, int cur_pitch, MeCfgDev cfg,
                         MeSubpelArgs a, MePatArgs pa, int *result /* mvx, mvy, cost, ticket */, int ticket)
{
  __shared__ BlockScratch sc;
  int mvx, mvy, cost;
  pattern_block(r1, r2, cur, cur_pitch, cfg, a, pa, sc, mvx, mvy, cost);
  if (threadIdx.x == 0) {
    result[0] = mvx; result[1] = mvy; result[2] = cost;
    __threadfence_system();
    *(volatile int *)&result[3] = ticket;
  }
}

template <typename Pel>
__global__ void __launch_bounds__(128)
me_subpel_single_kernel(MeRefDev r1, MeRefDev r2, const Pel *__restrict__ cur, int cur_pitch, MeCfgDev cfg,
                        MeSubpelArgs a, int *result /* mvx, mvy, cost, ticket */, int ticket)
{
  __shared__ BlockScratch sc;
  int mvx, mvy, cost;
  subpel_block(r1, r2, cur, cur_pitch, cfg, a, sc, mvx, mvy, cost);
  if (threadIdx.x == 0) {
    result[0] = mvx; result[1] = mvy; result[2] = cost;
    __threadfence_system();                              // `result` is mapped host memory: the host spins on the ticket
    *(volatile int *)&result[3] = ticket;
  }
}

// The block search over the stored surface followed at once by the sub-pel refinement BlockMotionSearch will ask for next
// (mv-search.c:745 -> :770-774 mv <<= 2 -> :785-788 -> :823): one launch and one host wake-up instead of two.  The host
// hands the second result out when the reference's SubPelBlockMotionSearch call arrives with exactly the predicted
// arguments, and runs the plain kernel otherwise.  result: {mvx, mvy, cost, ticket, sub mvx, sub mvy, sub cost}.
template <typename Pel>
__global__ void __launch_bounds__(256)
me_ffs_block_subpel_kernel(const int32_t *__restrict__ surf_p, int setup_range, int offx, int offy, int search_range, int min_mcost,
                           int rdopt, MeRefDev r1, const Pel *__restrict__ cur, int cur_pitch, MeCfgDev cfg,
                           MeSubpelArgs a /* block, predictor, lambda[3], check0 = everything but mv == 0 */, int *result, int ticket)
{
  __shared__ BlockScratch sc;
  int mvx, mvy, cost;
  me_ffs_block_argmin(surf_p, setup_range, offx, offy, a.pmv_x, a.pmv_y, search_range, min_mcost, a.lambda[0], rdopt, sc.red, mvx, mvy, cost);
  a.mv_x = mvx * 4; a.mv_y = mvy * 4;                                   // mv <<= 2 on shorts, mv-search.c:770-774
  a.min_mcost = cfg.start_hp ? cost : ME_INT_MAX;                       // mv-search.c:785-788
  a.check0 = a.check0 && a.mv_x == 0 && a.mv_y == 0;                    // me_fullsearch.c:361
  int smx, smy, scost;
  subpel_block(r1, r1, cur, cur_pitch, cfg, a, sc, smx, smy, scost);
  if (threadIdx.x == 0) {
    result[0] = mvx; result[1] = mvy; result[2] = cost;
    result[4] = smx; result[5] = smy; result[6] = scost;
    __threadfence_system();
    *(volatile int *)&result[3] = ticket;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// One launch per PartitionMotionSearch call (me_cuda_pms_begin): CTA (item, list-ref) runs what PartitionMotionSearch
// (mv-search.c:1378-1454) does for one (blocktype, block8x8) and one (list, ref) -- for every sub-block in call order the motion
// vector predictor, the block search over the stored surface, the sub-pel refinement, and the write-back of reference index and
// vector into the motion state the next sub-block's predictor reads.

// SetMotionVectorPredictor (mv-search.c:87-281), frame macroblocks (no MBAFF): st = ref_idx / mv of the 4x4 blocks around and
// inside the macroblock ([row + 1][col + 1]); block at 4x4 position (bx4, by4), size bsx x bsy pixels.
// d_ref / d_mv: MBAFF frames only (mbaff != 0) -- neighbour D of a block in column 0 and row k, sample (-1, 4k - 1) (me_cuda.h)
struct PmsState { int8_t ref_idx[5][6]; int16_t mv[5][6][2]; int8_t d_ref[4]; int16_t d_mv[4][2]; int mbaff; };
__device__ void pms_predictor(const PmsState &st, int ref_frame, int bx4, int by4, int bsx, int bsy, int &pmx, int &pmy)
{
  const int mb_x = 4 * bx4, mb_y = 4 * by4;
  // getLuma4x4Neighbour (mb_access.c:549, :143): pixel (x, y) relative to the macroblock -> cell of the state, availability
  struct Nb { int r, c; bool ok; };
  auto nb = [&](int x, int y) {
    Nb n;
    n.c = (x < 0 ? -1 : x >> 2) + 1; n.r = (y < 0 ? -1 : y >> 2) + 1;
    if (x >= 16 && y >= 0) { n.ok = false; n.r = n.c = 0; return n; }        // to the right of the macroblock: not coded yet
    if (x >= 16) n.c = 5;                                                    // C: the macroblock above right, its block column 0
    n.ok = st.ref_idx[n.r][n.c] != ME_PMS_UNAVAILABLE;
    return n;
  };
  const Nb a = nb(mb_x - 1, mb_y), b = nb(mb_x, mb_y - 1);
  Nb d = nb(mb_x - 1, mb_y - 1);
  // MBAFF: in column 0 the neighbour D below the first row lies in the left macroblock pair at sample row mb_y - 1, which need not be
  // the 4x4 row above A's (frame / field pairs interleave their rows): its own entry, marked by row index -1
  const bool d_sep = st.mbaff && mb_x == 0 && mb_y > 0;
  if (d_sep) { d.r = -1; d.c = by4; d.ok = st.d_ref[by4] != ME_PMS_UNAVAILABLE; }
  Nb c = nb(mb_x + bsx, mb_y - 1);
  if (mb_y > 0) {                                                            // mv-search.c:111-131
    if (mb_x < 8) {
      if (mb_y == 8) { if (bsx == 16) c.ok = false; }
      else if (mb_x + bsx == 8) c.ok = false;
    } else if (mb_x + bsx == 16) c.ok = false;
  }
  if (!c.ok) c = d;
  auto ref_of_nb = [&](const Nb &n) -> int { return !n.ok ? -1 : n.r < 0 ? st.d_ref[n.c] : st.ref_idx[n.r][n.c]; };
  auto mv_of_nb = [&](const Nb &n, int hv) -> int { return !n.ok ? 0 : n.r < 0 ? st.d_mv[n.c][hv] : st.mv[n.r][n.c][hv]; };
  const int rL = ref_of_nb(a), rU = ref_of_nb(b), rUR = ref_of_nb(c);
  enum { MEDIAN, L, U, UR };
  int type = MEDIAN;
  if (rL == ref_frame && rU != ref_frame && rUR != ref_frame) type = L;
  else if (rL != ref_frame && rU == ref_frame && rUR != ref_frame) type = U;
  else if (rL != ref_frame && rU != ref_frame && rUR == ref_frame) type = UR;
  if (bsx == 8 && bsy == 16) {
    if (mb_x == 0) { if (rL == ref_frame) type = L; }
    else if (rUR == ref_frame) type = UR;
  } else if (bsx == 16 && bsy == 8) {
    if (mb_y == 0) { if (rU == ref_frame) type = U; }
    else if (rL == ref_frame) type = L;
  }
  int pm[2];
  for (int hv = 0; hv < 2; hv++) {
    const int ma = mv_of_nb(a, hv), mb = mv_of_nb(b, hv), mc = mv_of_nb(c, hv);
    int pv;
    if (type == MEDIAN) pv = !(b.ok || c.ok) ? ma : ma + mb + mc - min(ma, min(mb, mc)) - max(ma, max(mb, mc));
    else pv = type == L ? ma : type == U ? mb : mc;
    pm[hv] = (int)(int16_t)pv;                                               // pmv[] is short
  }
  pmx = pm[0]; pmy = pm[1];
}

template <typename Pel>
__global__ void __launch_bounds__(256)
me_pms_group_kernel(const __grid_constant__ MePmsParams P, const Pel *__restrict__ cur, int cur_pitch, MeCfgDev cfg /* sub-pel stage */,
                    int rdopt, MePmsRecord *result /* mapped host memory: [slot][list-ref][4] */, int *done_count, int *ticket_out, int ticket)
{
  __shared__ BlockScratch sc;
  __shared__ PmsState st;
  __shared__ int pmv_s[2];
  const int item = blockIdx.x / P.n_refs, lr = blockIdx.x - item * P.n_refs;
  const MePmsRef &R = P.refs[lr];
  const int t = P.item_type[item], bsx = me_bsx(t), bsy = me_bsy(t);
  if (threadIdx.x < 30) {
    const int r = threadIdx.x / 6, c = threadIdx.x - 6 * r;
    st.ref_idx[r][c] = P.st.ref_idx[R.list][r][c];
    st.mv[r][c][0] = P.st.mv[R.list][r][c][0]; st.mv[r][c][1] = P.st.mv[R.list][r][c][1];
  }
  if (threadIdx.x >= 32 && threadIdx.x < 36) {
    const int k = threadIdx.x - 32;
    st.d_ref[k] = P.st.d_left_ref_idx[R.list][k];
    st.d_mv[k][0] = P.st.d_left_mv[R.list][k][0]; st.d_mv[k][1] = P.st.d_left_mv[R.list][k][1];
    if (k == 0) st.mbaff = P.st.mbaff;
  }
  __syncthreads();
  int hs[4], vs[4];
  const int n = me_pms_blocks(t, P.item_blk[item], hs, vs);
  const int Pset = (2 * R.setup_range + 1) * (2 * R.setup_range + 1);
  MeCfgDev icfg = cfg; icfg.rdopt = rdopt;
  const int srange = me_block_range(icfg, R.ref, t);
  for (int k = 0; k < n; k++) {
    const int h = hs[k], v = vs[k];
    if (threadIdx.x == 0) { int px, py; pms_predictor(st, R.ref, h, v, bsx, bsy, px, py); pmv_s[0] = px; pmv_s[1] = py; }
    __syncthreads();
    const int pmx = pmv_s[0], pmy = pmv_s[1];
    const int p = me_part_index(t, 4 * v + h);
    int mvx, mvy, cost;
    me_ffs_block_argmin(R.surf + (size_t)p * Pset, R.setup_range, R.offx, R.offy, pmx, pmy, srange, ME_INT_MAX, P.lambda[0], rdopt,
                        sc.red, mvx, mvy, cost);
    MeSubpelArgs a;
    a.pic_pix_x = P.opix_x + 4 * h; a.pic_pix_y = P.opix_y + 4 * v; a.blocktype = t;
    a.pmv_x = pmx; a.pmv_y = pmy;
    a.mv_x = (int)(int16_t)(mvx * 4); a.mv_y = (int)(int16_t)(mvy * 4);       // mv <<= 2 on shorts, mv-search.c:770-774
    a.s_mv_x = a.s_mv_y = 0;
    a.search_pos2 = 9; a.search_pos4 = 9;
    a.min_mcost = cfg.start_hp ? cost : ME_INT_MAX;                          // mv-search.c:785-788
    for (int i = 0; i < 3; i++) a.lambda[i] = P.lambda[i];
    a.check0 = !rdopt && P.slice_type != ME_B_SLICE && R.ref == 0 && t == 1 && a.mv_x == 0 && a.mv_y == 0;   // me_fullsearch.c:361
    a.bi = 0;
    a.bw = MeBiW{ 0, 0, 0, 0, 0, { 0, 0 }, { 0, 0 }, { 0, 0 } };
    int smx, smy, scost;
    subpel_block(R.dev, R.dev, cur, cur_pitch, cfg, a, sc, smx, smy, scost);
    if (threadIdx.x == 0) {
      MePmsRecord &rec = result[(P.item_slot[item] * ME_PMS_MAX_REFS + lr) * 4 + k];
      rec.pmv_x = pmx; rec.pmv_y = pmy; rec.int_mvx = mvx; rec.int_mvy = mvy; rec.int_cost = cost;
      rec.mvx = smx; rec.mvy = smy; rec.cost = scost;
      // PartitionMotionSearch, mv-search.c:1441-1449: reference index and vector of the block for the predictors that follow
      for (int j = v; j < v + (bsy >> 2); j++)
        for (int i = h; i < h + (bsx >> 2); i++) { st.ref_idx[j + 1][i + 1] = (int8_t)R.ref; st.mv[j + 1][i + 1][0] = (int16_t)smx; st.mv[j + 1][i + 1][1] = (int16_t)smy; }
    }
    __syncthreads();
  }
  // the last CTA to finish hands the ticket to the host (results and ticket live in mapped host memory)
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (atomicAdd(done_count, 1) == (int)gridDim.x - 1) {
      *done_count = 0;
      __threadfence_system();
      *(volatile int *)ticket_out = ticket;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// FullPelBlockMotionSearch (me_fullsearch.c:47-155): the per-block spiral full search of SearchMode = -1.  One CTA; a thread
// scores every 256th spiral position with the whole-block distortion computeUniPred[F_PEL] (SAD / SSE: one origin clamp per
// block; SATD: one per transform tile), the first strict minimum in spiral order wins.
struct MeFullPelArgs {
  int pic_pix_x, pic_pix_y, blocktype;
  int pmv_x, pmv_y, mv_x, mv_y;
  int search_range, min_mcost, lambda;
  int check00;          // blocktype == 1 && !rdopt && slice != B && ref == 0 (me_fullsearch.c:72)
};

// whole-block distortion; thr / use_thr: the reference's early exits (after a row of SAD / SSE when the sum has reached the
// threshold, after a tile of SATD when it exceeds it), needed for the one position where they change the result (below)
template <typename Pel>
__device__ int fullpel_dist(const MeRefDev &r, const Pel *__restrict__ cur, int cur_pitch, const MeCfgDev &cfg, const MeFullPelArgs &a,
                            int x4, int y4, bool use_thr, int thr)
{
  const int metric = cfg.metric[0], bsx = me_bsx(a.blocktype), bsy = me_bsy(a.blocktype);
  const Pel *blk = cur + (size_t)a.pic_pix_y * cur_pitch + a.pic_pix_x;
  const MeBiW nobw = { 0, 0, 0, 0, 0, { 0, 0 }, { 0, 0 }, { 0, 0 } };
  int m = 0;
  if (metric == 2) {
    const bool use8 = cfg.transform8x8 && a.blocktype <= 4;
    const int n = use8 ? 8 : 4;
    for (int by = 0; by < bsy; by += n)
      for (int bx = 0; bx < bsx; bx += n) {
        const Pel *rp = me_umv_ptr_p<Pel>(r, y4 + (by << 2), x4 + (bx << 2));
        m += use8 ? tile_dist<8, Pel>(2, blk, cur_pitch, bsx, bx, by, rp, r.pitch, nullptr, 0, nobw)
                  : tile_dist<4, Pel>(2, blk, cur_pitch, bsx, bx, by, rp, r.pitch, nullptr, 0, nobw);
        if (use_thr && m > thr) return m;
      }
    return m;
  }
  const Pel *rp = me_umv_ptr_p<Pel>(r, y4, x4);
  for (int y = 0; y < bsy; y++) {
    for (int x = 0; x < bsx; x++) {
      const int d = (int)blk[(size_t)y * cur_pitch + x] - (int)rp[(size_t)y * r.pitch + x];
      m += metric ? d * d : abs(d);
    }
    if (use_thr && m >= thr) return m;
  }
  if constexpr (sizeof(Pel) == 1)
  if (cfg.chroma.on && r.chroma[0])
    for (int k = 0; k < 2; k++) {               // me_distortion.c:376-402: one component after the other, same early exit
      m += cfg.chroma.weight * me_cr_block(cfg.chroma, metric, k, a.pic_pix_x, a.pic_pix_y, bsx, bsy, r, x4, y4, nullptr, 0, 0);
      if (use_thr && m >= thr) return m;
    }
  return m;
}

template <typename Pel>
__global__ void __launch_bounds__(256)
me_fullpel_block_kernel(MeRefDev r, const Pel *__restrict__ cur, int cur_pitch, MeCfgDev cfg, MeFullPelArgs a, int *result, int ticket)
{
  __shared__ unsigned long long red[8];
  const int P = (2 * a.search_range + 1) * (2 * a.search_range + 1);
  const int pred_x = (a.pic_pix_x << 2) + a.pmv_x, pred_y = (a.pic_pix_y << 2) + a.pmv_y;
  const int center_x = a.pic_pix_x + a.mv_x, center_y = a.pic_pix_y + a.mv_y;
  unsigned long long best = ~0ull;
  for (int s = threadIdx.x; s < P; s += blockDim.x) {
    int dx, dy;
    me_spiral_xy(s, dx, dy);
    const int cand_x = (center_x + dx) << 2, cand_y = (center_y + dy) << 2;
    int mcost = me_mv_cost(a.lambda, cand_x, cand_y, pred_x, pred_y);
    // the zero-vector bonus: a quarter-pel candidate compared with integer-pel block coordinates (me_fullsearch.c:129), as is
    if (a.check00 && cand_x == a.pic_pix_x && cand_y == a.pic_pix_y) mcost -= (a.lambda * 16) >> 16;
    // Position 0 is scored against the caller's min_mcost with the early exits in force: the bonus is subtracted before the
    // distortion call and its threshold min_mcost - mcost (me_fullsearch.c:138) wraps negative for min_mcost = INT_MAX, mcost < 0
    // -- the distortion function then leaves after one row / tile and the candidate is accepted with that partial sum.
    // For every later position the running minimum is finite and a partial sum can only reject what the full sum rejects.
    const bool first = s == 0;
    if (first && mcost >= a.min_mcost) continue;
    const int thr = (int)((unsigned)a.min_mcost - (unsigned)mcost);
    mcost += fullpel_dist(r, cur, cur_pitch, cfg, a, cand_x + 4 * ME_PAD, cand_y + 4 * ME_PAD, first, thr);
    const unsigned long long key = ((unsigned long long)(unsigned)(mcost ^ 0x80000000) << 32) | (unsigned)s;   // costs can be negative
    best = key < best ? key : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other < best ? other : best;
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; i++) best = red[i] < best ? red[i] : best;
    int cost = (int)((unsigned)(best >> 32) ^ 0x80000000u), pos = (int)(unsigned)(best & 0xffffffffu);
    int dx = 0, dy = 0;
    if (best != ~0ull && cost < a.min_mcost) me_spiral_xy(pos, dx, dy); else cost = a.min_mcost;
    result[0] = a.mv_x + dx; result[1] = a.mv_y + dy; result[2] = cost;
    __threadfence_system();
    *(volatile int *)&result[3] = ticket;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// FullPelBlockMotionBiPred (me_fullsearch.c:164-332): r1 at the fixed integer vector, r2 searched over (2r+1)^2
struct MeBiFullArgs {
  int pic_pix_x, pic_pix_y, blocktype;
  int pmv_x1, pmv_y1, pmv_x2, pmv_y2;
  int mv_x, mv_y, s_mv_x, s_mv_y;
  int search_range, min_mcost, lambda;
  MeBiW bw;
};

// one bi-predictive integer-pel search; called by all threads of the CTA, returns the result to all of them
template <typename Pel>
__device__ void bipred_full_block(const MeRefDev &r1, const MeRefDev &r2, const Pel *__restrict__ cur, int cur_pitch,
                                  const MeCfgDev &cfg, const MeBiFullArgs &a, BlockScratch &sc, int &out_mvx, int &out_mvy, int &out_cost)
{
  const int P = (2 * a.search_range + 1) * (2 * a.search_range + 1);
  const int bsx = me_bsx(a.blocktype), bsy = me_bsy(a.blocktype);
  const int metric = cfg.metric[0];
  // the reference keeps these centres in `short` (me_fullsearch.c:196-199)
  const int c2x = (short)(a.pic_pix_x + a.mv_x), c2y = (short)(a.pic_pix_y + a.mv_y);
  const int c1x = (short)(a.pic_pix_x + a.s_mv_x), c1y = (short)(a.pic_pix_y + a.s_mv_y);
  const int px1 = (a.pic_pix_x << 2) + a.pmv_x1, py1 = (a.pic_pix_y << 2) + a.pmv_y1;
  const int px2 = (a.pic_pix_x << 2) + a.pmv_x2, py2 = (a.pic_pix_y << 2) + a.pmv_y2;
  const Pel *p1 = me_umv_ptr_p<Pel>(r1, (c1y << 2) + 4 * ME_PAD, (c1x << 2) + 4 * ME_PAD);
  const int fixed_cost = me_mv_cost(a.lambda, c1x << 2, c1y << 2, px1, py1);
  unsigned long long best = ~0ull;
  for (int s = threadIdx.x; s < P; s += blockDim.x) {
    int dx, dy;
    me_spiral_xy(s, dx, dy);
    int cx = (c2x + dx) << 2, cy = (c2y + dy) << 2;
    int mc = fixed_cost + me_mv_cost(a.lambda, cx, cy, px2, py2);
    const Pel *p2 = me_umv_ptr_p<Pel>(r2, cy + 4 * ME_PAD, cx + 4 * ME_PAD);
    int d = 0;
    if (metric == 2) {
      // computeBiPredSATD1 / SATD2 (me_distortion.c:824, :907): Hadamard SAD tile by tile, each tile with its own origin clamp
      const bool use8 = cfg.transform8x8 && a.blocktype <= 4;
      const int n = use8 ? 8 : 4;
      const Pel *blk = cur + (size_t)a.pic_pix_y * cur_pitch + a.pic_pix_x;
      for (int by = 0; by < bsy; by += n)
        for (int bx = 0; bx < bsx; bx += n) {
          const Pel *t1 = me_umv_ptr_p<Pel>(r1, (c1y << 2) + 4 * ME_PAD + (by << 2), (c1x << 2) + 4 * ME_PAD + (bx << 2));
          const Pel *t2 = me_umv_ptr_p<Pel>(r2, cy + 4 * ME_PAD + (by << 2), cx + 4 * ME_PAD + (bx << 2));
          d += use8 ? tile_dist<8, Pel>(2, blk, cur_pitch, bsx, bx, by, t1, r1.pitch, t2, r2.pitch, a.bw)
                    : tile_dist<4, Pel>(2, blk, cur_pitch, bsx, bx, by, t1, r1.pitch, t2, r2.pitch, a.bw);
        }
    } else
    for (int y = 0; y < bsy; y++)
      for (int x = 0; x < bsx; x++) {
        int e = (int)cur[(size_t)(a.pic_pix_y + y) * cur_pitch + a.pic_pix_x + x]
                - bipred_pel(p1[(size_t)y * r1.pitch + x], p2[(size_t)y * r2.pitch + x], a.bw);
        d += metric ? e * e : abs(e);                      // computeBiPred1[F_PEL] / computeBiPred2[F_PEL]: SAD, SSE
      }
    if constexpr (sizeof(Pel) == 1)
    if (cfg.chroma.on && metric != 2 && r1.chroma[0] && r2.chroma[0])     // computeBiPredSAD1, me_distortion.c:516-543
      for (int k = 0; k < 2; k++)
        d += cfg.chroma.weight * me_cr_block(cfg.chroma, metric, k, a.pic_pix_x, a.pic_pix_y, bsx, bsy, r1, (c1x << 2) + 4 * ME_PAD,
                                             (c1y << 2) + 4 * ME_PAD, &r2, cx + 4 * ME_PAD, cy + 4 * ME_PAD,
                                             a.bw.apply, a.bw.cw1[k], a.bw.cw2[k], a.bw.cobi[k], a.bw.denom);
    mc += d;
    unsigned long long key = ((unsigned long long)(unsigned)mc << 32) | (unsigned)s;
    best = key < best ? key : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other < best ? other : best;
  }
  __syncthreads();                      // the scratch may still be read from a previous call
  if ((threadIdx.x & 31) == 0) sc.red[threadIdx.x >> 5] = best;
  __syncthreads();
  best = sc.red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); i++) best = sc.red[i] < best ? sc.red[i] : best;
  int cost = (int)(unsigned)(best >> 32), pos = (int)(unsigned)(best & 0xffffffffu);
  int dx = 0, dy = 0;
  if (cost < a.min_mcost) me_spiral_xy(pos, dx, dy); else cost = a.min_mcost;
  out_mvx = a.mv_x + dx; out_mvy = a.mv_y + dy; out_cost = cost;
}

// The same search for the SAD metric with the data staged in shared memory and four pixels per instruction: the fixed block
// (r1), the current block and the (2r+16)^2 window of r2 around the searched vector come from the wide-border integer
// planes (pure edge replication, equal to the reference's origin clamp, refbuf.c:37, for whole blocks); per position and
// row: 5 LDS + 4 funnel shifts for the window words, byte-wise rounded average with the fixed block ((p1 + p2 + 1) >> 1,
// me_distortion.c:502) and one VABSDIFF4 per word.  win: (2 * BI_FAST_RANGE + 16 + 1) rows of BI_WIN_PITCH bytes.
constexpr int BI_FAST_RANGE = 32;
constexpr int BI_WIN_PITCH = 2 * BI_FAST_RANGE + 16 + 8;
struct BiFastScratch {
  uint32_t curw[64], p1w[64];
  uint8_t  win[(2 * BI_FAST_RANGE + 16 + 1) * BI_WIN_PITCH];
};

__device__ void bipred_full_block_sad(const MeRefDev &r1, const MeRefDev &r2, const uint8_t *__restrict__ cur, int cur_pitch,
                                      const MeBiFullArgs &a, BlockScratch &sc, BiFastScratch &fs, int &out_mvx, int &out_mvy, int &out_cost)
{
  const int R = a.search_range, P = (2 * R + 1) * (2 * R + 1), n = 2 * R + 16;
  const int c2x = (short)(a.pic_pix_x + a.mv_x), c2y = (short)(a.pic_pix_y + a.mv_y);
  const int c1x = (short)(a.pic_pix_x + a.s_mv_x), c1y = (short)(a.pic_pix_y + a.s_mv_y);
  const int px1 = (a.pic_pix_x << 2) + a.pmv_x1, py1 = (a.pic_pix_y << 2) + a.pmv_y1;
  const int px2 = (a.pic_pix_x << 2) + a.pmv_x2, py2 = (a.pic_pix_y << 2) + a.pmv_y2;
  __syncthreads();                      // the scratch may still be read from a previous call
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
    int y = i / n, x = i - y * n;
    int sy = me_clamp(c2y - R + y + r2.spad, 0, r2.sh - 1), sx = me_clamp(c2x - R + x + r2.spad, 0, r2.sw - 1);
    fs.win[y * BI_WIN_PITCH + x] = r2.search[(size_t)sy * r2.spitch + sx];
  }
  {
    // 256 threads: one pixel each of the fixed block and of the current block
    int y = threadIdx.x >> 4, x = threadIdx.x & 15;
    int sy = me_clamp(c1y + y + r1.spad, 0, r1.sh - 1), sx = me_clamp(c1x + x + r1.spad, 0, r1.sw - 1);
    reinterpret_cast<uint8_t *>(fs.p1w)[threadIdx.x] = r1.search[(size_t)sy * r1.spitch + sx];
    reinterpret_cast<uint8_t *>(fs.curw)[threadIdx.x] = cur[(size_t)(a.pic_pix_y + y) * cur_pitch + a.pic_pix_x + x];
  }
  __syncthreads();
  const int fixed_cost = me_mv_cost(a.lambda, c1x << 2, c1y << 2, px1, py1);
  unsigned long long best = ~0ull;
  for (int s = threadIdx.x; s < P; s += blockDim.x) {
    int dx, dy;
    me_spiral_xy(s, dx, dy);
    const int wx = dx + R, wy = dy + R, sh = (wx & 3) * 8;
    unsigned sad = 0;
#pragma unroll 4
    for (int y = 0; y < 16; y++) {
      const uint32_t *w = reinterpret_cast<const uint32_t *>(fs.win + (wy + y) * BI_WIN_PITCH + (wx & ~3));
      const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
      const uint32_t q[4] = { __funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh) };
#pragma unroll
      for (int c = 0; c < 4; c++) sad = __vsadu4(__vavgu4(fs.p1w[4 * y + c], q[c]), fs.curw[4 * y + c]) + sad;
    }
    const int mc = fixed_cost + me_mv_cost(a.lambda, (c2x + dx) << 2, (c2y + dy) << 2, px2, py2) + (int)sad;
    const unsigned long long key = ((unsigned long long)(unsigned)mc << 32) | (unsigned)s;
    best = key < best ? key : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other < best ? other : best;
  }
  if ((threadIdx.x & 31) == 0) sc.red[threadIdx.x >> 5] = best;
  __syncthreads();
  best = sc.red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); i++) best = sc.red[i] < best ? sc.red[i] : best;
  int cost = (int)(unsigned)(best >> 32), pos = (int)(unsigned)(best & 0xffffffffu);
  int dx = 0, dy = 0;
  if (cost < a.min_mcost) me_spiral_xy(pos, dx, dy); else cost = a.min_mcost;
  out_mvx = a.mv_x + dx; out_mvy = a.mv_y + dy; out_cost = cost;
}

template <typename Pel>
__global__ void __launch_bounds__(256)
me_bipred_full_kernel(MeRefDev r1, MeRefDev r2, const Pel *__restrict__ cur, int cur_pitch, MeCfgDev cfg,
                      MeBiFullArgs a, int *result, int ticket)
{
  __shared__ BlockScratch sc;
  int mvx, mvy, cost;
  if constexpr (sizeof(Pel) == 1) {
    __shared__ BiFastScratch fs;                          // the byte-SIMD path: 8-bit pictures
    if (cfg.metric[0] == 0 && a.search_range <= BI_FAST_RANGE && !a.bw.apply && !cfg.chroma.on) bipred_full_block_sad(r1, r2, cur, cur_pitch, a, sc, fs, mvx, mvy, cost);
    else bipred_full_block(r1, r2, cur, cur_pitch, cfg, a, sc, mvx, mvy, cost);
  } else bipred_full_block(r1, r2, cur, cur_pitch, cfg, a, sc, mvx, mvy, cost);
  if (threadIdx.x == 0) {
    result[0] = mvx; result[1] = mvy; result[2] = cost;
    __threadfence_system();
    *(volatile int *)&result[3] = ticket;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Frame batch of the bi-predictive refinement: one CTA per (macroblock, list) job runs the whole block of
// BlockMotionSearch at mv-search.c:864-1034 (block type 1, reference 0): vectors to integer pel ((mv + 2) >> 2, :884-885);
// 1 + refinements integer-pel searches in which the two lists take turns, cost carried over, range halved each time
// (:889-975); back to quarter pel (:976-982); sub-pel refinement of bimv with mv fixed (:984-1001) and, for
// BiPredMESubPel == 2, of mv with bimv fixed (:1004-1021).  Every thread keeps the same copy of the (short) vectors.
__global__ void __launch_bounds__(256)
me_bipred_batch_kernel(const uint32_t *__restrict__ jobs, int n_jobs, const MeRefDev *__restrict__ ref_table,
                       const uint8_t *__restrict__ cur, int cur_pitch, MeCfgDev cfg, int refinements, int search_range,
                       int subpel, MeBiW bw0, MeBiW bw1 /* weights of calls with list argument 0 / 1 */, me_bipred_result *__restrict__ out)
{
  __shared__ BlockScratch sc;
  __shared__ BiFastScratch fs;
  __shared__ uint32_t jw[8];
  const int job_id = blockIdx.x;
  if (job_id >= n_jobs) return;
  if (threadIdx.x < 8) jw[threadIdx.x] = jobs[(size_t)job_id * 8 + threadIdx.x];    // whole-word loads (see me_load_job)
  __syncthreads();
  const int opix_x = (int)(int16_t)(jw[0] & 0xffffu), opix_y = (int)(int16_t)(jw[0] >> 16);
  const int list = (int)(int8_t)(jw[1] & 0xffu);
  const int pred_mv[2] = { (int)(int16_t)(jw[2] & 0xffffu), (int)(int16_t)(jw[2] >> 16) };
  const int pred_mv_bi[2] = { (int)(int16_t)(jw[3] & 0xffffu), (int)(int16_t)(jw[3] >> 16) };
  short mv[2] = { (short)(jw[4] & 0xffffu), (short)(jw[4] >> 16) }, bimv[2] = { 0, 0 }, tempmv[2] = { 0, 0 };
  const int lambda[3] = { (int)jw[5], (int)jw[6], (int)jw[7] };
  const MeRefDev refs[2] = { ref_table[0 * ME_MAX_REFS + 0], ref_table[1 * ME_MAX_REFS + 0] };
  const int *pred_mv1 = pred_mv, *pred_mv2 = pred_mv_bi;
  int min_mcostbi = ME_INT_MAX, iterlist = list;
  mv[0] = (short)((mv[0] + 2) >> 2);
  mv[1] = (short)((mv[1] + 2) >> 2);
  for (int i = 0; i <= refinements; i++) {
    if (i & 1) {
      pred_mv2 = pred_mv; pred_mv1 = pred_mv_bi;
      tempmv[0] = bimv[0]; tempmv[1] = bimv[1];
      bimv[0] = mv[0]; bimv[1] = mv[1];
      iterlist = list ^ 1;
    } else {
      pred_mv1 = pred_mv; pred_mv2 = pred_mv_bi;
      if (i != 0) {
        tempmv[0] = bimv[0]; tempmv[1] = bimv[1];
        bimv[0] = mv[0]; bimv[1] = mv[1];
      } else {
        tempmv[0] = mv[0]; tempmv[1] = mv[1];
        bimv[0] = (short)((pred_mv2[0] + 2) >> 2);
        bimv[1] = (short)((pred_mv2[1] + 2) >> 2);
      }
      iterlist = list;
    }
    mv[0] = bimv[0]; mv[1] = bimv[1];
    MeBiFullArgs a;
    a.pic_pix_x = opix_x; a.pic_pix_y = opix_y; a.blocktype = 1;
    a.pmv_x1 = pred_mv1[0]; a.pmv_y1 = pred_mv1[1]; a.pmv_x2 = pred_mv2[0]; a.pmv_y2 = pred_mv2[1];
    a.mv_x = bimv[0]; a.mv_y = bimv[1]; a.s_mv_x = tempmv[0]; a.s_mv_y = tempmv[1];
    a.search_range = search_range >> i; a.min_mcost = min_mcostbi; a.lambda = lambda[0];
    int mx, my;
    a.bw = iterlist ? bw1 : bw0;
    if (cfg.metric[0] == 0 && a.search_range <= BI_FAST_RANGE && !a.bw.apply && !cfg.chroma.on)
      bipred_full_block_sad(refs[iterlist], refs[iterlist ^ 1], cur, cur_pitch, a, sc, fs, mx, my, min_mcostbi);
    else
      bipred_full_block(refs[iterlist], refs[iterlist ^ 1], cur, cur_pitch, cfg, a, sc, mx, my, min_mcostbi);
    bimv[0] = (short)mx; bimv[1] = (short)my;
    mv[0] = tempmv[0]; mv[1] = tempmv[1];
  }
  mv[0] = (short)(tempmv[0] << 2); mv[1] = (short)(tempmv[1] << 2);
  bimv[0] = (short)(bimv[0] << 2); bimv[1] = (short)(bimv[1] << 2);
  for (int pass = 0; pass < subpel && pass < 2; pass++) {
    // pass 0 refines bimv (list iterlist ^ 1) with mv fixed; pass 1 refines mv (list iterlist) with bimv fixed
    if (pass == 0 ? !cfg.start_hp : (!cfg.start_hp || !cfg.start_qp)) min_mcostbi = ME_INT_MAX;
    MeSubpelArgs a;
    a.pic_pix_x = opix_x; a.pic_pix_y = opix_y; a.blocktype = 1;
    a.pmv_x = pass == 0 ? pred_mv2[0] : pred_mv1[0]; a.pmv_y = pass == 0 ? pred_mv2[1] : pred_mv1[1];
    a.mv_x = pass == 0 ? bimv[0] : mv[0]; a.mv_y = pass == 0 ? bimv[1] : mv[1];
    a.s_mv_x = pass == 0 ? mv[0] : bimv[0]; a.s_mv_y = pass == 0 ? mv[1] : bimv[1];
    a.search_pos2 = 9; a.search_pos4 = 9; a.min_mcost = min_mcostbi;
    a.lambda[0] = lambda[0]; a.lambda[1] = lambda[1]; a.lambda[2] = lambda[2];
    a.check0 = 0; a.bi = 1;
    const int l1 = pass == 0 ? iterlist : iterlist ^ 1;            // the `list` argument of SubPelBlockSearchBiPred: r1 = its reference 0
    a.bw = l1 ? bw1 : bw0;
    int mx, my;
    subpel_block(refs[l1], refs[l1 ^ 1], cur, cur_pitch, cfg, a, sc, mx, my, min_mcostbi);
    if (pass == 0) { bimv[0] = (short)mx; bimv[1] = (short)my; } else { mv[0] = (short)mx; mv[1] = (short)my; }
  }
  if (threadIdx.x == 0) {
    uint32_t *o = reinterpret_cast<uint32_t *>(&out[job_id]);
    o[0] = ((unsigned)(unsigned short)mv[0]) | ((unsigned)(unsigned short)mv[1] << 16);
    o[1] = ((unsigned)(unsigned short)bimv[0]) | ((unsigned)(unsigned short)bimv[1] << 16);
    o[2] = (uint32_t)min_mcostbi;
    o[3] = (uint32_t)iterlist;
  }
}

extern "C" void me_launch_subpel_batch(const MeJob *jobs, int n_jobs, const MeRefDev *ref_table, const uint8_t *cur,
                                       int cur_pitch, const MeCfgDev *cfg, const MeResultInt *in_int, MeResult *out,
                                       cudaStream_t st)
{
  if (cfg->pel16)
    me_subpel_batch_kernel<uint16_t><<<n_jobs, 256, 0, st>>>(jobs, n_jobs, ref_table, reinterpret_cast<const uint16_t *>(cur), cur_pitch, *cfg, in_int, out);
  else
    me_subpel_batch_kernel<uint8_t><<<n_jobs, 256, 0, st>>>(jobs, n_jobs, ref_table, cur, cur_pitch, *cfg, in_int, out);
}

extern "C" void me_launch_bipred_batch(const me_bipred_job *jobs, int n_jobs, const MeRefDev *ref_table, const uint8_t *cur,
                                       int cur_pitch, const MeCfgDev *cfg, int refinements, int search_range, int subpel,
                                       const MeBiW *bw /* [2] */, me_bipred_result *out, cudaStream_t st)
{
  static_assert(sizeof(me_bipred_job) == 32 && sizeof(me_bipred_result) == 16, "ABI struct sizes");
  me_bipred_batch_kernel<<<n_jobs, 256, 0, st>>>(reinterpret_cast<const uint32_t *>(jobs), n_jobs, ref_table, cur, cur_pitch, *cfg,
                                                 refinements, search_range, subpel, bw[0], bw[1], out);
}

extern "C" void me_launch_ffs_block_subpel(const int32_t *surf_p, int setup_range, int offx, int offy, int search_range, int min_mcost,
                                           int rdopt, const MeRefDev *r1, const uint8_t *cur, int cur_pitch, const MeCfgDev *cfg,
                                           const MeSubpelArgs *a, int *result, int ticket, cudaStream_t st)
{
  if (cfg->pel16)
    me_ffs_block_subpel_kernel<uint16_t><<<1, 256, 0, st>>>(surf_p, setup_range, offx, offy, search_range, min_mcost, rdopt, *r1,
                                                            reinterpret_cast<const uint16_t *>(cur), cur_pitch, *cfg, *a, result, ticket);
  else
    me_ffs_block_subpel_kernel<uint8_t><<<1, 256, 0, st>>>(surf_p, setup_range, offx, offy, search_range, min_mcost, rdopt, *r1, cur, cur_pitch,
                                                           *cfg, *a, result, ticket);
}

extern "C" void me_launch_pms_group(const MePmsParams *p, const uint8_t *cur, int cur_pitch, const MeCfgDev *sub_cfg, int rdopt,
                                    MePmsRecord *result, int *done_count, int *ticket_out, int ticket, cudaStream_t st)
{
  if (sub_cfg->pel16)
    me_pms_group_kernel<uint16_t><<<p->n_items * p->n_refs, 256, 0, st>>>(*p, reinterpret_cast<const uint16_t *>(cur), cur_pitch, *sub_cfg, rdopt,
                                                                          result, done_count, ticket_out, ticket);
  else
    me_pms_group_kernel<uint8_t><<<p->n_items * p->n_refs, 256, 0, st>>>(*p, cur, cur_pitch, *sub_cfg, rdopt, result, done_count, ticket_out, ticket);
}

extern "C" void me_launch_subpel_single(const MeRefDev *r1, const MeRefDev *r2, const uint8_t *cur, int cur_pitch,
                                        const MeCfgDev *cfg, const MeSubpelArgs *a, int *result, int ticket, cudaStream_t st)
{
  if (cfg->pel16)
    me_subpel_single_kernel<uint16_t><<<1, 128, 0, st>>>(*r1, r2 ? *r2 : *r1, reinterpret_cast<const uint16_t *>(cur), cur_pitch, *cfg, *a, result, ticket);
  else
    me_subpel_single_kernel<uint8_t><<<1, 128, 0, st>>>(*r1, r2 ? *r2 : *r1, cur, cur_pitch, *cfg, *a, result, ticket);
}

extern "C" void me_launch_subpel_pattern(const MeRefDev *r1, const MeRefDev *r2, const uint8_t *cur, int cur_pitch,
                                         const MeCfgDev *cfg, const MeSubpelArgs *a, const MePatArgs *pa, int *result, int ticket, cudaStream_t st)
{
  if (cfg->pel16)
    me_subpel_pattern_kernel<uint16_t><<<1, 256, 0, st>>>(*r1, r2 ? *r2 : *r1, reinterpret_cast<const uint16_t *>(cur), cur_pitch, *cfg, *a, *pa, result, ticket);
  else
    me_subpel_pattern_kernel<uint8_t><<<1, 256, 0, st>>>(*r1, r2 ? *r2 : *r1, cur, cur_pitch, *cfg, *a, *pa, result, ticket);
}

extern "C" void me_launch_fullpel_block(const MeRefDev *r, const uint8_t *cur, int cur_pitch, const MeCfgDev *cfg,
                                        const MeFullPelArgs *a, int *result, int ticket, cudaStream_t st)
{
  if (cfg->pel16)
    me_fullpel_block_kernel<uint16_t><<<1, 256, 0, st>>>(*r, reinterpret_cast<const uint16_t *>(cur), cur_pitch, *cfg, *a, result, ticket);
  else
    me_fullpel_block_kernel<uint8_t><<<1, 256, 0, st>>>(*r, cur, cur_pitch, *cfg, *a, result, ticket);
}

extern "C" void me_launch_bipred_full(const MeRefDev *r1, const MeRefDev *r2, const uint8_t *cur, int cur_pitch,
                                      const MeCfgDev *cfg, const MeBiFullArgs *a, int *result, int ticket, cudaStream_t st)
{
  if (cfg->pel16)
    me_bipred_full_kernel<uint16_t><<<1, 256, 0, st>>>(*r1, *r2, reinterpret_cast<const uint16_t *>(cur), cur_pitch, *cfg, *a, result, ticket);
  else
    me_bipred_full_kernel<uint8_t><<<1, 256, 0, st>>>(*r1, *r2, cur, cur_pitch, *cfg, *a, result, ticket);
}
