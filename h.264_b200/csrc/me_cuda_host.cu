// me_cuda_host.cu -- host side of libme_cuda.so: the C ABI of include/me_cuda.h.
//
// Owns the device mirrors of the encoder's pictures (current picture = pCurImg, reference pictures = the imgY_sub
// plane sets of StorablePicture), the per-macroblock fast-full-search state that lives in the static arrays of
// the reference's me_fullfast.c:41-46 (search_setup_done / search_center_x/y / BlockSAD), and the launch sequences.
// There is no CPU fallback in here: every entry point needs a working CUDA device.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "me_common.cuh"

// ---- kernel launchers (planes.cu, ffs_basic.cu, ffs_tuned.cu, subpel.cu) ----
struct MeBiW { int apply, w1, w2, obi, denom; int cw1[2], cw2[2], cobi[2]; int maxv; };   // weighted bi-prediction of one `list` argument (subpel.cu)
struct MeSubpelArgs {
  int pic_pix_x, pic_pix_y, blocktype;
  int pmv_x, pmv_y, mv_x, mv_y, s_mv_x, s_mv_y;
  int search_pos2, search_pos4, min_mcost;
  int lambda[3];
  int check0;
  int bi;
  MeBiW bw;
};
struct MePatArgs { int pattern, subthres, uplayer_x, uplayer_y, thr1, thr3, pmv2_x, pmv2_y; };   // subpel.cu
struct MeBiFullArgs {
  int pic_pix_x, pic_pix_y, blocktype;
  int pmv_x1, pmv_y1, pmv_x2, pmv_y2;
  int mv_x, mv_y, s_mv_x, s_mv_y;
  int search_range, min_mcost, lambda;
  MeBiW bw;
};
struct MeFullPelArgs {
  int pic_pix_x, pic_pix_y, blocktype;
  int pmv_x, pmv_y, mv_x, mv_y;
  int search_range, min_mcost, lambda;
  int check00;
};
extern "C" {
void me_launch_planes(const uint8_t *src, int src_stride, int w, int h, uint8_t *planes, int pitch, size_t plane_stride,
                      uint8_t *search, int spad, int spitch, cudaStream_t st);
void me_launch_planes16(const uint16_t *src, int src_stride, int w, int h, uint16_t *planes, int pitch, size_t plane_stride, int max_pel, cudaStream_t st);
void me_launch_ffs_surfaces16(const MeRefDev *ref, const uint16_t *cur, int cur_pitch, int opix_x, int opix_y, int cx, int cy, int R, int sse,
                              int32_t *surf, cudaStream_t st);
void me_launch_widen_planes(const uint8_t *planes, int pitch, size_t plane_stride, int wp, int hp, uint16_t *dst, cudaStream_t st);
void me_launch_check_jobs(const MeJob *jobs, int n_jobs, int n_refs0, int n_refs1, int width, int height, int *bad, cudaStream_t st);
void me_launch_luma_prediction(const void *blocks, int n, int bsx, int bsy, const MeRefDev *table, uint8_t *pred, cudaStream_t st);
void me_launch_pred_cost8x8(const void *blocks, int n_regions, int metric, const MeRefDev *table, const uint8_t *cur, int cur_pitch,
                            int pel16, int max_pel, int32_t *cost4, int32_t *cost8, cudaStream_t st);
void me_launch_widen_region(const uint8_t *plane, int pitch, int x0, int y0, int w, int h, uint16_t *dst, cudaStream_t st);
void me_launch_deblock(uint8_t *Y, int ypitch, uint8_t *U, uint8_t *V, int cpitch, int mbw, int mbh, const void *mbs, int *progress, cudaStream_t st);
void me_launch_weight_planes16(const uint16_t *src, uint16_t *dst, size_t n, int weight, int offset, int log_denom, int max_pel, cudaStream_t st);
void me_launch_weight_planes(const uint8_t *src, uint8_t *dst, size_t bytes, int weight, int offset, int log_denom, cudaStream_t st);
void me_launch_ffs_batch_basic(const MeJob *jobs, int n_jobs, const MeRefDev *ref_table, const uint8_t *cur, int cur_pitch,
                               const MeCfgDev *cfg, MeResultInt *out, const int *job_list, cudaStream_t st);
int  me_launch_ffs_batch_tuned(const MeJob *jobs, int n_jobs, const MeRefDev *ref_table, const void *win_maps, int n_refs0, int n_refs1,
                               const void *cur_map, const uint8_t *cur_pic, int cur_pitch, const MeCfgDev *cfg, MeResultInt *out,
                               int *deferred, int *ticket, cudaStream_t st);   // kernels launched, 0 = not applicable
int  me_make_tensor_map_2d(void *out, const void *base, int w, int h, int pitch, int bw, int bh);
int  me_make_planes_tensor_map(void *out, const void *planes, int pitch, int hp, size_t plane_stride);
int  me_launch_subpel_batch_tuned(const MeJob *jobs, int n_jobs, const MeRefDev *ref_table, const void *plane_maps, int n_refs0, int n_refs1,
                                  const uint8_t *cur, int cur_pitch, const MeCfgDev *cfg, const MeResultInt *in_int, MeResult *out,
                                  int *ticket, unsigned long long *stats, cudaStream_t st);   // 1 = launched, 0 = not applicable
int  me_ffs_tuned_window(int search_range);
void me_launch_ffs_surfaces(const MeRefDev *ref, const uint8_t *cur, int cur_pitch, int opix_x, int opix_y, int cx, int cy,
                            int R, int sse, const MeChromaDev *chroma, int32_t *surf, cudaStream_t st);
void me_launch_ffs_block(const int32_t *surf_p, int setup_range, int offx, int offy, int pmv_x, int pmv_y, int search_range,
                         int min_mcost, int lambda, int rdopt, int *result, int ticket, cudaStream_t st);
void me_launch_subpel_batch(const MeJob *jobs, int n_jobs, const MeRefDev *ref_table, const uint8_t *cur, int cur_pitch,
                            const MeCfgDev *cfg, const MeResultInt *in_int, MeResult *out, cudaStream_t st);
void me_launch_ffs_block_subpel(const int32_t *surf_p, int setup_range, int offx, int offy, int search_range, int min_mcost, int rdopt,
                                const MeRefDev *r1, const uint8_t *cur, int cur_pitch, const MeCfgDev *cfg, const MeSubpelArgs *a,
                                int *result, int ticket, cudaStream_t st);
void me_launch_subpel_single(const MeRefDev *r1, const MeRefDev *r2, const uint8_t *cur, int cur_pitch, const MeCfgDev *cfg,
                             const MeSubpelArgs *a, int *result, int ticket, cudaStream_t st);
void me_launch_subpel_pattern(const MeRefDev *r1, const MeRefDev *r2, const uint8_t *cur, int cur_pitch, const MeCfgDev *cfg,
                              const MeSubpelArgs *a, const MePatArgs *pa, int *result, int ticket, cudaStream_t st);
void me_launch_pms_group(const MePmsParams *p, const uint8_t *cur, int cur_pitch, const MeCfgDev *sub_cfg, int rdopt,
                         MePmsRecord *result, int *done_count, int *ticket_out, int ticket, cudaStream_t st);
void me_launch_bipred_batch(const me_bipred_job *jobs, int n_jobs, const MeRefDev *ref_table, const uint8_t *cur, int cur_pitch,
                            const MeCfgDev *cfg, int refinements, int search_range, int subpel, const MeBiW *bw, me_bipred_result *out,
                            cudaStream_t st);
void me_launch_bipred_full(const MeRefDev *r1, const MeRefDev *r2, const uint8_t *cur, int cur_pitch, const MeCfgDev *cfg,
                           const MeBiFullArgs *a, int *result, int ticket, cudaStream_t st);
void me_launch_fullpel_block(const MeRefDev *r, const uint8_t *cur, int cur_pitch, const MeCfgDev *cfg, const MeFullPelArgs *a,
                             int *result, int ticket, cudaStream_t st);
}

namespace {

constexpr int ME_PIPE_CHUNKS = 8;       // most chunks a batch is cut into (events are created for this many)

struct TensorMap { alignas(64) unsigned char bytes[128]; };   // CUtensorMap

struct RefPic {
  bool used = false;
  uint8_t *chroma = nullptr;    // ChromaMEEnable: both reconstructed chroma planes, one allocation (U then V)
  me_ref_handle chroma444[2] = {0, 0};   // 4:4:4: the pictures made from the U / V planes (their plane sets are the chroma sub-images)
  bool has_chroma() const { return chroma != nullptr || chroma444[0] != 0; }
  unsigned long long gen = 0;   // changes whenever the planes are (re)generated: weighted copies made from them go stale
  cudaStream_t written_on = nullptr;   // stream of the last plane generation (a caller's stream in the *_device calls)
  uint8_t *planes = nullptr, *search = nullptr;
  MeRefDev dev{};
  bool has_map = false;
  TensorMap win_map{};          // search plane, box = one search window of the tuned kernel
  bool has_pmap = false;
  TensorMap planes_map{};       // the 16 planes as a 3-D tensor, box = one sub-pel region of the tuned kernel
};

// Weighted-prediction ME: the weights of one reference-list entry (wp_weight / wp_offset [list][ref][0], luma_log_weight_denom)
struct WeightSet {
  int apply = 0, weight = 32, offset = 0, denom = 5;
  bool identity() const { return !apply || (weight == (1 << denom) && offset == 0); }
  bool operator==(const WeightSet &o) const { return apply == o.apply && weight == o.weight && offset == o.offset && denom == o.denom; }
};
// ... and the weighted copy of that entry's planes the uni-predictive searches read instead (planes.cu: me_weight_planes_kernel)
struct Twin {
  uint8_t *planes = nullptr, *search = nullptr;
  size_t planes_bytes = 0, search_bytes = 0;
  MeRefDev dev{};
  bool has_map = false, has_pmap = false;
  TensorMap win_map{}, planes_map{};
  me_ref_handle src = 0; unsigned long long src_gen = 0; WeightSet w; bool valid = false;
};

struct State {
  bool inited = false;
  me_cuda_params p{};
  MeCfgDev cfg{};
  int device = 0;
  cudaStream_t stream = nullptr;
  std::vector<RefPic> refs;
  unsigned long long gen_counter = 0;
  WeightSet wp[2][ME_MAX_REFS];
  Twin twin[2][ME_MAX_REFS];
  MeBiW bw[2] = {};   // weighted bi-prediction per `list` argument of the bi-predictive searches
  // current picture
  uint8_t *d_cur = nullptr;
  int cur_pitch = 0;
  int pel_bytes = 1;            // 2: pictures of more than 8 bits per sample (cfg.pel16): 16-bit planes and current picture
  int cur_h = 0;                // rows of the picture last set (the initialised height, or half of it: a field picture)
  // what the drop-in calls read as the current picture: the buffer itself, or one field of it (me_cuda_select_current_view)
  int view = 0; const uint8_t *cur_view = nullptr; int cur_view_pitch = 0, cur_view_h = 0;
  int chroma_base_pitch = 0;    // row pitch of d_cur_c (cfg.chroma.cur_pitch is that of the view)
  int cur_ch = 0;               // rows of each original chroma plane last uploaded (V follows U after that many rows)
  uint8_t *d_cur_c = nullptr;   // ChromaMEEnable: original chroma planes (U then V), rows of cfg.chroma.cur_pitch bytes
  int chroma_w = 0, chroma_h = 0;
  // staging
  uint8_t *h_pin = nullptr; size_t h_pin_bytes = 0;
  uint16_t *d_wide = nullptr; uint16_t *h_wide = nullptr; size_t wide_elems = 0;
  // reference lists
  MeRefDev h_table[2 * ME_MAX_REFS];
  MeRefDev *d_table = nullptr;
  TensorMap h_maps[2 * ME_MAX_REFS];
  TensorMap h_pmaps[2 * ME_MAX_REFS];
  bool pmaps_ok = false;
  bool maps_ok = false;         // every listed reference has a window tensor map
  TensorMap cur_map{}; bool has_cur_map = false;   // current picture, box 16 x 16
  unsigned long long *d_subpel_stats = nullptr;      // {partitions the sub-pel kernel served from its staged region, all partitions}, while enabled
  int *d_deferred = nullptr; int deferred_cap = 0;  // {job ticket counters of the two tuned kernels; jobs the search kernel hands back: count, ids...}
  cudaEvent_t ev_order = nullptr;                   // orders a caller's stream (the *_device calls) against the library stream
  cudaStream_t s_in = nullptr, s_out = nullptr;     // host <-> device pipeline of me_cuda_frame_search
  cudaEvent_t ev_in[ME_PIPE_CHUNKS] = {}, ev_done[ME_PIPE_CHUNKS] = {};
  me_ref_handle list[2][ME_MAX_REFS];
  int list_n[2] = {0, 0};
  bool table_dirty = true;
  int slice_type = ME_P_SLICE;
  // per-macroblock fast full search state
  int opix_x = 0, opix_y = 0;
  int setup_done[2][ME_MAX_REFS];
  int centre_x[2][ME_MAX_REFS], centre_y[2][ME_MAX_REFS], setup_range[2][ME_MAX_REFS];
  int32_t *d_surf[2][ME_MAX_REFS];
  int *d_result = nullptr, *h_result = nullptr;   // one mapped pinned buffer: the drop-in kernels write {mvx, mvy, cost, ticket}
  int ticket = 0;                                 // straight into host memory and the host spins on the ticket (no copy, no sync)
  // The reference follows every integer-pel block search with the sub-pel refinement of the same block (mv-search.c:745 ->
  // :823).  Once a sub-pel call has shown the lambda factors and position counts in use, the block search kernel also
  // computes that refinement; the result is handed out when the sub-pel call arrives with exactly the predicted arguments.
  struct { bool seen = false; int lambda[3] = {0, 0, 0}, pos2 = 9, pos4 = 9; } sub_params;
  struct { bool valid = false; int list, ref, ppx, ppy, bt, pmvx, pmvy, mv_in_x, mv_in_y, min_in, lambda[3], pos2, pos4;
           int mvx, mvy, cost; } spec;
  // One launch per PartitionMotionSearch (me_cuda_pms_begin): result records and the ticket of the last group launch live in one
  // mapped pinned buffer (h_pms / d_pms); an item = one (blocktype, block8x8) call whose results wait for the drop-in calls
  struct Pms {
    MePmsRecord *h_rec = nullptr, *d_rec = nullptr;     // [ME_PMS_SLOTS][ME_PMS_MAX_REFS][4]
    volatile int *h_ticket = nullptr; int *d_ticket = nullptr;
    int *d_done = nullptr;
    int ticket = 0;                                     // of the last launch; launches finish in order, h_ticket follows
    struct Item { bool valid = false; int blocktype = 0, block8x8 = 0, ticket = 0, lambda[3] = {0, 0, 0}; unsigned age = 0; } item[ME_PMS_SLOTS];
    unsigned clock = 0;
    int n_lr = 0, lr_list[ME_PMS_MAX_REFS], lr_ref[ME_PMS_MAX_REFS];
    bool ffs_served[ME_PMS_SLOTS][ME_PMS_MAX_REFS][4];
    long long launches = 0, served = 0, missed = 0;
  } pms;
  // batch scratch
  MeJob *d_jobs = nullptr; MeResult *d_out = nullptr; MeResultInt *d_out_int = nullptr; int batch_cap = 0;
  MeResultInt *d_int_scratch = nullptr; int int_scratch_cap = 0;
  me_bipred_job *d_bi_jobs = nullptr; me_bipred_result *d_bi_out = nullptr; int bi_cap = 0;
  MeRefDev *d_bi_table = nullptr;
  // motion-compensated prediction / mode-decision costs (mc.cu): the listed pictures' own planes, and a staging area that is
  // mapped pinned memory up to MC_MAPPED bytes (one launch + one wait per call, no copies) and device memory beyond
  MeRefDev *d_mc_table = nullptr; MeRefDev h_mc_table[2 * ME_MAX_REFS]; bool mc_table_valid = false;
  uint8_t *h_mc = nullptr, *d_mc_mapped = nullptr;        // mapped pinned
  uint8_t *d_mc = nullptr; size_t d_mc_bytes = 0;         // device
  // deblocking (deblock.cu): the picture being filtered (luma, then U, V), the macroblock records and the row progress counters
  uint8_t *d_db = nullptr; size_t d_db_bytes = 0; int db_w = 0, db_h = 0; bool db_valid = false;
  long long launches = 0;
  long long spec_hits = 0;                        // sub-pel calls answered from the fused block-search kernel
  // optional per-kernel-class timing (bench.py): event pairs around every launch of a class
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[3];   // 0 planes, 1 integer-pel search, 2 sub-pel
};

State g;
char g_err[512] = "";
me_cuda_error_fn g_handler = nullptr;

int fail(int code, const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

// drop-in search calls return a cost, so they report through the handler
void fatal(const char *what)
{
  if (g_handler) { g_handler(g_err[0] ? g_err : what, 500); return; }
  fprintf(stderr, "libme_cuda: fatal: %s\n", g_err[0] ? g_err : what);
  abort();
}

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(2, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

int align_up(int v, int a) { return (v + a - 1) / a * a; }

// The *_device entry points take a caller's stream.  Everything else -- plane sets, the current picture, the reference
// table, the scratch arrays of the batch -- is produced and consumed on the library stream, so work on a caller's stream is
// bracketed: it starts after everything queued on the library stream so far (enter) and the library stream continues after
// it (leave).  Program order of the calls is then the order on the device, whichever streams they name.
int enter_stream(cudaStream_t st)
{
  if (st == g.stream) return 0;
  CK(cudaEventRecord(g.ev_order, g.stream));
  CK(cudaStreamWaitEvent(st, g.ev_order, 0));
  return 0;
}
int leave_stream(cudaStream_t st)
{
  if (st == g.stream) return 0;
  CK(cudaEventRecord(g.ev_order, st));
  CK(cudaStreamWaitEvent(g.stream, g.ev_order, 0));
  return 0;
}

struct ScopedTimer {
  int which; cudaStream_t st; cudaEvent_t e1 = nullptr; bool on;
  ScopedTimer(int w, cudaStream_t s) : which(w), st(s), on(g.timing) {
    if (!on) return;
    cudaEvent_t e0 = nullptr;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { if (e0) cudaEventDestroy(e0); e1 = nullptr; on = false; return; }
    cudaEventRecord(e0, st);
    g.ev[which].push_back({e0, e1});
  }
  ~ScopedTimer() { if (on) cudaEventRecord(e1, st); }
};

RefPic *ref_of(me_ref_handle h)
{
  if (h <= 0 || (size_t)h > g.refs.size() || !g.refs[h - 1].used) return nullptr;
  return &g.refs[h - 1];
}

int ensure_pin(size_t bytes)
{
  if (g.h_pin_bytes >= bytes) return 0;
  if (g.h_pin) cudaFreeHost(g.h_pin);
  g.h_pin = nullptr; g.h_pin_bytes = 0;
  CK(cudaMallocHost((void **)&g.h_pin, bytes));
  g.h_pin_bytes = bytes;
  return 0;
}

// The planes a uni-predictive search of reference-list entry (list, idx) reads: the picture's own, or -- under weighted-
// prediction ME with non-trivial weights -- their weighted copy, (re)built here when the picture or the weights changed.
int search_planes(int list, int idx, RefPic *r, const MeRefDev **dev, const TensorMap **win_map, bool *has_map,
                  const TensorMap **planes_map, bool *has_pmap)
{
  const WeightSet &w = g.wp[list][idx];
  if (w.identity()) {
    *dev = &r->dev;
    if (win_map) { *win_map = &r->win_map; *has_map = r->has_map; *planes_map = &r->planes_map; *has_pmap = r->has_pmap; }
    return 0;
  }
  Twin &t = g.twin[list][idx];
  const me_ref_handle h = g.list[list][idx];
  if (!(t.valid && t.src == h && t.src_gen == r->gen && t.w == w)) {
    const MeRefDev &d = r->dev;
    const size_t pb = d.plane_stride * 16 * g.pel_bytes, sb = (size_t)d.spitch * d.sh;      // 16-bit pictures have no search plane (sb = 0)
    t.valid = false;
    if (t.planes_bytes != pb || t.search_bytes != sb) {
      CK(cudaStreamSynchronize(g.stream));
      if (t.planes) cudaFree(t.planes);
      if (t.search) cudaFree(t.search);
      t.planes = t.search = nullptr; t.planes_bytes = t.search_bytes = 0;
      CK(cudaMalloc((void **)&t.planes, pb));
      if (sb) CK(cudaMalloc((void **)&t.search, sb));
      t.planes_bytes = pb; t.search_bytes = sb;
      t.has_map = false; t.has_pmap = false;
      if (g.pel_bytes == 1) {
        if (int n = me_ffs_tuned_window(g.cfg.search_range))
          t.has_map = me_make_tensor_map_2d(t.win_map.bytes, t.search, d.sw, d.sh, d.spitch, n + 16, n) == 0;
        t.has_pmap = me_make_planes_tensor_map(t.planes_map.bytes, t.planes, d.pitch, d.hp, d.plane_stride) == 0;
      }
    }
    // everything but the luma planes (geometry, chroma planes) is the picture's own: taken afresh, the slot may have served another picture
    t.dev = d; t.dev.planes = t.planes; t.dev.search = t.search;
    // the planes may have been generated on a caller's stream (me_cuda_ref_create_device / me_cuda_ref_update_device)
    if (r->written_on && r->written_on != g.stream) CK(cudaStreamSynchronize(r->written_on));
    if (g.pel_bytes == 2) {
      ScopedTimer tm(0, g.stream);
      me_launch_weight_planes16((const uint16_t *)r->planes, (uint16_t *)t.planes, d.plane_stride * 16, w.weight, w.offset, w.denom, g.cfg.max_pel, g.stream);
    } else
    { ScopedTimer tm(0, g.stream);
      me_launch_weight_planes(r->planes, t.planes, pb, w.weight, w.offset, w.denom, g.stream);
      me_launch_weight_planes(r->search, t.search, sb, w.weight, w.offset, w.denom, g.stream); }
    g.launches += 2;
    CK(cudaGetLastError());
    t.src = h; t.src_gen = r->gen; t.w = w; t.valid = true;
  }
  *dev = &t.dev;
  if (win_map) { *win_map = &t.win_map; *has_map = t.has_map; *planes_map = &t.planes_map; *has_pmap = t.has_pmap; }
  return 0;
}

int upload_table()
{
  if (!g.table_dirty) return 0;
  memset(g.h_table, 0, sizeof g.h_table);
  memset(g.h_maps, 0, sizeof g.h_maps);
  memset(g.h_pmaps, 0, sizeof g.h_pmaps);
  g.maps_ok = true; g.pmaps_ok = true;
  for (int l = 0; l < 2; l++)
    for (int i = 0; i < g.list_n[l]; i++) {
      RefPic *r = ref_of(g.list[l][i]);
      if (!r) return fail(3, "reference list %d entry %d holds an invalid handle", l, i);
      const MeRefDev *dev; const TensorMap *wm, *pm; bool hm, hp;
      if (g.p.chroma_me_enable && !r->has_chroma())
        return fail(3, "reference list %d entry %d has no chroma planes (ChromaMEEnable: call me_cuda_ref_set_chroma after creating the picture)", l, i);
      if (int rc = search_planes(l, i, r, &dev, &wm, &hm, &pm, &hp)) return rc;
      g.h_table[l * ME_MAX_REFS + i] = *dev;
      g.h_maps[l * ME_MAX_REFS + i] = *wm;
      g.maps_ok = g.maps_ok && hm;
      g.h_pmaps[l * ME_MAX_REFS + i] = *pm;
      g.pmaps_ok = g.pmaps_ok && hp;
    }
  CK(cudaMemcpyAsync(g.d_table, g.h_table, sizeof g.h_table, cudaMemcpyHostToDevice, g.stream));
  // h_table is pageable: the copy above is staged synchronously by the runtime, safe to reuse
  g.table_dirty = false;
  return 0;
}

int ensure_batch(int n)
{
  if (n <= g.batch_cap) return 0;
  if (g.d_jobs) cudaFree(g.d_jobs);
  if (g.d_out) cudaFree(g.d_out);
  if (g.d_out_int) cudaFree(g.d_out_int);
  g.d_jobs = nullptr; g.d_out = nullptr; g.d_out_int = nullptr; g.batch_cap = 0;
  CK(cudaMalloc((void **)&g.d_jobs, sizeof(MeJob) * (size_t)n));
  CK(cudaMalloc((void **)&g.d_out, sizeof(MeResult) * (size_t)n));
  CK(cudaMalloc((void **)&g.d_out_int, sizeof(MeResultInt) * (size_t)n));
  g.batch_cap = n;
  return 0;
}

int setup_range_of(int ref)
{
  return (g.cfg.full_search == 2 || ref == 0) ? g.cfg.search_range : g.cfg.search_range / 2;
}

int create_ref_common(const uint8_t *d_luma, int stride, int w, int h, cudaStream_t st, me_ref_handle *out)
{
  // find a free slot
  size_t slot = 0;
  for (; slot < g.refs.size(); slot++) if (!g.refs[slot].used) break;
  if (slot == g.refs.size()) g.refs.emplace_back();
  RefPic &r = g.refs[slot];
  MeRefDev &d = r.dev;
  d.w = w; d.h = h; d.wp = w + 2 * ME_PAD; d.hp = h + 2 * ME_PAD;
  d.pitch = align_up(d.wp, 128);
  d.plane_stride = (size_t)d.pitch * d.hp;
  d.spad = (g.cfg.search_range + 32 + 1) & ~1;      // even: the search plane is written four bytes at a time
  d.sw = w + 2 * d.spad; d.sh = h + 2 * d.spad;
  d.spitch = align_up(d.sw, 128);
  if (g.pel_bytes == 2) {
    // 16-bit samples: pitch / plane_stride count samples; no search-plane copy, no tensor maps (the throughput kernels are 8-bit)
    CK(cudaMalloc((void **)&r.planes, d.plane_stride * 16 * sizeof(uint16_t)));
    r.search = nullptr; d.planes = r.planes; d.search = nullptr; d.spad = d.sw = d.sh = d.spitch = 0;
    r.has_map = r.has_pmap = false;
    { ScopedTimer t(0, st); me_launch_planes16((const uint16_t *)d_luma, stride, w, h, (uint16_t *)r.planes, d.pitch, d.plane_stride, g.cfg.max_pel, st); }
    g.launches += 1;
    CK(cudaGetLastError());
    r.used = true; r.written_on = st; r.gen = ++g.gen_counter;
    *out = (me_ref_handle)(slot + 1);
    return 0;
  }
  CK(cudaMalloc((void **)&r.planes, d.plane_stride * 16));
  CK(cudaMalloc((void **)&r.search, (size_t)d.spitch * d.sh));
  d.planes = r.planes; d.search = r.search;
  if (int n = me_ffs_tuned_window(g.cfg.search_range))
    r.has_map = me_make_tensor_map_2d(r.win_map.bytes, r.search, d.sw, d.sh, d.spitch, n + 16, n) == 0;
  r.has_pmap = me_make_planes_tensor_map(r.planes_map.bytes, r.planes, d.pitch, d.hp, d.plane_stride) == 0;
  { ScopedTimer t(0, st); me_launch_planes(d_luma, stride, w, h, r.planes, d.pitch, d.plane_stride, r.search, d.spad, d.spitch, st); }
  g.launches += 2;
  CK(cudaGetLastError());
  r.used = true;
  r.written_on = st;
  r.gen = ++g.gen_counter;
  *out = (me_ref_handle)(slot + 1);
  return 0;
}

// regenerate the planes of an existing picture slot in place (a DPB frame store being reused)
int update_ref_common(me_ref_handle h, const uint8_t *d_luma, int stride, cudaStream_t st)
{
  RefPic *r = ref_of(h);
  if (!r) return fail(3, "me_cuda_ref_update: invalid handle %d", (int)h);
  MeRefDev &d = r->dev;
  { ScopedTimer t(0, st); me_launch_planes(d_luma, stride, d.w, d.h, r->planes, d.pitch, d.plane_stride, r->search, d.spad, d.spitch, st); }
  g.launches += 2;
  CK(cudaGetLastError());
  r->gen = ++g.gen_counter;
  r->written_on = st;
  for (int l = 0; l < 2; l++)    // weighted copies of these planes are rebuilt with the next table
    for (int i = 0; i < g.list_n[l]; i++) if (g.twin[l][i].valid && g.twin[l][i].src == h) g.table_dirty = true;
  return 0;
}

// staging buffer for pictures on the device (u8, pitch = w)
uint8_t *g_d_stage = nullptr; size_t g_d_stage_bytes = 0;
int ensure_stage(size_t bytes)
{
  if (g_d_stage_bytes >= bytes) return 0;
  if (g_d_stage) cudaFree(g_d_stage);
  g_d_stage = nullptr; g_d_stage_bytes = 0;
  CK(cudaMalloc((void **)&g_d_stage, bytes));
  g_d_stage_bytes = bytes;
  return 0;
}

// Pictures are frames of the initialised size or fields of such frames (half the rows; interlace: PicInterlace / MbInterlace,
// dpb_split_field, mbuffer.c:2971): every reference picture carries its own size in its descriptor, the kernels take theirs from it
int check_size(int w, int h)
{
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  const bool frame = w == g.p.width && h == g.p.height, field = w == g.p.width && 2 * h == g.p.height && !(h & 15);
  if (!frame && !field) return fail(1, "picture size %dx%d is neither the initialised %dx%d nor a field of it", w, h, g.p.width, g.p.height);
  return 0;
}

// the current picture as the drop-in calls see it: view 0 the buffer as uploaded, 1 / 2 its top / bottom field (rows 2y / 2y + 1)
void apply_view(int view)
{
  const int par = view == 2 ? 1 : 0, mul = view ? 2 : 1;
  g.view = view;
  g.cur_view = g.d_cur + (size_t)par * g.cur_pitch * g.pel_bytes;
  g.cur_view_pitch = g.cur_pitch * mul;
  g.cur_view_h = g.cur_h / mul;
  if (g.d_cur_c) {
    g.cfg.chroma.cur_pitch = g.chroma_base_pitch * mul;
    for (int k = 0; k < 2; k++) g.cfg.chroma.cur[k] = g.d_cur_c + (size_t)k * g.chroma_base_pitch * g.cur_ch + (size_t)par * g.chroma_base_pitch;
  }
}
void picture_set(int h) { g.cur_h = h; apply_view(0); }

// More than 8 bits per sample: served by the row-pointer entry points (imgpel rows, as the reference holds its pictures) and the
// drop-in searches of the fast-full-search path; everything else says so instead of treating 16-bit samples as bytes.
int only8(const char *who)
{
  if (g.inited && g.pel_bytes != 1)
    return fail(1, "%s: not available for pictures of %d bits per sample (8-bit only)", who, g.p.bit_depth);
  return 0;
}

}  // namespace

extern "C" {

void me_cuda_set_error_handler(me_cuda_error_fn fn) { g_handler = fn; }
const char *me_cuda_last_error(void) { return g_err; }

int me_cuda_init(const me_cuda_params *p)
{
  if (!p) return fail(1, "me_cuda_init: null parameters");
  if (p->abi_version != ME_CUDA_ABI_VERSION) return fail(1, "ABI version %d, library is %d", p->abi_version, ME_CUDA_ABI_VERSION);
  if (g.inited) me_cuda_shutdown();
  if (p->bit_depth < 8 || p->bit_depth > 14) return fail(1, "bit depth %d outside 8..14", p->bit_depth);
  if (p->bit_depth > 8 && p->chroma_me_enable)
    return fail(1, "bit depth %d: chroma-aware ME is implemented for 8-bit pictures only", p->bit_depth);
  if (p->chroma_me_enable < 0 || p->chroma_me_enable > 2) return fail(1, "ChromaMEEnable %d outside 0..2", p->chroma_me_enable);
  if (p->chroma_me_enable) {
    if (p->yuv_format < 1 || p->yuv_format > 3)
      return fail(1, "ChromaMEEnable needs yuv_format 1 (4:2:0), 2 (4:2:2) or 3 (4:4:4)");
  }
  if (p->width <= 0 || p->height <= 0 || (p->width & 15) || (p->height & 15)) return fail(1, "picture size %dx%d must be positive multiples of 16", p->width, p->height);
  if (p->search_range < 1 || p->search_range > 64) return fail(1, "search range %d outside 1..64", p->search_range);
  if (p->max_num_references < 1 || p->max_num_references > ME_MAX_REFS) return fail(1, "max_num_references %d outside 1..%d", p->max_num_references, ME_MAX_REFS);
  int dev = p->device;
  if (dev < 0) { const char *e = getenv("ME_CUDA_DEVICE"); dev = e ? atoi(e) : 0; }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail(2, "no CUDA device available (%s); libme_cuda has no CPU path", cudaGetErrorString(e));
  if (dev >= count) return fail(2, "CUDA device %d requested, %d present", dev, count);
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major < 10) return fail(2, "device %s is sm_%d%d; libme_cuda is built for sm_100a only", prop.name, prop.major, prop.minor);
  g = State();
  g.p = *p; g.device = dev;
  g.pel_bytes = p->bit_depth > 8 ? 2 : 1;
  g.cfg.pel16 = p->bit_depth > 8; g.cfg.max_pel = (1 << p->bit_depth) - 1;
  g.cfg.search_range = p->search_range; g.cfg.rdopt = p->rdopt; g.cfg.full_search = p->full_search;
  for (int i = 0; i < 3; i++) g.cfg.metric[i] = p->me_metric[i];
  g.cfg.transform8x8 = p->transform8x8 != 0;
  g.cfg.mv_limit_min_y = p->mv_limit_min_y; g.cfg.mv_limit_max_y = p->mv_limit_max_y;
  g.cfg.start_hp = (p->chroma_me_enable == 1 || p->me_metric[0] != p->me_metric[1]) ? 0 : 1;      // mv-search.c:396
  g.cfg.start_qp = (p->chroma_me_enable == 1 || p->me_metric[1] != p->me_metric[2]) ? 0 : 1;      // mv-search.c:397
  memset(&g.cfg.chroma, 0, sizeof g.cfg.chroma);
  if (p->chroma_me_enable) {
    g.cfg.chroma.on = p->chroma_me_enable;                             // until me_cuda_set_chroma_state says otherwise
    g.cfg.chroma.weight = p->chroma_me_weight;
    g.cfg.chroma.shift_x = p->yuv_format == 3 ? 2 : 3; g.cfg.chroma.shift_y = p->yuv_format == 1 ? 3 : 2;       // lencod.c:2856-2879
    g.chroma_w = p->yuv_format == 3 ? p->width : p->width >> 1; g.chroma_h = p->yuv_format == 1 ? p->height >> 1 : p->height;
    g.cfg.chroma.cur_pitch = g.chroma_base_pitch = align_up(g.chroma_w, 16);
    g.cur_ch = g.chroma_h;
    CK(cudaMalloc((void **)&g.d_cur_c, (size_t)2 * g.cfg.chroma.cur_pitch * g.chroma_h));
    CK(cudaMemset(g.d_cur_c, 128, (size_t)2 * g.cfg.chroma.cur_pitch * g.chroma_h));
    g.cfg.chroma.cur[0] = g.d_cur_c; g.cfg.chroma.cur[1] = g.d_cur_c + (size_t)g.cfg.chroma.cur_pitch * g.chroma_h;
  }
  CK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&g.ev_order, cudaEventDisableTiming));
  g.cur_pitch = align_up(p->width, 128);
  CK(cudaMalloc((void **)&g.d_cur, (size_t)g.cur_pitch * p->height * g.pel_bytes));
  CK(cudaMalloc((void **)&g.d_table, sizeof g.h_table));
  g.has_cur_map = g.pel_bytes == 1 && me_make_tensor_map_2d(g.cur_map.bytes, g.d_cur, p->width, p->height, g.cur_pitch, 16, 16) == 0;
  picture_set(p->height);
  CK(cudaHostAlloc((void **)&g.h_result, 64, cudaHostAllocMapped));
  CK(cudaHostGetDevicePointer((void **)&g.d_result, g.h_result, 0));
  memset(g.h_result, 0, 64);
  {
    const size_t rec_bytes = sizeof(MePmsRecord) * ME_PMS_SLOTS * ME_PMS_MAX_REFS * 4;
    void *hp = nullptr, *dp = nullptr;
    CK(cudaHostAlloc(&hp, rec_bytes + 64, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer(&dp, hp, 0));
    memset(hp, 0, rec_bytes + 64);
    g.pms.h_rec = (MePmsRecord *)hp; g.pms.d_rec = (MePmsRecord *)dp;
    g.pms.h_ticket = (volatile int *)((char *)hp + rec_bytes); g.pms.d_ticket = (int *)((char *)dp + rec_bytes);
    CK(cudaMalloc((void **)&g.pms.d_done, sizeof(int)));
    CK(cudaMemset(g.pms.d_done, 0, sizeof(int)));
  }
  memset(g.d_surf, 0, sizeof g.d_surf);
  memset(g.setup_done, 0, sizeof g.setup_done);
  g_err[0] = 0;
  g.inited = true;
  return 0;
}

void me_cuda_shutdown(void)
{
  if (!g.inited) return;
  cudaSetDevice(g.device);
  cudaStreamSynchronize(g.stream);
  for (auto &r : g.refs) if (r.used) { cudaFree(r.planes); cudaFree(r.search); if (r.chroma) cudaFree(r.chroma); }
  g.refs.clear();
  if (g.d_cur_c) cudaFree(g.d_cur_c);
  for (int l = 0; l < 2; l++) for (int i = 0; i < ME_MAX_REFS; i++) { if (g.twin[l][i].planes) cudaFree(g.twin[l][i].planes); if (g.twin[l][i].search) cudaFree(g.twin[l][i].search); }
  for (int l = 0; l < 2; l++) for (int i = 0; i < ME_MAX_REFS; i++) if (g.d_surf[l][i]) cudaFree(g.d_surf[l][i]);
  cudaFree(g.d_cur); cudaFree(g.d_table); cudaFreeHost(g.h_result);
  if (g.pms.h_rec) cudaFreeHost(g.pms.h_rec);
  if (g.pms.d_done) cudaFree(g.pms.d_done);
  if (g.d_deferred) cudaFree(g.d_deferred);
  if (g.d_subpel_stats) cudaFree(g.d_subpel_stats);
  if (g.s_in) {
    cudaStreamSynchronize(g.s_in); cudaStreamSynchronize(g.s_out);
    for (int c = 0; c < ME_PIPE_CHUNKS; c++) { cudaEventDestroy(g.ev_in[c]); cudaEventDestroy(g.ev_done[c]); }
    cudaStreamDestroy(g.s_in); cudaStreamDestroy(g.s_out);
  }
  if (g.h_pin) cudaFreeHost(g.h_pin);
  if (g.d_wide) cudaFree(g.d_wide);
  if (g.h_wide) cudaFreeHost(g.h_wide);
  if (g.d_jobs) cudaFree(g.d_jobs);
  if (g.d_out) cudaFree(g.d_out);
  if (g.d_out_int) cudaFree(g.d_out_int);
  if (g.d_int_scratch) cudaFree(g.d_int_scratch);
  if (g.d_bi_jobs) cudaFree(g.d_bi_jobs);
  if (g.d_bi_out) cudaFree(g.d_bi_out);
  if (g.d_bi_table) cudaFree(g.d_bi_table);
  if (g.d_mc_table) cudaFree(g.d_mc_table);
  if (g.h_mc) cudaFreeHost(g.h_mc);
  if (g.d_mc) cudaFree(g.d_mc);
  if (g.d_db) cudaFree(g.d_db);
  if (g_d_stage) { cudaFree(g_d_stage); g_d_stage = nullptr; g_d_stage_bytes = 0; }
  for (int k = 0; k < 3; k++) { for (auto &p : g.ev[k]) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); } g.ev[k].clear(); }
  if (g.ev_order) cudaEventDestroy(g.ev_order);
  cudaStreamDestroy(g.stream);
  g = State();
}

void *me_cuda_stream(void) { return (void *)g.stream; }
int me_cuda_sync(void) { CK(cudaStreamSynchronize(g.stream)); return 0; }
int64_t me_cuda_launch_count(void) { return g.launches; }

// Share of the partitions the throughput sub-pel kernel reads from its staged region (the others take the clamped reads from
// global memory).  enable != 0 starts (and zeroes) the count, 0 stops it; served / total may be null.
int me_cuda_subpel_region_stats(int enable, int64_t *served, int64_t *total)
{
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  CK(cudaStreamSynchronize(g.stream));
  if (g.d_subpel_stats && (served || total)) {
    unsigned long long h[2] = { 0, 0 };
    CK(cudaMemcpy(h, g.d_subpel_stats, sizeof h, cudaMemcpyDeviceToHost));
    if (served) *served = (int64_t)h[0];
    if (total) *total = (int64_t)h[1];
  } else {
    if (served) *served = 0;
    if (total) *total = 0;
  }
  if (enable) {
    if (!g.d_subpel_stats) CK(cudaMalloc((void **)&g.d_subpel_stats, 2 * sizeof(unsigned long long)));
    CK(cudaMemset(g.d_subpel_stats, 0, 2 * sizeof(unsigned long long)));
  } else if (g.d_subpel_stats) {
    cudaFree(g.d_subpel_stats);
    g.d_subpel_stats = nullptr;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ pictures
int me_cuda_ref_create_device(const void *d_luma, int stride, int w, int h, void *stream, me_ref_handle *out)
{
  g.spec.valid = false;
  if (int rc = only8("me_cuda_ref_create_device")) return rc;
  if (int rc = check_size(w, h)) return rc;
  cudaStream_t st = stream ? (cudaStream_t)stream : g.stream;
  if (int rc = enter_stream(st)) return rc;
  if (int rc = create_ref_common((const uint8_t *)d_luma, stride, w, h, st, out)) return rc;
  return leave_stream(st);
}

int me_cuda_ref_create_u8(const uint8_t *luma, int stride, int w, int h, me_ref_handle *out)
{
  g.spec.valid = false;
  if (int rc = only8("me_cuda_ref_create_u8")) return rc;
  if (int rc = check_size(w, h)) return rc;
  size_t bytes = (size_t)w * h;
  if (int rc = ensure_pin(bytes)) return rc;
  if (int rc = ensure_stage(bytes)) return rc;
  CK(cudaStreamSynchronize(g.stream));               // the pinned staging buffer may still be in flight
  for (int y = 0; y < h; y++) memcpy(g.h_pin + (size_t)y * w, luma + (size_t)y * stride, w);
  CK(cudaMemcpyAsync(g_d_stage, g.h_pin, bytes, cudaMemcpyHostToDevice, g.stream));
  return create_ref_common(g_d_stage, w, w, h, g.stream, out);
}

int me_cuda_ref_create(const me_imgpel *const *rows, int w, int h, me_imgpel ****host_sub, me_ref_handle *out)
{
  g.spec.valid = false;
  if (int rc = check_size(w, h)) return rc;
  size_t bytes = (size_t)w * h * g.pel_bytes;
  if (int rc = ensure_pin(bytes)) return rc;
  if (int rc = ensure_stage(bytes)) return rc;
  CK(cudaStreamSynchronize(g.stream));
  for (int y = 0; y < h; y++) {
    const me_imgpel *s = rows[y];
    if (g.pel_bytes == 2) { memcpy(g.h_pin + (size_t)y * w * 2, s, (size_t)w * 2); continue; }     // imgpel is the 16-bit sample
    uint8_t *d = g.h_pin + (size_t)y * w;
    for (int x = 0; x < w; x++) d[x] = (uint8_t)s[x];
  }
  CK(cudaMemcpyAsync(g_d_stage, g.h_pin, bytes, cudaMemcpyHostToDevice, g.stream));
  if (int rc = create_ref_common(g_d_stage, w, w, h, g.stream, out)) return rc;
  if (host_sub && g.pel_bytes == 2) {
    // 16-bit planes go back as they are
    RefPic *r = ref_of(*out);
    const int wp = r->dev.wp, hp = r->dev.hp;
    const size_t n = (size_t)16 * wp * hp;
    if (g.wide_elems < n) {
      if (g.d_wide) cudaFree(g.d_wide);
      if (g.h_wide) cudaFreeHost(g.h_wide);
      g.d_wide = nullptr; g.h_wide = nullptr; g.wide_elems = 0;
      CK(cudaMalloc((void **)&g.d_wide, n * sizeof(uint16_t)));
      CK(cudaMallocHost((void **)&g.h_wide, n * sizeof(uint16_t)));
      g.wide_elems = n;
    }
    CK(cudaMemcpy2DAsync(g.h_wide, (size_t)wp * 2, r->planes, (size_t)r->dev.pitch * 2, (size_t)wp * 2, (size_t)16 * hp, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    for (int fy = 0; fy < 4; fy++)
      for (int fx = 0; fx < 4; fx++)
        for (int y = 0; y < hp; y++)
          memcpy(host_sub[fy][fx][y], g.h_wide + ((size_t)(fy * 4 + fx) * hp + y) * wp, (size_t)wp * sizeof(uint16_t));
    return 0;
  }
  if (host_sub) {
    // the encoder's motion compensation still reads imgY_sub on the host (macroblock.c:816): hand the planes back
    RefPic *r = ref_of(*out);
    int wp = r->dev.wp, hp = r->dev.hp;
    size_t n = (size_t)16 * wp * hp;
    if (g.wide_elems < n) {
      if (g.d_wide) cudaFree(g.d_wide);
      if (g.h_wide) cudaFreeHost(g.h_wide);
      g.d_wide = nullptr; g.h_wide = nullptr; g.wide_elems = 0;
      CK(cudaMalloc((void **)&g.d_wide, n * sizeof(uint16_t)));
      CK(cudaMallocHost((void **)&g.h_wide, n * sizeof(uint16_t)));
      g.wide_elems = n;
    }
    me_launch_widen_planes(r->planes, r->dev.pitch, r->dev.plane_stride, wp, hp, g.d_wide, g.stream);
    g.launches++;
    CK(cudaMemcpyAsync(g.h_wide, g.d_wide, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    for (int fy = 0; fy < 4; fy++)
      for (int fx = 0; fx < 4; fx++)
        for (int y = 0; y < hp; y++)
          memcpy(host_sub[fy][fx][y], g.h_wide + ((size_t)(fy * 4 + fx) * hp + y) * wp, (size_t)wp * sizeof(uint16_t));
  }
  return 0;
}

int me_cuda_ref_destroy(me_ref_handle h)
{
  g.spec.valid = false;
  RefPic *r = ref_of(h);
  if (!r) return fail(3, "me_cuda_ref_destroy: invalid handle %d", (int)h);
  CK(cudaStreamSynchronize(g.stream));
  cudaFree(r->planes); cudaFree(r->search);
  if (r->chroma) cudaFree(r->chroma);
  *r = RefPic();
  for (int l = 0; l < 2; l++) for (int i = 0; i < g.list_n[l]; i++) if (g.list[l][i] == h) { g.list[l][i] = 0; g.table_dirty = true; }
  return 0;
}

int me_cuda_ref_download_plane(me_ref_handle h, int fy, int fx, uint8_t *dst, int dst_stride)
{
  RefPic *r = ref_of(h);
  if (!r) return fail(3, "me_cuda_ref_download_plane: invalid handle %d", (int)h);
  if (fy < 0 || fy > 3 || fx < 0 || fx > 3) return fail(1, "plane index out of range");
  if (int rc = only8("me_cuda_ref_download_plane")) return rc;
  CK(cudaMemcpy2DAsync(dst, dst_stride, r->planes + (size_t)(fy * 4 + fx) * r->dev.plane_stride, r->dev.pitch,
                       r->dev.wp, r->dev.hp, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  return 0;
}

int me_cuda_ref_update_device(me_ref_handle hd, const void *d_luma, int stride, int w, int h, void *stream)
{
  g.spec.valid = false;
  if (int rc = only8("me_cuda_ref_update_device")) return rc;
  if (int rc = check_size(w, h)) return rc;
  if (RefPic *r = ref_of(hd)) if (r->dev.w != w || r->dev.h != h) return fail(1, "me_cuda_ref_update_device: a %dx%d picture into a %dx%d slot", w, h, r->dev.w, r->dev.h);
  cudaStream_t st = stream ? (cudaStream_t)stream : g.stream;
  if (int rc = enter_stream(st)) return rc;
  if (int rc = update_ref_common(hd, (const uint8_t *)d_luma, stride, st)) return rc;
  return leave_stream(st);
}

int me_cuda_ref_update_u8(me_ref_handle hd, const uint8_t *luma, int stride, int w, int h)
{
  g.spec.valid = false;
  if (int rc = only8("me_cuda_ref_update_u8")) return rc;
  if (int rc = check_size(w, h)) return rc;
  if (RefPic *r = ref_of(hd)) if (r->dev.w != w || r->dev.h != h) return fail(1, "me_cuda_ref_update_u8: a %dx%d picture into a %dx%d slot", w, h, r->dev.w, r->dev.h);
  size_t bytes = (size_t)w * h;
  if (int rc = ensure_stage(bytes)) return rc;
  // straight from the caller's buffer (pinned memory makes this an asynchronous DMA)
  CK(cudaMemcpy2DAsync(g_d_stage, w, luma, stride, w, h, cudaMemcpyHostToDevice, g.stream));
  return update_ref_common(hd, g_d_stage, w, g.stream);
}

int me_cuda_set_timing(int on)
{
  for (int k = 0; k < 3; k++) {
    for (auto &p : g.ev[k]) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    g.ev[k].clear();
  }
  g.timing = on != 0;
  return 0;
}

int me_cuda_kernel_time_ms(int which, double *total_ms, int *n_launches)
{
  if (which < 0 || which > 2) return fail(1, "me_cuda_kernel_time_ms: class %d out of range", which);
  double tot = 0;
  for (auto &p : g.ev[which]) {
    CK(cudaEventSynchronize(p.second));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, p.first, p.second));
    tot += ms;
  }
  *total_ms = tot;
  *n_launches = (int)g.ev[which].size();
  return 0;
}

int me_cuda_set_current_device(const void *d_luma, int stride, int w, int h, void *stream)
{
  g.spec.valid = false;
  if (int rc = only8("me_cuda_set_current_device")) return rc;
  if (int rc = check_size(w, h)) return rc;
  cudaStream_t st = stream ? (cudaStream_t)stream : g.stream;
  if (int rc = enter_stream(st)) return rc;
  CK(cudaMemcpy2DAsync(g.d_cur, g.cur_pitch, d_luma, stride, w, h, cudaMemcpyDeviceToDevice, st));
  picture_set(h);
  return leave_stream(st);
}

int me_cuda_set_current_u8(const uint8_t *luma, int stride, int w, int h)
{
  g.spec.valid = false;
  if (int rc = only8("me_cuda_set_current_u8")) return rc;
  if (int rc = check_size(w, h)) return rc;
  CK(cudaMemcpy2DAsync(g.d_cur, g.cur_pitch, luma, stride, w, h, cudaMemcpyHostToDevice, g.stream));
  picture_set(h);
  return 0;
}

int me_cuda_set_current(const me_imgpel *const *rows, int w, int h)
{
  g.spec.valid = false;
  if (int rc = check_size(w, h)) return rc;
  size_t bytes = (size_t)w * h * g.pel_bytes;
  if (int rc = ensure_pin(bytes)) return rc;
  CK(cudaStreamSynchronize(g.stream));
  for (int y = 0; y < h; y++) {
    const me_imgpel *s = rows[y];
    if (g.pel_bytes == 2) { memcpy(g.h_pin + (size_t)y * w * 2, s, (size_t)w * 2); continue; }
    uint8_t *d = g.h_pin + (size_t)y * w;
    for (int x = 0; x < w; x++) d[x] = (uint8_t)s[x];
  }
  CK(cudaMemcpy2DAsync(g.d_cur, (size_t)g.cur_pitch * g.pel_bytes, g.h_pin, (size_t)w * g.pel_bytes, (size_t)w * g.pel_bytes, h, cudaMemcpyHostToDevice, g.stream));
  picture_set(h);
  return 0;
}

// MBAFF / field coding of a frame whose original picture is on the device: the macroblock about to be searched is a frame
// macroblock (view 0), or belongs to the top (1) / bottom (2) field -- pCurImg = imgY_org_frm / imgY_org_top / imgY_org_bot
// (macroblock.c:165-190, image.c:2983, :3001; the field pictures are the even / odd rows of the frame, lencod.c:1970).  Applies to
// the original chroma planes as well.  The drop-in calls and me_cuda_pred_cost8x8 read the view; the frame batch needs view 0.
int me_cuda_select_current_view(int view)
{
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  if (view < 0 || view > 2) return fail(1, "me_cuda_select_current_view: view %d", view);
  if (view && (g.cur_h & 31)) return fail(1, "me_cuda_select_current_view: a picture of %d rows has no field views", g.cur_h);
  if (view != g.view) { g.spec.valid = false; apply_view(view); }
  return 0;
}

// ---- chroma (ChromaMEEnable)
// the weighted copies of a picture carry a copy of its descriptor (chroma pointers included): rebuilt after a change
static void invalidate_twins(me_ref_handle h)
{
  for (int l = 0; l < 2; l++)
    for (int i = 0; i < ME_MAX_REFS; i++) if (g.twin[l][i].src == h) g.twin[l][i].valid = false;
}
// luma_h: rows of the luma picture the planes belong to (the initialised height, or half of it: a field)
static int chroma_check(const char *who, int cw, int ch, int luma_h)
{
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  if (!g.p.chroma_me_enable) return fail(1, "%s: the library was initialised with chroma_me_enable = 0", who);
  const int want_h = luma_h == g.p.height ? g.chroma_h : g.chroma_h / 2;
  if (cw != g.chroma_w || ch != want_h) return fail(1, "%s: chroma size %dx%d differs from the expected %dx%d", who, cw, ch, g.chroma_w, want_h);
  return 0;
}
// both planes through the pinned staging buffer into a device allocation of 2 * pitch * ch bytes
static int upload_chroma(uint8_t *d_dst, int pitch, const uint8_t *u, const uint8_t *v, int stride, const me_imgpel *const *ur,
                         const me_imgpel *const *vr, int cw, int ch)
{
  const size_t bytes = (size_t)2 * cw * ch;
  if (int rc = ensure_pin(bytes)) return rc;
  CK(cudaStreamSynchronize(g.stream));               // the pinned staging buffer may still be in flight
  for (int k = 0; k < 2; k++)
    for (int y = 0; y < ch; y++) {
      uint8_t *d = g.h_pin + ((size_t)k * ch + y) * cw;
      if (ur) { const me_imgpel *s = (k ? vr : ur)[y]; for (int x = 0; x < cw; x++) d[x] = (uint8_t)s[x]; }
      else memcpy(d, (k ? v : u) + (size_t)y * stride, cw);
    }
  CK(cudaMemcpy2DAsync(d_dst, pitch, g.h_pin, cw, cw, (size_t)2 * ch, cudaMemcpyHostToDevice, g.stream));
  return 0;
}
static int ref_set_chroma_common(me_ref_handle h, const uint8_t *u, const uint8_t *v, int stride, const me_imgpel *const *ur,
                                 const me_imgpel *const *vr, int cw, int ch)
{
  g.spec.valid = false;
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  if (g.p.yuv_format == 3) return fail(1, "me_cuda_ref_set_chroma: in 4:4:4 the chroma sub-images are plane sets (me_cuda_ref_set_chroma_planes)");
  RefPic *r = ref_of(h);
  if (!r) return fail(3, "me_cuda_ref_set_chroma: invalid handle %d", (int)h);
  if (int rc = chroma_check("me_cuda_ref_set_chroma", cw, ch, r->dev.h)) return rc;
  const int pitch = align_up(cw, 16);
  if (!r->chroma) CK(cudaMalloc((void **)&r->chroma, (size_t)2 * pitch * ch));
  if (int rc = upload_chroma(r->chroma, pitch, u, v, stride, ur, vr, cw, ch)) return rc;
  r->dev.chroma[0] = r->chroma; r->dev.chroma[1] = r->chroma + (size_t)pitch * ch;
  r->dev.cw = cw; r->dev.ch = ch; r->dev.cpitch = pitch;
  invalidate_twins(h);
  g.table_dirty = true;
  return 0;
}
int me_cuda_ref_set_chroma(me_ref_handle h, const me_imgpel *const *u_rows, const me_imgpel *const *v_rows, int cw, int ch)
{
  return ref_set_chroma_common(h, nullptr, nullptr, 0, u_rows, v_rows, cw, ch);
}
int me_cuda_ref_set_chroma_u8(me_ref_handle h, const uint8_t *u, const uint8_t *v, int stride, int cw, int ch)
{
  return ref_set_chroma_common(h, u, v, stride, nullptr, nullptr, cw, ch);
}
int me_cuda_ref_set_chroma_planes(me_ref_handle h, me_ref_handle hu, me_ref_handle hv)
{
  g.spec.valid = false;
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  if (!g.p.chroma_me_enable || g.p.yuv_format != 3) return fail(1, "me_cuda_ref_set_chroma_planes is for ChromaMEEnable in 4:4:4");
  RefPic *r = ref_of(h), *ru = ref_of(hu), *rv = ref_of(hv);
  if (!r || !ru || !rv) return fail(3, "me_cuda_ref_set_chroma_planes: invalid handle (%d, %d, %d)", (int)h, (int)hu, (int)hv);
  r->chroma444[0] = hu; r->chroma444[1] = hv;
  r->dev.chroma[0] = ru->planes; r->dev.chroma[1] = rv->planes;
  r->dev.cw = ru->dev.w; r->dev.ch = ru->dev.h; r->dev.cpitch = ru->dev.pitch; r->dev.cplane_stride = ru->dev.plane_stride;
  invalidate_twins(h);
  g.table_dirty = true;
  return 0;
}
int me_cuda_set_current_chroma(const me_imgpel *const *u_rows, const me_imgpel *const *v_rows, int cw, int ch)
{
  g.spec.valid = false;
  if (int rc = chroma_check("me_cuda_set_current_chroma", cw, ch, g.cur_h)) return rc;      // after me_cuda_set_current of the same picture
  g.cur_ch = ch; apply_view(g.view);
  return upload_chroma(g.d_cur_c, g.chroma_base_pitch, nullptr, nullptr, 0, u_rows, v_rows, cw, ch);
}
int me_cuda_set_current_chroma_u8(const uint8_t *u, const uint8_t *v, int stride, int cw, int ch)
{
  g.spec.valid = false;
  if (int rc = chroma_check("me_cuda_set_current_chroma_u8", cw, ch, g.cur_h)) return rc;
  g.cur_ch = ch; apply_view(g.view);
  return upload_chroma(g.d_cur_c, g.chroma_base_pitch, u, v, stride, nullptr, nullptr, cw, ch);
}
int me_cuda_set_chroma_state(int enable)
{
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  if (enable && !g.p.chroma_me_enable) return fail(1, "me_cuda_set_chroma_state: the library was initialised with chroma_me_enable = 0");
  g.cfg.chroma.on = enable;
  g.spec.valid = false;
  return 0;
}

int me_cuda_set_ref_list(int list, int n, const me_ref_handle *handles)
{
  g.spec.valid = false;
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  if (list < 0 || list > 1 || n < 0 || n > ME_MAX_REFS) return fail(1, "me_cuda_set_ref_list: list %d / count %d out of range", list, n);
  for (int i = 0; i < n; i++) if (!ref_of(handles[i])) return fail(3, "me_cuda_set_ref_list: invalid handle %d at %d", (int)handles[i], i);
  bool same = g.list_n[list] == n;
  for (int i = 0; same && i < n; i++) same = g.list[list][i] == handles[i];
  if (same) return 0;
  g.list_n[list] = n;
  for (int i = 0; i < n; i++) g.list[list][i] = handles[i];
  g.table_dirty = true;
  return 0;
}

int me_cuda_set_slice_type(int t) { g.slice_type = t; g.spec.valid = false; return 0; }

int me_cuda_set_bipred_weights(int list, int apply, int weight1, int weight2, int offset_bi, int log_denom)
{
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  if (list < 0 || list > 1) return fail(1, "me_cuda_set_bipred_weights: list %d out of range", list);
  MeBiW w = {};
  if (apply) {
    // implicit weights reach 128 (w0 = 64 - w1 with w1 in -64..128, weighted_prediction.c:218-223), e.g. a list-1 picture that
    // lies in the past; the kernels compute in int
    if (weight1 < -128 || weight1 > 128 || weight2 < -128 || weight2 > 128 || offset_bi < -128 || offset_bi > 127 || log_denom < 0 || log_denom > 7)
      return fail(1, "me_cuda_set_bipred_weights: weights %d, %d / offset %d / log denominator %d out of range", weight1, weight2, offset_bi, log_denom);
    w.apply = 1; w.w1 = weight1; w.w2 = weight2; w.obi = offset_bi; w.denom = log_denom; w.maxv = g.cfg.max_pel;
    // chroma: like luma until me_cuda_set_bipred_chroma_weights says otherwise (explicit mode leaves chroma at the default
    // weights, i.e. the plain average; implicit mode gives chroma the luma weights, weighted_prediction.c:243-252)
    for (int k = 0; k < 2; k++) { w.cw1[k] = 1 << log_denom; w.cw2[k] = 1 << log_denom; w.cobi[k] = 0; }
  }
  g.bw[list] = w;
  g.spec.valid = false;
  return 0;
}

int me_cuda_set_bipred_chroma_weights(int list, int component, int weight1, int weight2, int offset_bi)
{
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  if (list < 0 || list > 1 || component < 0 || component > 1) return fail(1, "me_cuda_set_bipred_chroma_weights: list %d / component %d out of range", list, component);
  if (weight1 < -128 || weight1 > 128 || weight2 < -128 || weight2 > 128 || offset_bi < -128 || offset_bi > 127)
    return fail(1, "me_cuda_set_bipred_chroma_weights: weights %d, %d / offset %d out of range", weight1, weight2, offset_bi);
  if (!g.bw[list].apply) return fail(1, "me_cuda_set_bipred_chroma_weights: call me_cuda_set_bipred_weights(list %d, apply = 1, ...) first", list);
  g.bw[list].cw1[component] = weight1; g.bw[list].cw2[component] = weight2; g.bw[list].cobi[component] = offset_bi;
  g.spec.valid = false;
  return 0;
}

int me_cuda_set_ref_weights(int list, int ref, int apply, int weight, int offset, int log_denom)
{
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  if (list < 0 || list > 1 || ref < 0 || ref >= ME_MAX_REFS) return fail(1, "me_cuda_set_ref_weights: (list %d, ref %d) out of range", list, ref);
  WeightSet w;
  w.apply = apply != 0;
  if (w.apply) {
    if (weight < -128 || weight > 127 || offset < -128 || offset > 127 || log_denom < 0 || log_denom > 7)
      return fail(1, "me_cuda_set_ref_weights: weight %d / offset %d / log denominator %d out of range", weight, offset, log_denom);
    w.weight = weight; w.offset = offset; w.denom = log_denom;
  }
  if (g.wp[list][ref] == w) return 0;
  g.wp[list][ref] = w;
  g.table_dirty = true;
  g.spec.valid = false;
  return 0;
}

// ------------------------------------------------------------------------------------------------ drop-in calls
int me_cuda_mb_begin(int opix_x, int opix_y)
{
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  g.opix_x = opix_x; g.opix_y = opix_y;
  g.spec.valid = false;
  for (auto &it : g.pms.item) it.valid = false;
  memset(g.setup_done, 0, sizeof g.setup_done);
  return 0;
}

int me_cuda_ffs_is_set_up(int list, int ref)
{
  if (!g.inited || list < 0 || list > 1 || ref < 0 || ref >= ME_MAX_REFS) return 0;
  return g.setup_done[list][ref];
}

int me_cuda_ffs_setup(int list, int ref, int pmv16_x, int pmv16_y)
{
  g.spec.valid = false;
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  if (list < 0 || list > 1 || ref < 0 || ref >= g.list_n[list]) return fail(1, "me_cuda_ffs_setup: (list %d, ref %d) not in the reference lists", list, ref);
  RefPic *r = ref_of(g.list[list][ref]);
  if (!r) return fail(3, "me_cuda_ffs_setup: stale reference handle");
  if (g.cfg.chroma.on && !r->has_chroma()) return fail(3, "me_cuda_ffs_setup: (list %d, ref %d) has no chroma planes (me_cuda_ref_set_chroma)", list, ref);
  int R = setup_range_of(ref);
  int cx = pmv16_x / 4, cy = pmv16_y / 4;                                          // me_fullfast.c:552-553
  if (!g.cfg.rdopt) { cx = me_clamp(cx, -R, R); cy = me_clamp(cy, -R, R); }       // :555-560
  cx = me_clamp(cx, -2047 + R, 2047 - R);                                         // :562
  cy = me_clamp(cy, g.cfg.mv_limit_min_y + R, g.cfg.mv_limit_max_y - R);           // :563
  cx += g.opix_x; cy += g.opix_y;
  int P = (2 * g.cfg.search_range + 1) * (2 * g.cfg.search_range + 1);
  if (!g.d_surf[list][ref]) CK(cudaMalloc((void **)&g.d_surf[list][ref], sizeof(int32_t) * (size_t)ME_NUM_PARTITIONS * P));
  const MeRefDev *dev;
  if (int rc = search_planes(list, ref, r, &dev, nullptr, nullptr, nullptr, nullptr)) return rc;
  if (g.pel_bytes == 2)
    me_launch_ffs_surfaces16(dev, (const uint16_t *)g.cur_view, g.cur_view_pitch, g.opix_x, g.opix_y, cx, cy, R, g.cfg.metric[0] != 0, g.d_surf[list][ref], g.stream);
  else
  me_launch_ffs_surfaces(dev, g.cur_view, g.cur_view_pitch, g.opix_x, g.opix_y, cx, cy, R, g.cfg.metric[0] != 0, &g.cfg.chroma, g.d_surf[list][ref], g.stream);
  g.launches++;
  CK(cudaGetLastError());
  g.centre_x[list][ref] = cx; g.centre_y[list][ref] = cy; g.setup_range[list][ref] = R;
  g.setup_done[list][ref] = 1;
  return 0;
}

static int fetch_result(int *mvx, int *mvy, int *cost)
{
  // the kernel launched with g.ticket writes its three results and then the ticket into mapped host memory
  volatile int *r = g.h_result;
  for (unsigned spin = 0;; spin++) {
    if (r[3] == g.ticket) break;
    if ((spin & 0x3fff) == 0x3fff) {
      // the stream has drained without the ticket arriving: the kernel failed
      cudaError_t e = cudaStreamQuery(g.stream);
      if (e != cudaErrorNotReady && r[3] != g.ticket) {
        if (e == cudaSuccess) e = cudaGetLastError();
        return fail(2, "drop-in search kernel did not deliver its result: %s", cudaGetErrorString(e));
      }
    }
  }
  __sync_synchronize();
  *mvx = r[0]; *mvy = r[1]; *cost = r[2];
  return 0;
}

// ---- one launch per PartitionMotionSearch (include/me_cuda.h)
int me_cuda_pms_begin(const me_mb_motion *state, int n_items, const me_pms_item *items, const int *lambda)
{
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  if (!state || !items || !lambda || n_items < 1 || n_items > ME_PMS_MAX_ITEMS) return fail(1, "me_cuda_pms_begin: bad arguments");
  const int n_lists = g.slice_type == ME_B_SLICE ? 2 : 1;
  MePmsParams P;
  memset(&P, 0, sizeof P);
  P.st = *state;
  int n = 0;
  for (int l = 0; l < n_lists; l++)
    for (int r = 0; r < g.list_n[l]; r++) {
      if (n == ME_PMS_MAX_REFS) return -1;
      if (!g.setup_done[l][r]) return fail(1, "me_cuda_pms_begin: me_cuda_ffs_setup(list %d, ref %d) has not run for this macroblock", l, r);
      RefPic *rp = ref_of(g.list[l][r]);
      if (!rp) return fail(3, "me_cuda_pms_begin: stale reference handle");
      const MeRefDev *dev;
      if (int rc = search_planes(l, r, rp, &dev, nullptr, nullptr, nullptr, nullptr)) return rc;
      MePmsRef &R = P.refs[n];
      R.dev = *dev; R.surf = g.d_surf[l][r];
      R.offx = g.centre_x[l][r] - g.opix_x; R.offy = g.centre_y[l][r] - g.opix_y; R.setup_range = g.setup_range[l][r];
      R.list = l; R.ref = r;
      g.pms.lr_list[n] = l; g.pms.lr_ref[n] = r;
      n++;
    }
  if (!n) return -1;
  P.n_refs = n; P.n_items = n_items;
  for (int i = 0; i < n_items; i++) {
    if (items[i].blocktype < 1 || items[i].blocktype > 7 || items[i].block8x8 < 0 || items[i].block8x8 > 3)
      return fail(1, "me_cuda_pms_begin: item %d (block type %d, block %d) out of range", i, (int)items[i].blocktype, (int)items[i].block8x8);
    P.item_type[i] = items[i].blocktype; P.item_blk[i] = items[i].block8x8;
    // result slot: the one that already holds this call (stale), else a free one, else the oldest
    int slot = -1;
    for (int k = 0; k < ME_PMS_SLOTS && slot < 0; k++)
      if (g.pms.item[k].valid && g.pms.item[k].blocktype == items[i].blocktype && g.pms.item[k].block8x8 == items[i].block8x8) slot = k;
    for (int k = 0; k < ME_PMS_SLOTS && slot < 0; k++) if (!g.pms.item[k].valid) slot = k;
    if (slot < 0) { slot = 0; for (int k = 1; k < ME_PMS_SLOTS; k++) if (g.pms.item[k].age < g.pms.item[slot].age) slot = k; }
    auto &it = g.pms.item[slot];
    it.valid = true; it.blocktype = items[i].blocktype; it.block8x8 = items[i].block8x8; it.ticket = g.pms.ticket + 1; it.age = ++g.pms.clock;
    for (int k = 0; k < 3; k++) it.lambda[k] = lambda[k];
    memset(g.pms.ffs_served[slot], 0, sizeof g.pms.ffs_served[slot]);
    P.item_slot[i] = slot;
  }
  for (int i = 0; i < 3; i++) P.lambda[i] = lambda[i];
  P.opix_x = g.opix_x; P.opix_y = g.opix_y; P.slice_type = g.slice_type;
  // the sub-pel stage of BlockMotionSearch runs with ChromaMEEnable = (input->ChromaMEEnable == 2) (mv-search.c:779)
  MeCfgDev sub_cfg = g.cfg;
  sub_cfg.chroma.on = g.p.chroma_me_enable == 2 ? 1 : 0;
  g.spec.valid = false;
  g.pms.n_lr = n;
  me_launch_pms_group(&P, g.cur_view, g.cur_view_pitch, &sub_cfg, g.cfg.rdopt, g.pms.d_rec, g.pms.d_done, g.pms.d_ticket, ++g.pms.ticket, g.stream);
  g.launches++; g.pms.launches++;
  CK(cudaGetLastError());
  return 0;
}

int me_cuda_pms_cached(int blocktype, int block8x8)
{
  for (const auto &it : g.pms.item) if (it.valid && it.blocktype == blocktype && it.block8x8 == block8x8) return 1;
  return 0;
}

int me_cuda_pms_end(int blocktype, int block8x8)
{
  for (auto &it : g.pms.item) if (it.valid && it.blocktype == blocktype && it.block8x8 == block8x8) it.valid = false;
  return 0;
}

int me_cuda_pms_stats(int64_t *launches, int64_t *served, int64_t *missed)
{
  if (launches) *launches = g.pms.launches;
  if (served) *served = g.pms.served;
  if (missed) *missed = g.pms.missed;
  return 0;
}

// the record of block search (list, ref, blocktype, block at pic_pix) in the results of the last group launch, or nullptr;
// waits for the launch to finish
static MePmsRecord *pms_lookup(int list, int ref, int blocktype, int pic_pix_x, int pic_pix_y, bool **ffs_flag, const int **lambda)
{
  const int h = (pic_pix_x - g.opix_x) >> 2, v = (pic_pix_y - g.opix_y) >> 2;
  if (h < 0 || h > 3 || v < 0 || v > 3) return nullptr;
  for (int i = 0; i < ME_PMS_SLOTS; i++) {
    const auto &it = g.pms.item[i];
    if (!it.valid || it.blocktype != blocktype) continue;
    int hs[4], vs[4];
    const int n = me_pms_blocks(it.blocktype, it.block8x8, hs, vs);
    for (int k = 0; k < n; k++) {
      if (hs[k] != h || vs[k] != v) continue;
      for (int lr = 0; lr < g.pms.n_lr; lr++) {
        if (g.pms.lr_list[lr] != list || g.pms.lr_ref[lr] != ref) continue;
        // results and ticket are written by the kernel into mapped host memory; launches complete in ticket order
        for (unsigned spin = 0; (int)(*g.pms.h_ticket - it.ticket) < 0; spin++) {
          if ((spin & 0x3fff) == 0x3fff) {
            cudaError_t e = cudaStreamQuery(g.stream);
            if (e != cudaErrorNotReady && (int)(*g.pms.h_ticket - it.ticket) < 0) {
              fail(2, "partition-search kernel did not deliver its results: %s", cudaGetErrorString(e == cudaSuccess ? cudaGetLastError() : e));
              fatal("pms");
              return nullptr;
            }
          }
        }
        __sync_synchronize();
        *ffs_flag = &g.pms.ffs_served[i][lr][k];
        *lambda = it.lambda;
        return &g.pms.h_rec[(i * ME_PMS_MAX_REFS + lr) * 4 + k];
      }
    }
  }
  return nullptr;
}

int me_cuda_FastFullPelBlockMotionSearch(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                         short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y,
                                         int search_range, int min_mcost, int lambda_factor)
{
  if (g.inited && list >= 0 && list <= 1 && ref >= 0 && ref < ME_MAX_REFS && g.setup_done[list][ref]) {
    bool *flag = nullptr;
    const int *plam = nullptr;
    if (MePmsRecord *rec = pms_lookup(list, ref, blocktype, pic_pix_x, pic_pix_y, &flag, &plam)) {
      // answered from the group launch only when it searched with exactly these arguments
      if (rec->pmv_x == pred_mv_x && rec->pmv_y == pred_mv_y && lambda_factor == plam[0] && min_mcost == ME_INT_MAX &&
          search_range == me_block_range(g.cfg, ref, blocktype)) {
        *mv_x = (short)rec->int_mvx; *mv_y = (short)rec->int_mvy;
        *flag = true;
        g.spec.valid = false;
        g.pms.served++;
        return rec->int_cost;
      }
      g.pms.missed++;
    }
  }
  if (!g.inited || list < 0 || list > 1 || ref < 0 || ref >= ME_MAX_REFS || !g.setup_done[list][ref]) {
    fail(1, "me_cuda_FastFullPelBlockMotionSearch: me_cuda_ffs_setup(list %d, ref %d) has not run for this macroblock", list, (int)ref);
    fatal("ffs"); return ME_INT_MAX;
  }
  int blk = (pic_pix_y - g.opix_y) + ((pic_pix_x - g.opix_x) >> 2);                 // me_fullfast.c:854
  int p = me_part_index(blocktype, blk);
  if (p < 0) { fail(1, "block (%d,%d) is not a type-%d partition origin", pic_pix_x, pic_pix_y, blocktype); fatal("ffs"); return ME_INT_MAX; }
  int R = g.setup_range[list][ref], P = (2 * R + 1) * (2 * R + 1);
  const int offx = g.centre_x[list][ref] - g.opix_x, offy = g.centre_y[list][ref] - g.opix_y;
  RefPic *rp = ref < g.list_n[list] ? ref_of(g.list[list][ref]) : nullptr;
  g.spec.valid = false;
  const MeRefDev *rdev = nullptr;
  if (rp && search_planes(list, ref, rp, &rdev, nullptr, nullptr, nullptr, nullptr)) { fatal("ffs"); return ME_INT_MAX; }
  // (with ChromaMEEnable the sub-pel stage runs under a different chroma state than this call: no speculation)
  const bool speculate = g.sub_params.seen && g.sub_params.lambda[0] == lambda_factor && rp != nullptr && !g.p.chroma_me_enable;
  if (speculate) {
    MeSubpelArgs a{};
    a.pic_pix_x = pic_pix_x; a.pic_pix_y = pic_pix_y; a.blocktype = blocktype;
    a.pmv_x = pred_mv_x; a.pmv_y = pred_mv_y;
    a.search_pos2 = g.sub_params.pos2; a.search_pos4 = g.sub_params.pos4;
    for (int i = 0; i < 3; i++) a.lambda[i] = g.sub_params.lambda[i];
    a.check0 = (!g.cfg.rdopt && g.slice_type != ME_B_SLICE && ref == 0 && blocktype == 1);   // and mv == 0: decided on the device
    a.bi = 0;
    me_launch_ffs_block_subpel(g.d_surf[list][ref] + (size_t)p * P, R, offx, offy, search_range, min_mcost, g.cfg.rdopt, rdev,
                               g.cur_view, g.cur_view_pitch, &g.cfg, &a, g.d_result, ++g.ticket, g.stream);
  } else {
    me_launch_ffs_block(g.d_surf[list][ref] + (size_t)p * P, R, offx, offy, pred_mv_x, pred_mv_y, search_range, min_mcost,
                        lambda_factor, g.cfg.rdopt, g.d_result, ++g.ticket, g.stream);
  }
  g.launches++;
  int mx, my, cost;
  if (fetch_result(&mx, &my, &cost)) { fatal("ffs"); return ME_INT_MAX; }
  *mv_x = (short)mx; *mv_y = (short)my;
  if (speculate) {
    auto &sp = g.spec;
    sp.valid = true; sp.list = list; sp.ref = ref; sp.ppx = pic_pix_x; sp.ppy = pic_pix_y; sp.bt = blocktype;
    sp.pmvx = pred_mv_x; sp.pmvy = pred_mv_y;
    sp.mv_in_x = (short)(mx * 4); sp.mv_in_y = (short)(my * 4);
    sp.min_in = g.cfg.start_hp ? cost : ME_INT_MAX;
    for (int i = 0; i < 3; i++) sp.lambda[i] = g.sub_params.lambda[i];
    sp.pos2 = g.sub_params.pos2; sp.pos4 = g.sub_params.pos4;
    sp.mvx = g.h_result[4]; sp.mvy = g.h_result[5]; sp.cost = g.h_result[6];
  }
  return cost;
}

int me_cuda_FullPelBlockMotionSearch(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                     short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y,
                                     int search_range, int min_mcost, int lambda_factor)
{
  g.spec.valid = false;
  RefPic *r = (g.inited && list >= 0 && list <= 1 && ref >= 0 && ref < g.list_n[list]) ? ref_of(g.list[list][ref]) : nullptr;
  if (!r) { fail(1, "me_cuda_FullPelBlockMotionSearch: (list %d, ref %d) not in the reference lists", list, (int)ref); fatal("fullpel"); return ME_INT_MAX; }
  if (blocktype < 1 || blocktype > 7 || search_range < 0 || search_range > 64) {
    fail(1, "me_cuda_FullPelBlockMotionSearch: block type %d / search range %d out of range", blocktype, search_range); fatal("fullpel"); return ME_INT_MAX;
  }
  if (g.cfg.chroma.on && !r->has_chroma()) { fail(3, "me_cuda_FullPelBlockMotionSearch: reference without chroma planes (me_cuda_ref_set_chroma)"); fatal("fullpel"); return ME_INT_MAX; }
  const MeRefDev *dev;
  if (search_planes(list, ref, r, &dev, nullptr, nullptr, nullptr, nullptr)) { fatal("fullpel"); return ME_INT_MAX; }
  MeFullPelArgs a{};
  a.pic_pix_x = pic_pix_x; a.pic_pix_y = pic_pix_y; a.blocktype = blocktype;
  a.pmv_x = pred_mv_x; a.pmv_y = pred_mv_y; a.mv_x = *mv_x; a.mv_y = *mv_y;
  a.search_range = search_range; a.min_mcost = min_mcost; a.lambda = lambda_factor;
  a.check00 = (blocktype == 1 && !g.cfg.rdopt && g.slice_type != ME_B_SLICE && ref == 0);       // me_fullsearch.c:72
  me_launch_fullpel_block(dev, g.cur_view, g.cur_view_pitch, &g.cfg, &a, g.d_result, ++g.ticket, g.stream);
  g.launches++;
  int mx, my, cost;
  if (fetch_result(&mx, &my, &cost)) { fatal("fullpel"); return ME_INT_MAX; }
  *mv_x = (short)mx; *mv_y = (short)my;
  return cost;
}

int me_cuda_SubPelBlockMotionSearch(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                    short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y,
                                    int search_pos2, int search_pos4, int min_mcost, const int *lambda)
{
  RefPic *r = (g.inited && list >= 0 && list <= 1 && ref >= 0 && ref < g.list_n[list]) ? ref_of(g.list[list][ref]) : nullptr;
  if (!r) { fail(1, "me_cuda_SubPelBlockMotionSearch: (list %d, ref %d) not in the reference lists", list, (int)ref); fatal("subpel"); return ME_INT_MAX; }
  g.sub_params.seen = true;
  for (int i = 0; i < 3; i++) g.sub_params.lambda[i] = lambda[i];
  g.sub_params.pos2 = search_pos2; g.sub_params.pos4 = search_pos4;
  {
    bool *flag = nullptr;
    const int *plam = nullptr;
    MePmsRecord *rec = pms_lookup(list, ref, blocktype, pic_pix_x, pic_pix_y, &flag, &plam);
    if (rec && *flag) {                      // the integer-pel call of this block was answered from the same record
      if (rec->pmv_x == pred_mv_x && rec->pmv_y == pred_mv_y && *mv_x == (short)(rec->int_mvx * 4) && *mv_y == (short)(rec->int_mvy * 4) &&
          min_mcost == (g.cfg.start_hp ? rec->int_cost : ME_INT_MAX) && search_pos2 == 9 && search_pos4 == 9 &&
          lambda[0] == plam[0] && lambda[1] == plam[1] && lambda[2] == plam[2]) {
        *mv_x = (short)rec->mvx; *mv_y = (short)rec->mvy;
        g.spec.valid = false;
        g.pms.served++;
        return rec->cost;
      }
      g.pms.missed++;
    }
  }
  if (g.spec.valid) {
    const auto &sp = g.spec;
    const bool hit = sp.list == list && sp.ref == ref && sp.ppx == pic_pix_x && sp.ppy == pic_pix_y && sp.bt == blocktype &&
                     sp.pmvx == pred_mv_x && sp.pmvy == pred_mv_y && sp.mv_in_x == *mv_x && sp.mv_in_y == *mv_y &&
                     sp.min_in == min_mcost && sp.pos2 == search_pos2 && sp.pos4 == search_pos4 &&
                     sp.lambda[0] == lambda[0] && sp.lambda[1] == lambda[1] && sp.lambda[2] == lambda[2];
    g.spec.valid = false;
    if (hit) { *mv_x = (short)sp.mvx; *mv_y = (short)sp.mvy; g.spec_hits++; return sp.cost; }
  }
  MeSubpelArgs a{};
  a.pic_pix_x = pic_pix_x; a.pic_pix_y = pic_pix_y; a.blocktype = blocktype;
  a.pmv_x = pred_mv_x; a.pmv_y = pred_mv_y; a.mv_x = *mv_x; a.mv_y = *mv_y;
  a.search_pos2 = search_pos2; a.search_pos4 = search_pos4; a.min_mcost = min_mcost;
  for (int i = 0; i < 3; i++) a.lambda[i] = lambda[i];
  a.check0 = (!g.cfg.rdopt && g.slice_type != ME_B_SLICE && ref == 0 && blocktype == 1 && *mv_x == 0 && *mv_y == 0);   // me_fullsearch.c:361
  a.bi = 0;
  const MeRefDev *dev;
  if (search_planes(list, ref, r, &dev, nullptr, nullptr, nullptr, nullptr)) { fatal("subpel"); return ME_INT_MAX; }
  me_launch_subpel_single(dev, nullptr, g.cur_view, g.cur_view_pitch, &g.cfg, &a, g.d_result, ++g.ticket, g.stream);
  g.launches++;
  int mx, my, cost;
  if (fetch_result(&mx, &my, &cost)) { fatal("subpel"); return ME_INT_MAX; }
  *mv_x = (short)mx; *mv_y = (short)my;
  return cost;
}

int me_cuda_FullPelBlockMotionBiPred(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                     short pred_mv_x1, short pred_mv_y1, short pred_mv_x2, short pred_mv_y2,
                                     short *mv_x, short *mv_y, short *s_mv_x, short *s_mv_y,
                                     int search_range, int min_mcost, int lambda_factor)
{
  g.spec.valid = false;
  RefPic *r1 = (g.inited && list >= 0 && list <= 1 && ref >= 0 && ref < g.list_n[list]) ? ref_of(g.list[list][ref]) : nullptr;
  RefPic *r2 = (g.inited && g.list_n[list ^ 1] > 0) ? ref_of(g.list[list ^ 1][0]) : nullptr;
  if (!r1 || !r2) { fail(1, "me_cuda_FullPelBlockMotionBiPred: reference lists incomplete"); fatal("bipred"); return ME_INT_MAX; }
  if (search_range > 64) { fail(1, "bi-pred search range %d too large", search_range); fatal("bipred"); return ME_INT_MAX; }
  if (g.cfg.chroma.on && (!r1->has_chroma() || !r2->has_chroma())) { fail(3, "me_cuda_FullPelBlockMotionBiPred: reference without chroma planes (me_cuda_ref_set_chroma)"); fatal("bipred"); return ME_INT_MAX; }
  MeBiFullArgs a{};
  a.pic_pix_x = pic_pix_x; a.pic_pix_y = pic_pix_y; a.blocktype = blocktype;
  a.pmv_x1 = pred_mv_x1; a.pmv_y1 = pred_mv_y1; a.pmv_x2 = pred_mv_x2; a.pmv_y2 = pred_mv_y2;
  a.mv_x = *mv_x; a.mv_y = *mv_y; a.s_mv_x = *s_mv_x; a.s_mv_y = *s_mv_y;
  a.search_range = search_range; a.min_mcost = min_mcost; a.lambda = lambda_factor;
  a.bw = g.bw[list];
  me_launch_bipred_full(&r1->dev, &r2->dev, g.cur_view, g.cur_view_pitch, &g.cfg, &a, g.d_result, ++g.ticket, g.stream);
  g.launches++;
  int mx, my, cost;
  if (fetch_result(&mx, &my, &cost)) { fatal("bipred"); return ME_INT_MAX; }
  *mv_x = (short)mx; *mv_y = (short)my;
  return cost;
}

int me_cuda_SubPelBlockSearchBiPred(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                    short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y, short *s_mv_x, short *s_mv_y,
                                    int search_pos2, int search_pos4, int min_mcost, const int *lambda)
{
  g.spec.valid = false;
  RefPic *r1 = (g.inited && list >= 0 && list <= 1 && ref >= 0 && ref < g.list_n[list]) ? ref_of(g.list[list][ref]) : nullptr;
  RefPic *r2 = (g.inited && g.list_n[list ^ 1] > 0) ? ref_of(g.list[list ^ 1][0]) : nullptr;
  if (!r1 || !r2) { fail(1, "me_cuda_SubPelBlockSearchBiPred: reference lists incomplete"); fatal("bipred"); return ME_INT_MAX; }
  MeSubpelArgs a{};
  a.pic_pix_x = pic_pix_x; a.pic_pix_y = pic_pix_y; a.blocktype = blocktype;
  a.pmv_x = pred_mv_x; a.pmv_y = pred_mv_y; a.mv_x = *mv_x; a.mv_y = *mv_y; a.s_mv_x = *s_mv_x; a.s_mv_y = *s_mv_y;
  a.search_pos2 = search_pos2; a.search_pos4 = search_pos4; a.min_mcost = min_mcost;
  for (int i = 0; i < 3; i++) a.lambda[i] = lambda[i];
  a.check0 = 0; a.bi = 1;
  a.bw = g.bw[list];
  me_launch_subpel_single(&r1->dev, &r2->dev, g.cur_view, g.cur_view_pitch, &g.cfg, &a, g.d_result, ++g.ticket, g.stream);
  g.launches++;
  int mx, my, cost;
  if (fetch_result(&mx, &my, &cost)) { fatal("bipred"); return ME_INT_MAX; }
  *mv_x = (short)mx; *mv_y = (short)my;
  return cost;
}

// ------------------------------------------------------------------------------------------------ sub-pel stages of the pattern searches
// one launch per call (subpel.cu: me_subpel_pattern_kernel), result through mapped memory like the other drop-in calls
static int subpel_pattern(const char *who, const MePatArgs &pa, short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                          int pmv_x, int pmv_y, short *mv_x, short *mv_y, const short *s_mv, int search_pos2, int search_pos4,
                          int min_mcost, int lam0, int lam1, int lam2, int check0)
{
  g.spec.valid = false;
  if (blocktype < 1 || blocktype > 7) { fail(1, "%s: block type %d", who, blocktype); fatal(who); return ME_INT_MAX; }
  MeSubpelArgs a{};
  a.pic_pix_x = pic_pix_x; a.pic_pix_y = pic_pix_y; a.blocktype = blocktype;
  a.pmv_x = pmv_x; a.pmv_y = pmv_y; a.mv_x = *mv_x; a.mv_y = *mv_y;
  a.search_pos2 = search_pos2; a.search_pos4 = search_pos4; a.min_mcost = min_mcost;
  a.lambda[0] = lam0; a.lambda[1] = lam1; a.lambda[2] = lam2;
  a.check0 = check0;
  if (s_mv) {
    RefPic *r1 = (g.inited && list >= 0 && list <= 1 && ref >= 0 && ref < g.list_n[list]) ? ref_of(g.list[list][ref]) : nullptr;
    RefPic *r2 = (g.inited && g.list_n[list ^ 1] > 0) ? ref_of(g.list[list ^ 1][0]) : nullptr;
    if (!r1 || !r2) { fail(1, "%s: reference lists incomplete", who); fatal(who); return ME_INT_MAX; }
    a.bi = 1; a.s_mv_x = s_mv[0]; a.s_mv_y = s_mv[1];
    a.bw = g.bw[list];
    me_launch_subpel_pattern(&r1->dev, &r2->dev, g.cur_view, g.cur_view_pitch, &g.cfg, &a, &pa, g.d_result, ++g.ticket, g.stream);
  } else {
    RefPic *r = (g.inited && list >= 0 && list <= 1 && ref >= 0 && ref < g.list_n[list]) ? ref_of(g.list[list][ref]) : nullptr;
    if (!r) { fail(1, "%s: (list %d, ref %d) not in the reference lists", who, list, (int)ref); fatal(who); return ME_INT_MAX; }
    const MeRefDev *dev;
    if (search_planes(list, ref, r, &dev, nullptr, nullptr, nullptr, nullptr)) { fatal(who); return ME_INT_MAX; }
    me_launch_subpel_pattern(dev, nullptr, g.cur_view, g.cur_view_pitch, &g.cfg, &a, &pa, g.d_result, ++g.ticket, g.stream);
  }
  g.launches++;
  int mx, my, cost;
  if (fetch_result(&mx, &my, &cost)) { fatal(who); return ME_INT_MAX; }
  *mv_x = (short)mx; *mv_y = (short)my;
  return cost;
}

static const int8_t smp_shift[8] = { 0, 0, 1, 1, 2, 3, 3, 1 };   // block_type_shift_factor, me_umhexsmp.c:46

int me_cuda_EPZSSubPelBlockMotionSearch(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype, const short pred_mv[2], short mv[2],
                                        int search_pos2, int search_pos4, int min_mcost, const int *lambda, int subthres)
{
  MePatArgs pa{}; pa.pattern = 1; pa.subthres = subthres;
  return subpel_pattern("me_cuda_EPZSSubPelBlockMotionSearch", pa, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv[0], pred_mv[1],
                        &mv[0], &mv[1], nullptr, search_pos2, search_pos4, min_mcost, lambda[0], lambda[1], lambda[2], 0);
}

int me_cuda_EPZSSubPelBlockSearchBiPred(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype, const short *pred_mv1,
                                        const short *pred_mv2, short mv[2], const short s_mv[2], int search_pos2, int search_pos4,
                                        int min_mcost, const int *lambda)
{
  MePatArgs pa{}; pa.pattern = 1; pa.pmv2_x = pred_mv2[0]; pa.pmv2_y = pred_mv2[1];
  return subpel_pattern("me_cuda_EPZSSubPelBlockSearchBiPred", pa, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv1[0], pred_mv1[1],
                        &mv[0], &mv[1], s_mv, search_pos2, search_pos4, min_mcost, lambda[0], lambda[1], lambda[2], 0);
}

int me_cuda_UMHEXSubPelBlockMotionSearch(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype, short pred_mv_x, short pred_mv_y,
                                         short *mv_x, short *mv_y, int search_pos2, int search_pos4, int min_mcost, int lambda_factor)
{
  MePatArgs pa{}; pa.pattern = 2;
  return subpel_pattern("me_cuda_UMHEXSubPelBlockMotionSearch", pa, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y,
                        mv_x, mv_y, nullptr, search_pos2, search_pos4, min_mcost, lambda_factor, lambda_factor, lambda_factor, 0);
}

int me_cuda_smpUMHEXSubPelBlockMotionSearch(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype, short pred_mv_x, short pred_mv_y,
                                            short *mv_x, short *mv_y, int search_pos2, int search_pos4, int min_mcost, int lambda_factor,
                                            int uplayer_x, int uplayer_y, int threshold1, int threshold3)
{
  MePatArgs pa{}; pa.pattern = 3; pa.uplayer_x = uplayer_x; pa.uplayer_y = uplayer_y;
  pa.thr1 = threshold1 >> smp_shift[blocktype & 7]; pa.thr3 = threshold3 >> smp_shift[blocktype & 7];
  return subpel_pattern("me_cuda_smpUMHEXSubPelBlockMotionSearch", pa, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y,
                        mv_x, mv_y, nullptr, search_pos2, search_pos4, min_mcost, lambda_factor, lambda_factor, lambda_factor, 0);
}

int me_cuda_smpUMHEXFullSubPelBlockMotionSearch(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype, short pred_mv_x, short pred_mv_y,
                                                short *mv_x, short *mv_y, int search_pos2, int search_pos4, int min_mcost, int lambda_factor,
                                                int threshold1, int threshold3)
{
  MePatArgs pa{}; pa.pattern = 4;
  pa.thr1 = threshold1 >> smp_shift[blocktype & 7]; pa.thr3 = threshold3 >> smp_shift[blocktype & 7];
  const int check0 = (!g.cfg.rdopt && g.slice_type != ME_B_SLICE && ref == 0 && blocktype == 1 && *mv_x == 0 && *mv_y == 0);   // me_umhexsmp.c:442
  return subpel_pattern("me_cuda_smpUMHEXFullSubPelBlockMotionSearch", pa, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y,
                        mv_x, mv_y, nullptr, search_pos2, search_pos4, min_mcost, lambda_factor, lambda_factor, lambda_factor, check0);
}

// ------------------------------------------------------------------------------------------------ frame batch
static int use_basic_only()
{
  const char *e = getenv("ME_CUDA_FFS_KERNEL");
  return e && !strcmp(e, "basic");
}

int me_cuda_frame_search_device(const void *d_jobs, int n_jobs, void *d_out, void *d_out_int, void *stream)
{
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  if (n_jobs <= 0) return 0;
  if (g.view) return fail(1, "me_cuda_frame_search: a field view of the current picture is selected (the frame batch reads the picture as uploaded: upload the field itself)");
  cudaStream_t st = stream ? (cudaStream_t)stream : g.stream;
  if (int rc = upload_table()) return rc;                  // on the library stream (weighted copies are rebuilt there too)
  MeResultInt *di = (MeResultInt *)d_out_int;
  if (!di) {
    if (g.int_scratch_cap < n_jobs) {
      if (g.d_int_scratch) cudaFree(g.d_int_scratch);
      g.d_int_scratch = nullptr; g.int_scratch_cap = 0;
      CK(cudaMalloc((void **)&g.d_int_scratch, sizeof(MeResultInt) * (size_t)n_jobs));
      g.int_scratch_cap = n_jobs;
    }
    di = g.d_int_scratch;
  }
  if (g.deferred_cap < n_jobs) {
    if (g.d_deferred) cudaFree(g.d_deferred);
    g.d_deferred = nullptr; g.deferred_cap = 0;
    CK(cudaMalloc((void **)&g.d_deferred, sizeof(int) * ((size_t)n_jobs + 3)));
    g.deferred_cap = n_jobs;
  }
  int launched = 0;
  // a caller's stream: after everything queued on the library stream so far (planes, current picture, table, scratch in use)
  if (int rc = enter_stream(st)) return rc;
  CK(cudaMemsetAsync(g.d_deferred, 0, 3 * sizeof(int), st));     // both ticket counters and the hand-back count
  {
    ScopedTimer t(1, st);
    if (!use_basic_only() && g.maps_ok && g.has_cur_map) {
      launched = me_launch_ffs_batch_tuned((const MeJob *)d_jobs, n_jobs, g.d_table, g.h_maps, g.list_n[0], g.list_n[1], g.cur_map.bytes,
                                           g.d_cur, g.cur_pitch, &g.cfg, di, g.d_deferred + 2, g.d_deferred, st);
      // the jobs the tuned kernel handed back (lambda factors beyond its 32-bit keys; normally none: the launch finds an empty list)
      if (launched) { me_launch_ffs_batch_basic((const MeJob *)d_jobs, n_jobs, g.d_table, g.d_cur, g.cur_pitch, &g.cfg, di, g.d_deferred + 2, st); launched++; }
    }
    if (!launched) { me_launch_ffs_batch_basic((const MeJob *)d_jobs, n_jobs, g.d_table, g.d_cur, g.cur_pitch, &g.cfg, di, nullptr, st); launched = 1; }
  }
  {
    ScopedTimer t(2, st);
    // the sub-pel stage of BlockMotionSearch runs with ChromaMEEnable = (input->ChromaMEEnable == 2) (mv-search.c:779)
    MeCfgDev sub_cfg = g.cfg;
    sub_cfg.chroma.on = (g.cfg.chroma.on && g.p.chroma_me_enable == 2) ? 1 : 0;
    int done = 0;
    if (!use_basic_only() && g.pmaps_ok)
      done = me_launch_subpel_batch_tuned((const MeJob *)d_jobs, n_jobs, g.d_table, g.h_pmaps, g.list_n[0], g.list_n[1], g.d_cur, g.cur_pitch,
                                          &sub_cfg, di, (MeResult *)d_out, g.d_deferred + 1, g.d_subpel_stats, st);
    if (!done) me_launch_subpel_batch((const MeJob *)d_jobs, n_jobs, g.d_table, g.d_cur, g.cur_pitch, &sub_cfg, di, (MeResult *)d_out, st);
  }
  g.launches += launched + 1;
  CK(cudaGetLastError());
  return leave_stream(st);
}

// job fields the kernels index with: checked on the host, chunk by chunk, while the chunk's upload is already in flight
static int check_jobs(const me_mb_job *jobs, int first, int n)
{
  for (int i = first; i < first + n; i++) {
    const me_mb_job &j = jobs[i];
    if (j.list < 0 || j.list > 1 || j.ref < 0 || j.ref >= g.list_n[j.list])
      return fail(1, "job %d: (list %d, ref %d) not in the reference lists", i, (int)j.list, (int)j.ref);
    if (j.opix_x < 0 || j.opix_y < 0 || j.opix_x + 16 > g.p.width || j.opix_y + 16 > g.cur_h || (j.opix_x & 15) || (j.opix_y & 15))
      return fail(1, "job %d: macroblock origin (%d,%d) invalid", i, (int)j.opix_x, (int)j.opix_y);
  }
  return 0;
}

int me_cuda_frame_search(const me_mb_job *jobs, int n_jobs, me_mb_result *out, me_mb_result_int *out_int)
{
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  if (n_jobs <= 0) return 0;
  // Pinned (page-locked) caller buffers are visible to the device under unified addressing: the kernels then read the job array and
  // write the result arrays in place -- descriptors are fetched two jobs / one job ahead of their use, results leave as contiguous
  // 164-byte runs -- and the whole batch is one search launch and one sub-pel launch with no copy in front or behind.
  // (ME_CUDA_ZEROCOPY=0: always the chunked copies below.)
  {
    static int zero_copy = -1;
    if (zero_copy < 0) { const char *e = getenv("ME_CUDA_ZEROCOPY"); zero_copy = !(e && e[0] == '0'); }
    auto dev_ptr = [](const void *p) -> void * {
      cudaPointerAttributes a;
      if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
      return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
    };
    void *dj = zero_copy ? dev_ptr(jobs) : nullptr, *dout = dj ? dev_ptr(out) : nullptr, *dint = (dout && out_int) ? dev_ptr(out_int) : nullptr;
    if (dj && dout && (!out_int || dint)) {
      // the host does not touch the job array either: a small kernel in front of the searches checks the fields they index with
      // (the check_jobs() of the copying path) and the batch kernels return at once when it objects
      volatile int *bad = g.h_result + 8;                              // mapped pinned, beside the drop-in result words
      *bad = 0;
      me_launch_check_jobs((const MeJob *)dj, n_jobs, g.list_n[0], g.list_n[1], g.p.width, g.cur_h, g.d_result + 8, g.stream);
      g.launches++;
      g.cfg.abort_flag = g.d_result + 8;
      int rc = me_cuda_frame_search_device(dj, n_jobs, dout, dint, g.stream);
      g.cfg.abort_flag = nullptr;
      cudaError_t e = cudaStreamSynchronize(g.stream);               // whatever happened, nothing touches the caller's buffers after return
      if (!rc && e != cudaSuccess) rc = fail(2, "cudaStreamSynchronize failed: %s", cudaGetErrorString(e));
      if (!rc && *bad) {
        const me_mb_job &j = jobs[*bad - 1];
        rc = fail(1, "job %d: (list %d, ref %d) / macroblock origin (%d,%d) invalid", *bad - 1, (int)j.list, (int)j.ref, (int)j.opix_x, (int)j.opix_y);
      }
      return rc;
    }
  }
  if (int rc = ensure_batch(n_jobs)) return rc;
  // Large batches are cut into chunks that move through three streams -- jobs in (H2D engine), search (library stream),
  // results out (D2H engine) -- so that with pinned caller buffers only the first upload and the last download are exposed.
  // chunk sizes in twentieths of the batch: a small first chunk starts the kernels early, a small last one keeps the exposed
  // download short (ME_CUDA_PIPE_WEIGHTS="a,b,..." overrides; measured on B200, 1080p x 5 references)
  static int weights[ME_PIPE_CHUNKS], n_weights = 0;
  if (!n_weights) {
    const char *e = getenv("ME_CUDA_PIPE_WEIGHTS");
    if (e) for (const char *q = e; *q && n_weights < ME_PIPE_CHUNKS; ) { int v = atoi(q); if (v > 0) weights[n_weights++] = v; while (*q && *q != ',') q++; if (*q) q++; }
    if (!n_weights) { const int def[4] = { 5, 5, 5, 5 }; for (int i = 0; i < 4; i++) weights[n_weights++] = def[i]; }
  }
  const int n_chunks = n_jobs >= 8192 ? n_weights : 1;
  if (!g.s_in) {
    CK(cudaStreamCreateWithFlags(&g.s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&g.s_out, cudaStreamNonBlocking));
    for (int c = 0; c < ME_PIPE_CHUNKS; c++) {
      CK(cudaEventCreateWithFlags(&g.ev_in[c], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&g.ev_done[c], cudaEventDisableTiming));
    }
  }
  int c_off[ME_PIPE_CHUNKS + 1];
  {
    int wsum = 0, acc = 0;
    for (int c = 0; c < n_chunks; c++) wsum += weights[c];
    c_off[0] = 0;
    for (int c = 0; c < n_chunks; c++) { acc += weights[c]; c_off[c + 1] = n_chunks == 1 ? n_jobs : (int)((long long)n_jobs * acc / wsum); }
  }
  // from here on errors are collected, not returned: whatever happens, the three streams are drained before the caller gets
  // its buffers back (nothing may still be reading the job array or writing the result arrays on return)
  int rc = 0;
  auto cu = [&](cudaError_t e, const char *what) {
    if (e != cudaSuccess && !rc) rc = fail(2, "%s failed: %s", what, cudaGetErrorString(e));
    return e == cudaSuccess;
  };
  for (int c = 0; c < n_chunks && !rc; c++) {
    const int off = c_off[c], n = c_off[c + 1] - off;
    if (n <= 0) continue;
    cu(cudaMemcpyAsync(g.d_jobs + off, jobs + off, sizeof(MeJob) * (size_t)n, cudaMemcpyHostToDevice, g.s_in), "job upload") &&
      cu(cudaEventRecord(g.ev_in[c], g.s_in), "cudaEventRecord");
  }
  if (!rc) rc = upload_table();                                // on the library stream, before any chunk's kernels
  for (int c = 0; c < n_chunks && !rc; c++) {
    const int off = c_off[c], n = c_off[c + 1] - off;
    if (n <= 0) continue;
    if ((rc = check_jobs(jobs, off, n)) != 0) break;           // host work hidden behind the uploads / the previous chunk's kernels
    if (!cu(cudaStreamWaitEvent(g.stream, g.ev_in[c], 0), "cudaStreamWaitEvent")) break;
    if ((rc = me_cuda_frame_search_device(g.d_jobs + off, n, g.d_out + off, g.d_out_int + off, g.stream)) != 0) break;
    if (!cu(cudaEventRecord(g.ev_done[c], g.stream), "cudaEventRecord") ||
        !cu(cudaStreamWaitEvent(g.s_out, g.ev_done[c], 0), "cudaStreamWaitEvent") ||
        !cu(cudaMemcpyAsync(out + off, g.d_out + off, sizeof(MeResult) * (size_t)n, cudaMemcpyDeviceToHost, g.s_out), "result download")) break;
    if (out_int) cu(cudaMemcpyAsync(out_int + off, g.d_out_int + off, sizeof(MeResultInt) * (size_t)n, cudaMemcpyDeviceToHost, g.s_out), "result download");
  }
  cu(cudaStreamSynchronize(g.s_in), "cudaStreamSynchronize");
  cu(cudaStreamSynchronize(g.stream), "cudaStreamSynchronize");
  cu(cudaStreamSynchronize(g.s_out), "cudaStreamSynchronize");
  return rc;
}


// ------------------------------------------------------------------------------------------------ prediction and mode-decision costs
static const size_t MC_MAPPED = 64 * 1024;

// table of the listed pictures' own plane sets (never the weighted copies of the searches), uploaded when the lists changed
static int mc_table()
{
  MeRefDev t[2 * ME_MAX_REFS];
  memset(t, 0, sizeof t);
  for (int l = 0; l < 2; l++)
    for (int i = 0; i < g.list_n[l]; i++) {
      RefPic *r = ref_of(g.list[l][i]);
      if (!r) return fail(3, "reference list %d entry %d holds an invalid handle", l, i);
      if (r->written_on && r->written_on != g.stream) CK(cudaStreamSynchronize(r->written_on));
      t[l * ME_MAX_REFS + i] = r->dev;
    }
  if (!g.d_mc_table) CK(cudaMalloc((void **)&g.d_mc_table, sizeof t));
  if (!g.mc_table_valid || memcmp(t, g.h_mc_table, sizeof t)) {
    CK(cudaStreamSynchronize(g.stream));                  // a kernel may still be reading the old table
    memcpy(g.h_mc_table, t, sizeof t);
    CK(cudaMemcpyAsync(g.d_mc_table, g.h_mc_table, sizeof t, cudaMemcpyHostToDevice, g.stream));
    g.mc_table_valid = true;
  }
  return 0;
}

static int mc_check_blocks(const me_pred_block *b, int n, int bsx, int bsy, const char *who)
{
  for (int i = 0; i < n; i++) {
    if (b[i].p_dir > 2) return fail(1, "%s: block %d: p_dir %d", who, i, (int)b[i].p_dir);
    if (b[i].p_dir < 0) continue;
    for (int l = 0; l < 2; l++)
      if ((b[i].p_dir == l || b[i].p_dir == 2) && (b[i].ref[l] < 0 || b[i].ref[l] >= g.list_n[l]))
        return fail(1, "%s: block %d: reference %d not in list %d", who, i, (int)b[i].ref[l], l);
    if (b[i].pix_x < 0 || b[i].pix_y < 0 || b[i].pix_x + bsx > g.p.width || b[i].pix_y + bsy > g.cur_view_h)
      return fail(1, "%s: block %d at (%d,%d) outside the picture", who, i, (int)b[i].pix_x, (int)b[i].pix_y);
    if (b[i].apply_weights && (b[i].log_denom < 0 || b[i].log_denom > 7))
      return fail(1, "%s: block %d: log weight denominator %d", who, i, (int)b[i].log_denom);
  }
  return 0;
}

// staging for `in` bytes of requests and `out` bytes of results: *d_in / *d_out device addresses, *h_out where the results
// appear on the host (nullptr: copy them from *d_out)
static int mc_stage(const void *src, size_t in, size_t out, uint8_t **d_in, uint8_t **d_out, uint8_t **h_out)
{
  const size_t in_al = (in + 255) & ~(size_t)255;
  if (in_al + out <= MC_MAPPED) {
    if (!g.h_mc) {
      CK(cudaHostAlloc((void **)&g.h_mc, MC_MAPPED, cudaHostAllocMapped));
      CK(cudaHostGetDevicePointer((void **)&g.d_mc_mapped, g.h_mc, 0));
    }
    memcpy(g.h_mc, src, in);
    *d_in = g.d_mc_mapped; *d_out = g.d_mc_mapped + in_al; *h_out = g.h_mc + in_al;
    return 0;
  }
  if (g.d_mc_bytes < in_al + out) {
    CK(cudaStreamSynchronize(g.stream));
    if (g.d_mc) cudaFree(g.d_mc);
    g.d_mc = nullptr; g.d_mc_bytes = 0;
    CK(cudaMalloc((void **)&g.d_mc, in_al + out));
    g.d_mc_bytes = in_al + out;
  }
  CK(cudaMemcpyAsync(g.d_mc, src, in, cudaMemcpyHostToDevice, g.stream));
  *d_in = g.d_mc; *d_out = g.d_mc + in_al; *h_out = nullptr;
  return 0;
}

int me_cuda_luma_prediction(const me_pred_block *blocks, int n, int bsx, int bsy, uint8_t *pred)
{
  if (int rc = only8("me_cuda_luma_prediction")) return rc;
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  if (n <= 0) return 0;
  if ((bsx != 4 && bsx != 8 && bsx != 16) || (bsy != 4 && bsy != 8 && bsy != 16)) return fail(1, "me_cuda_luma_prediction: block size %d x %d", bsx, bsy);
  if (int rc = mc_check_blocks(blocks, n, bsx, bsy, "me_cuda_luma_prediction")) return rc;
  if (int rc = mc_table()) return rc;
  const size_t in = sizeof(me_pred_block) * (size_t)n, out = (size_t)n * bsx * bsy;
  uint8_t *d_in, *d_out, *h_out;
  if (int rc = mc_stage(blocks, in, out, &d_in, &d_out, &h_out)) return rc;
  me_launch_luma_prediction(d_in, n, bsx, bsy, g.d_mc_table, d_out, g.stream);
  g.launches++;
  CK(cudaGetLastError());
  if (!h_out) CK(cudaMemcpyAsync(pred, d_out, out, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  if (h_out) memcpy(pred, h_out, out);
  return 0;
}

int me_cuda_pred_cost8x8(const me_pred_block *blocks, int n_regions, int metric, int32_t *cost4, int32_t *cost8)
{
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  if (n_regions <= 0) return 0;
  if (metric < 0 || metric > 2) return fail(1, "me_cuda_pred_cost8x8: metric %d", metric);
  if (int rc = mc_check_blocks(blocks, 4 * n_regions, 4, 4, "me_cuda_pred_cost8x8")) return rc;
  if (int rc = mc_table()) return rc;
  const size_t in = sizeof(me_pred_block) * 4 * (size_t)n_regions, out = 2 * sizeof(int32_t) * (size_t)n_regions;
  uint8_t *d_in, *d_out, *h_out;
  if (int rc = mc_stage(blocks, in, out, &d_in, &d_out, &h_out)) return rc;
  int32_t *d4 = (int32_t *)d_out, *d8 = d4 + n_regions;
  me_launch_pred_cost8x8(d_in, n_regions, metric, g.d_mc_table, g.cur_view, g.cur_view_pitch, g.pel_bytes == 2, g.cfg.max_pel, d4, d8, g.stream);
  g.launches++;
  CK(cudaGetLastError());
  if (!h_out) {
    CK(cudaMemcpyAsync(cost4, d4, sizeof(int32_t) * (size_t)n_regions, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaMemcpyAsync(cost8, d8, sizeof(int32_t) * (size_t)n_regions, cudaMemcpyDeviceToHost, g.stream));
  }
  CK(cudaStreamSynchronize(g.stream));
  if (h_out) {
    memcpy(cost4, h_out, sizeof(int32_t) * (size_t)n_regions);
    memcpy(cost8, h_out + sizeof(int32_t) * (size_t)n_regions, sizeof(int32_t) * (size_t)n_regions);
  }
  return 0;
}

int me_cuda_ref_download_region(me_ref_handle h, int fy, int fx, int x0, int y0, int w, int hgt, me_imgpel *dst, int dst_stride)
{
  RefPic *r = ref_of(h);
  if (!r) return fail(3, "me_cuda_ref_download_region: invalid handle %d", (int)h);
  if (fy < 0 || fy > 3 || fx < 0 || fx > 3) return fail(1, "plane index out of range");
  if (x0 < 0 || y0 < 0 || w <= 0 || hgt <= 0 || x0 + w > r->dev.wp || y0 + hgt > r->dev.hp)
    return fail(1, "me_cuda_ref_download_region: rectangle (%d,%d) %d x %d outside the %d x %d plane", x0, y0, w, hgt, r->dev.wp, r->dev.hp);
  const size_t bytes = (size_t)w * hgt * sizeof(uint16_t);
  uint8_t *d_in, *d_out, *h_out;
  uint32_t dummy = 0;
  if (r->written_on && r->written_on != g.stream) CK(cudaStreamSynchronize(r->written_on));
  if (g.pel_bytes == 2) {          // 16-bit planes: the rectangle as it is
    const uint16_t *src = (const uint16_t *)r->planes + (size_t)(fy * 4 + fx) * r->dev.plane_stride + (size_t)y0 * r->dev.pitch + x0;
    CK(cudaMemcpy2DAsync(dst, (size_t)dst_stride * sizeof(uint16_t), src, (size_t)r->dev.pitch * sizeof(uint16_t), (size_t)w * sizeof(uint16_t), hgt,
                         cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return 0;
  }
  if (int rc = mc_stage(&dummy, sizeof dummy, bytes, &d_in, &d_out, &h_out)) return rc;
  me_launch_widen_region(r->planes + (size_t)(fy * 4 + fx) * r->dev.plane_stride, r->dev.pitch, x0, y0, w, hgt, (uint16_t *)d_out, g.stream);
  g.launches++;
  CK(cudaGetLastError());
  if (!h_out) {
    CK(cudaMemcpy2DAsync(dst, (size_t)dst_stride * sizeof(uint16_t), d_out, (size_t)w * sizeof(uint16_t), (size_t)w * sizeof(uint16_t), hgt,
                         cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return 0;
  }
  CK(cudaStreamSynchronize(g.stream));
  for (int y = 0; y < hgt; y++) memcpy(dst + (size_t)y * dst_stride, h_out + (size_t)y * w * sizeof(uint16_t), (size_t)w * sizeof(uint16_t));
  return 0;
}

// ------------------------------------------------------------------------------------------------ deblocking
int me_cuda_deblock_frame(me_imgpel *const *y_rows, me_imgpel *const *u_rows, me_imgpel *const *v_rows, int w, int h, int cw, int ch,
                          const me_deblock_mb *mbs)
{
  if (int rc = only8("me_cuda_deblock_frame")) return rc;
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  if (w <= 0 || h <= 0 || (w & 15) || (h & 15)) return fail(1, "me_cuda_deblock_frame: picture size %d x %d", w, h);
  const bool chroma = u_rows && v_rows;
  if (chroma && (cw != w / 2 || ch != h / 2)) return fail(1, "me_cuda_deblock_frame: chroma %d x %d is not 4:2:0 of %d x %d", cw, ch, w, h);
  const int mbw = w / 16, mbh = h / 16;
  if (mbh > 148 * 16) return fail(1, "me_cuda_deblock_frame: %d macroblock rows exceed what is resident at once", mbh);
  const size_t ybytes = (size_t)w * h, cbytes = chroma ? (size_t)cw * ch : 0, pix = ybytes + 2 * cbytes;
  const size_t mb_bytes = sizeof(me_deblock_mb) * (size_t)mbw * mbh, mb_off = (pix + 255) & ~(size_t)255;
  const size_t prog_off = (mb_off + mb_bytes + 255) & ~(size_t)255, total = prog_off + sizeof(int) * (size_t)mbh;
  if (int rc = ensure_pin(mb_off + mb_bytes)) return rc;
  CK(cudaStreamSynchronize(g.stream));
  g.db_valid = false;
  if (g.d_db_bytes < total) {
    if (g.d_db) cudaFree(g.d_db);
    g.d_db = nullptr; g.d_db_bytes = 0;
    CK(cudaMalloc((void **)&g.d_db, total));
    g.d_db_bytes = total;
  }
  for (int y = 0; y < h; y++) { const me_imgpel *s = y_rows[y]; uint8_t *d = g.h_pin + (size_t)y * w; for (int x = 0; x < w; x++) d[x] = (uint8_t)s[x]; }
  for (int k = 0; k < (chroma ? 2 : 0); k++)
    for (int y = 0; y < ch; y++) {
      const me_imgpel *s = (k ? v_rows : u_rows)[y]; uint8_t *d = g.h_pin + ybytes + (size_t)k * cbytes + (size_t)y * cw;
      for (int x = 0; x < cw; x++) d[x] = (uint8_t)s[x];
    }
  memcpy(g.h_pin + mb_off, mbs, mb_bytes);
  CK(cudaMemcpyAsync(g.d_db, g.h_pin, mb_off + mb_bytes, cudaMemcpyHostToDevice, g.stream));
  me_launch_deblock(g.d_db, w, chroma ? g.d_db + ybytes : nullptr, chroma ? g.d_db + ybytes + cbytes : nullptr, cw, mbw, mbh,
                    g.d_db + mb_off, (int *)(g.d_db + prog_off), g.stream);
  g.launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(g.h_pin, g.d_db, pix, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  for (int y = 0; y < h; y++) { me_imgpel *d = y_rows[y]; const uint8_t *s = g.h_pin + (size_t)y * w; for (int x = 0; x < w; x++) d[x] = s[x]; }
  for (int k = 0; k < (chroma ? 2 : 0); k++)
    for (int y = 0; y < ch; y++) {
      me_imgpel *d = (k ? v_rows : u_rows)[y]; const uint8_t *s = g.h_pin + ybytes + (size_t)k * cbytes + (size_t)y * cw;
      for (int x = 0; x < cw; x++) d[x] = s[x];
    }
  g.db_w = w; g.db_h = h; g.db_valid = true;
  return 0;
}

int me_cuda_ref_create_deblocked(int w, int h, me_imgpel ****host_sub, me_ref_handle *out)
{
  if (int rc = only8("me_cuda_ref_create_deblocked")) return rc;
  g.spec.valid = false;
  if (int rc = check_size(w, h)) return rc;
  if (!g.db_valid || g.db_w != w || g.db_h != h) return fail(1, "me_cuda_ref_create_deblocked: no deblocked %d x %d picture on the device", w, h);
  if (host_sub) return fail(1, "me_cuda_ref_create_deblocked: the plane copy-back is not offered here (page the planes in with me_cuda_ref_download_region)");
  return create_ref_common(g.d_db, w, w, h, g.stream, out);
}

int me_cuda_frame_bipred(const me_bipred_job *jobs, int n_jobs, int refinements, int search_range, int subpel, me_bipred_result *out)
{
  if (int rc = only8("me_cuda_frame_bipred")) return rc;
  if (!g.inited) return fail(1, "me_cuda_init has not been called");
  if (n_jobs <= 0) return 0;
  if (g.list_n[0] < 1 || g.list_n[1] < 1) return fail(1, "me_cuda_frame_bipred needs a reference picture in both lists");
  if (g.view) return fail(1, "me_cuda_frame_bipred: a field view of the current picture is selected (upload the field itself)");
  if (refinements < 0 || refinements > 16 || search_range < 0 || search_range > 64 || subpel < 0 || subpel > 2)
    return fail(1, "me_cuda_frame_bipred: refinements %d / search range %d / sub-pel levels %d out of range", refinements, search_range, subpel);
  for (int i = 0; i < n_jobs; i++) {
    const me_bipred_job &j = jobs[i];
    if (j.list < 0 || j.list > 1) return fail(1, "bi-pred job %d: list %d", i, (int)j.list);
    if (j.opix_x < 0 || j.opix_y < 0 || j.opix_x + 16 > g.p.width || j.opix_y + 16 > g.cur_h || (j.opix_x & 15) || (j.opix_y & 15))
      return fail(1, "bi-pred job %d: macroblock origin (%d,%d) invalid", i, (int)j.opix_x, (int)j.opix_y);
  }
  if (g.bi_cap < n_jobs) {
    if (g.d_bi_jobs) cudaFree(g.d_bi_jobs);
    if (g.d_bi_out) cudaFree(g.d_bi_out);
    g.d_bi_jobs = nullptr; g.d_bi_out = nullptr; g.bi_cap = 0;
    CK(cudaMalloc((void **)&g.d_bi_jobs, sizeof(me_bipred_job) * (size_t)n_jobs));
    CK(cudaMalloc((void **)&g.d_bi_out, sizeof(me_bipred_result) * (size_t)n_jobs));
    g.bi_cap = n_jobs;
  }
  if (!ref_of(g.list[0][0]) || !ref_of(g.list[1][0])) return fail(3, "me_cuda_frame_bipred: stale reference handle in a list");
  CK(cudaMemcpyAsync(g.d_bi_jobs, jobs, sizeof(me_bipred_job) * (size_t)n_jobs, cudaMemcpyHostToDevice, g.stream));
  // the bi-predictive distortions read the pictures' own planes (their weights are applied to the pair of samples), never the
  // weighted copies the uni-predictive searches of the same list entries may be using: a two-entry table of its own
  {
    MeRefDev pair[2 * ME_MAX_REFS];
    memset(pair, 0, sizeof pair);
    pair[0] = ref_of(g.list[0][0])->dev; pair[ME_MAX_REFS] = ref_of(g.list[1][0])->dev;
    if (!g.d_bi_table) CK(cudaMalloc((void **)&g.d_bi_table, sizeof pair));
    CK(cudaMemcpyAsync(g.d_bi_table, pair, sizeof pair, cudaMemcpyHostToDevice, g.stream));
  }
  me_launch_bipred_batch(g.d_bi_jobs, n_jobs, g.d_bi_table, g.d_cur, g.cur_pitch, &g.cfg, refinements, search_range, subpel, g.bw, g.d_bi_out, g.stream);
  g.launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, g.d_bi_out, sizeof(me_bipred_result) * (size_t)n_jobs, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  return 0;
}

}  // extern "C"
