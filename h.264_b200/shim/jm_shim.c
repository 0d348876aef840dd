/*
 * jm_shim.c -- the reference-side binding of libme_cuda.so for the JM 12.4 encoder (lencod).
 *
 * This is what a maintainer adds to lencod to route its motion-estimation hot path through the GPU library.  It is
 * written as a link-time interposer: the UNMODIFIED reference objects are linked with
 *     -Wl,--wrap=FastFullPelBlockMotionSearch,--wrap=FullPelBlockMotionSearch,--wrap=SubPelBlockMotionSearch,--wrap=FullPelBlockMotionBiPred,
 *         --wrap=SubPelBlockSearchBiPred,--wrap=ResetFastFullIntegerSearch,--wrap=getSubImagesLuma,
 *         --wrap=encode_one_frame,--wrap=Init_Motion_Search_Module,--wrap=Clear_Motion_Search_Module,
 *         --wrap=free_storable_picture,--wrap=PartitionMotionSearch,--wrap=BIDPartitionCost,--wrap=GetDirectCost8x8,
 *         --wrap=UMVLine4X,--wrap=FastLine4X,--wrap=DeblockFrame
 * so every call site the SearchMode dispatch of mv-search.c reaches (mv-search.c:745, :823, :960, :998, :1018), the
 * per-macroblock reset (macroblock.c:509), the plane generation (image.c:1642) and the module init / clear
 * (lencod.c:255, :387) land here.  Everything else -- predictors (SetMotionVectorPredictor, mv-search.c:87), mode
 * decision, transforms, entropy coding -- stays the reference's host C.  (INTEGRATION.md shows the same binding as a
 * source patch.)  The library is loaded with dlopen, so the encoder binary has no link-time CUDA dependency; it is
 * compiled against the reference's own headers by oracle/Makefile (target _ref/lencod_cuda).
 *
 * Environment: ME_CUDA_VERIFY=1 runs the reference function next to every GPU call and reports differences (diagnostic),
 *              ME_CUDA_LIB = path of libme_cuda.so (default: <dir of executable>/../../h.264_b200/libme_cuda.so),
 *              ME_CUDA_DISABLE=1 turns the interposer into a pass-through (pure reference behaviour),
 *              ME_CUDA_PMS=0 keeps one device launch per block search (no launch per PartitionMotionSearch call),
 *              ME_CUDA_MC=0 leaves BIDPartitionCost / GetDirectCost8x8 (mv-search.c:1052, :1273) to the reference's host code,
 *              ME_CUDA_DEBLOCK=0 leaves the in-loop deblocking filter (loopFilter.c:87) to the reference's host code,
 *              ME_CUDA_HOST_PLANES=1 copies all 16 planes of every reference picture back into imgY_sub (70 MB per 1080p picture)
 *              instead of paging in what the host code touches, ME_CUDA_POISON_PLANES=1 fills imgY_sub with a pattern first
 *              (test aid: a host read that bypasses the paging then changes the bitstream).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "global.h"
#include "image.h"
#include "mbuffer.h"
#include "me_distortion.h"
#include "refbuf.h"
#include "mb_access.h"
#include "me_umhexsmp.h"

#include "me_cuda.h"

extern const int LEVELMVLIMIT[17][6];
extern void SetMotionVectorPredictor (short pmv[2], char **refPic, short ***tmp_mv, short ref_frame,
                                      int list, int block_x, int block_y, int blockshape_x, int blockshape_y);
extern void getLuma4x4Neighbour (int curr_mb_nr, int block_x_pos, int block_y_pos, PixelPos *pix);

/* ---- the library, resolved at run time ---- */
static struct {
  void *dl;
  int  (*init)(const me_cuda_params *);
  void (*shutdown)(void);
  void (*set_error_handler)(me_cuda_error_fn);
  const char *(*last_error)(void);
  int  (*ref_create)(const me_imgpel *const *, int, int, me_imgpel ****, me_ref_handle *);
  int  (*ref_destroy)(me_ref_handle);
  int  (*set_current)(const me_imgpel *const *, int, int);
  int  (*set_ref_list)(int, int, const me_ref_handle *);
  int  (*set_slice_type)(int);
  int  (*set_ref_weights)(int, int, int, int, int, int);
  int  (*set_bipred_weights)(int, int, int, int, int, int);
  int  (*set_bipred_chroma_weights)(int, int, int, int, int);
  int  (*set_current_chroma)(const me_imgpel *const *, const me_imgpel *const *, int, int);
  int  (*ref_set_chroma)(me_ref_handle, const me_imgpel *const *, const me_imgpel *const *, int, int);
  int  (*set_chroma_state)(int);
  int  (*ref_set_chroma_planes)(me_ref_handle, me_ref_handle, me_ref_handle);
  int  (*mb_begin)(int, int);
  int  (*select_view)(int);
  int  (*ffs_setup)(int, int, int, int);
  int  (*ffs_is_set_up)(int, int);
  int  (*ffs)(short, int, int, int, int, short, short, short *, short *, int, int, int);
  int  (*fullpel)(short, int, int, int, int, short, short, short *, short *, int, int, int);
  int  (*subpel)(short, int, int, int, int, short, short, short *, short *, int, int, int, const int *);
  int  (*bi_full)(short, int, int, int, int, short, short, short, short, short *, short *, short *, short *, int, int, int);
  int  (*bi_sub)(short, int, int, int, int, short, short, short *, short *, short *, short *, int, int, int, const int *);
  int  (*epzs_sub)(short, int, int, int, int, const short *, short *, int, int, int, const int *, int);
  int  (*epzs_bi_sub)(short, int, int, int, int, const short *, const short *, short *, const short *, int, int, int, const int *);
  int  (*umhex_sub)(short, int, int, int, int, short, short, short *, short *, int, int, int, int);
  int  (*smp_sub)(short, int, int, int, int, short, short, short *, short *, int, int, int, int, int, int, int, int);
  int  (*smp_full_sub)(short, int, int, int, int, short, short, short *, short *, int, int, int, int, int, int);
  int  (*pms_begin)(const me_mb_motion *, int, const me_pms_item *, const int *);
  int  (*pms_cached)(int, int);
  int  (*pms_end)(int, int);
  int  (*pms_stats)(int64_t *, int64_t *, int64_t *);
  int  (*pred_cost8x8)(const me_pred_block *, int, int, int32_t *, int32_t *);
  int  (*ref_download_region)(me_ref_handle, int, int, int, int, int, int, me_imgpel *, int);
  int  (*deblock_frame)(me_imgpel *const *, me_imgpel *const *, me_imgpel *const *, int, int, int, int, const me_deblock_mb *);
  int  (*ref_create_deblocked)(int, int, me_imgpel ****, me_ref_handle *);
} L;

static int active = 0;          /* GPU path in use for this run */
static int use_pms = 1;         /* one device launch per PartitionMotionSearch call (ME_CUDA_PMS=0: one per block search) */
static int use_pms_mbaff = 1;   /* ... also for the macroblocks of an MBAFF frame (ME_CUDA_PMS_MBAFF=0: one launch per block search there) */
static int verify = 0;          /* ME_CUDA_VERIFY=1: run the reference function too, report any difference, keep the reference result */
static long n_checked = 0, n_mismatch = 0;
/* calls the device answered, per wrapped function; printed when the module is cleared so that a run (and the tests around it)
 * can tell a GPU run from a silent pass-through */
static long served_ffs = 0, served_fullpel = 0, served_subpel = 0, served_bi_full = 0, served_bi_sub = 0, served_planes = 0;
static long served_field_mbs = 0, served_field_pictures_mbs = 0;   /* macroblocks searched on a field view: MBAFF field macroblocks, field pictures */
static long served_pattern[5] = {0, 0, 0, 0, 0}, served_pattern_bi = 0;   /* sub-pel stages of the pattern searches, by pattern */
static long served_bid_cost = 0, served_direct_cost = 0, paged_tiles = 0, paged_bytes = 0;
static int hbd = 0;             /* more than 8 bits per sample */
static int use_mc = 1;          /* BIDPartitionCost / GetDirectCost8x8 on the device (ME_CUDA_MC=0: the reference's host code) */
static int lazy_planes = 1;     /* imgY_sub paged in on demand (ME_CUDA_HOST_PLANES=1: all planes copied back at creation) */
static int poison_planes = 0;
static int use_deblock = 1;     /* DeblockFrame on the device (ME_CUDA_DEBLOCK=0: the reference's host code) */
static long served_deblock = 0, reused_deblocked = 0;
static imgpel **deblocked_rows = NULL;   /* the luma rows the device filtered last and still holds */
static int deblocked_pic = -1;           /* ... during this encode_one_frame call */
static int pic_counter = 0, pic_uploaded = -1, plane_uploaded = -1;

#define MAX_MAP 64
/* plane: 0, or the colour plane in 4:4:4 independent mode.  sub: the picture's host plane set (StorablePicture::p_curr_img_sub at
 * creation); valid: one flag per (plane, tile of TILE_W x TILE_H samples) -- has the tile been paged in from the device? */
#define TILE_W 64
#define TILE_H 32
static struct { StorablePicture *p; int plane; me_ref_handle h; imgpel ****sub; unsigned char *valid; int ntx, nty, wp, hp; } map[MAX_MAP];
static int n_map = 0;

static void die(const char *what)
{
  static char msg[600];
  snprintf(msg, sizeof msg, "me_cuda shim: %s: %s", what, L.last_error ? L.last_error() : "");
  error(msg, 500);                                      /* filehandle.c:38: prints, flushes the DPB, exits */
}
static void on_library_error(const char *text, int code)
{
  static char msg[600];
  snprintf(msg, sizeof msg, "me_cuda: %s", text);
  error(msg, code);
}

static void *sym(const char *name)
{
  void *p = dlsym(L.dl, name);
  if (!p) { fprintf(stderr, "me_cuda shim: symbol %s missing in libme_cuda.so\n", name); exit(501); }
  return p;
}

static void load_library(void)
{
  char path[PATH_MAX];
  const char *e = getenv("ME_CUDA_LIB");
  if (e && *e) snprintf(path, sizeof path, "%s", e);
  else {
    char exe[PATH_MAX];
    ssize_t n = readlink("/proc/self/exe", exe, sizeof exe - 1);
    char *slash;
    if (n <= 0) { fprintf(stderr, "me_cuda shim: cannot locate the executable\n"); exit(501); }
    exe[n] = 0;
    slash = strrchr(exe, '/');
    if (slash) *slash = 0;
    snprintf(path, sizeof path, "%s/../../h.264_b200/libme_cuda.so", exe);
  }
  L.dl = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!L.dl) { fprintf(stderr, "me_cuda shim: cannot load %s: %s (no CPU fallback: set ME_CUDA_DISABLE=1 for the pure reference)\n", path, dlerror()); exit(501); }
  L.init = sym("me_cuda_init"); L.shutdown = sym("me_cuda_shutdown");
  L.set_error_handler = sym("me_cuda_set_error_handler"); L.last_error = sym("me_cuda_last_error");
  L.ref_create = sym("me_cuda_ref_create"); L.ref_destroy = sym("me_cuda_ref_destroy");
  L.set_current = sym("me_cuda_set_current"); L.set_ref_list = sym("me_cuda_set_ref_list");
  L.set_slice_type = sym("me_cuda_set_slice_type"); L.mb_begin = sym("me_cuda_mb_begin");
  L.select_view = sym("me_cuda_select_current_view");
  L.set_ref_weights = sym("me_cuda_set_ref_weights"); L.set_bipred_weights = sym("me_cuda_set_bipred_weights");
  L.set_bipred_chroma_weights = sym("me_cuda_set_bipred_chroma_weights");
  L.set_current_chroma = sym("me_cuda_set_current_chroma"); L.ref_set_chroma = sym("me_cuda_ref_set_chroma");
  L.set_chroma_state = sym("me_cuda_set_chroma_state"); L.ref_set_chroma_planes = sym("me_cuda_ref_set_chroma_planes");
  L.ffs_setup = sym("me_cuda_ffs_setup"); L.ffs_is_set_up = sym("me_cuda_ffs_is_set_up");
  L.fullpel = sym("me_cuda_FullPelBlockMotionSearch");
  L.ffs = sym("me_cuda_FastFullPelBlockMotionSearch"); L.subpel = sym("me_cuda_SubPelBlockMotionSearch");
  L.bi_full = sym("me_cuda_FullPelBlockMotionBiPred"); L.bi_sub = sym("me_cuda_SubPelBlockSearchBiPred");
  L.epzs_sub = sym("me_cuda_EPZSSubPelBlockMotionSearch"); L.epzs_bi_sub = sym("me_cuda_EPZSSubPelBlockSearchBiPred");
  L.umhex_sub = sym("me_cuda_UMHEXSubPelBlockMotionSearch"); L.smp_sub = sym("me_cuda_smpUMHEXSubPelBlockMotionSearch");
  L.smp_full_sub = sym("me_cuda_smpUMHEXFullSubPelBlockMotionSearch");
  L.pms_begin = sym("me_cuda_pms_begin"); L.pms_cached = sym("me_cuda_pms_cached"); L.pms_end = sym("me_cuda_pms_end");
  L.pms_stats = sym("me_cuda_pms_stats");
  L.pred_cost8x8 = sym("me_cuda_pred_cost8x8"); L.ref_download_region = sym("me_cuda_ref_download_region");
  L.deblock_frame = sym("me_cuda_deblock_frame"); L.ref_create_deblocked = sym("me_cuda_ref_create_deblocked");
}

/* 4:4:4 independent mode (SeparateColourPlane = 1, IS_INDEPENDENT, defines.h:224): the three colour planes are coded one after
 * the other as luma pictures; every reference picture has three sets of quarter-pel planes (UnifiedOneForthPix_JV,
 * image.c:1672-1711, p_img_sub[0..2]) and the searches of plane img->colour_plane_id read the set of that plane. */
static int current_plane(void) { return IS_INDEPENDENT(input) ? img->colour_plane_id : 0; }

static me_ref_handle handle_of(StorablePicture *p)
{
  int i, plane = current_plane();
  for (i = n_map - 1; i >= 0; i--) if (map[i].p == p && map[i].plane == plane) return map[i].h;
  return 0;
}

static void unmap_index(int i)
{
  if (L.ref_destroy(map[i].h)) die("me_cuda_ref_destroy");
  free(map[i].valid);
  memmove(&map[i], &map[i + 1], sizeof(map[0]) * (size_t)(n_map - 1 - i));
  n_map--;
}

/* the frame store at index used_size is the one being filled: insert_picture_in_dpb (mbuffer.c:2548) makes the plane sets of a frame
 * and -- in a stream with field support -- of its two fields (dpb_split_field, :2971) before store_picture_in_dpb counts the store */
static int in_dpb(StorablePicture *p)
{
  unsigned i;
  for (i = 0; i <= dpb.used_size && i < dpb.size; i++)
    if (dpb.fs[i] && (dpb.fs[i]->frame == p || dpb.fs[i]->top_field == p || dpb.fs[i]->bottom_field == p)) return 1;
  return 0;
}

/* ---------------------------------------------------------------- module init / clear (mv-search.c:290, :443) */
void __real_Init_Motion_Search_Module(void);
void __wrap_Init_Motion_Search_Module(void)
{
  const char *dis = getenv("ME_CUDA_DISABLE");
  __real_Init_Motion_Search_Module();
  active = 0;
  if (dis && atoi(dis)) return;
  /* SearchMode 0 (fast full search) and -1 (per-block full search) run every stage on the device.  With the pattern searches
   * (1 UMHexagonS, 2 simplified UMHexagonS, 3 EPZS) the integer-pel stage stays the reference's host code -- a data-dependent walk
   * of a few positions -- and the device serves everything the dispatch calls after it: the plane sets, SubPelBlockMotionSearch
   * (UMHexagonS block types 1..3, mv-search.c:799; EPZS without EPZSSubPelME, :823), SubPelBlockSearchBiPred (:1018), and the
   * modes' own sub-pel stages UMHEXSubPelBlockMotionSearch (:794), smpUMHEX[Full]SubPelBlockMotionSearch (:807, :812),
   * EPZSSubPelBlockMotionSearch (:818) and EPZSSubPelBlockSearchBiPred (:993, :1013) */
  {
    me_cuda_params p;
    int apply_wp = (input->WeightedPrediction || input->WeightedBiprediction) && input->UseWeightedReferenceME;
    memset(&p, 0, sizeof p);
    load_library();
    L.set_error_handler(on_library_error);
    p.abi_version = ME_CUDA_ABI_VERSION;
    p.width = img->width; p.height = img->height;
    /* more than 8 bits per sample (imgpel is u16, global.h:50): plane sets, every search stage and the launch per
     * PartitionMotionSearch and the prediction costs run on 16-bit planes; the deblocking filter stays the reference's host code in
     * such runs; chroma-aware ME is refused */
    p.bit_depth = img->bitdepth_luma;
    hbd = img->bitdepth_luma > 8;
    if (hbd && IS_INDEPENDENT(input) && img->bitdepth_chroma != img->bitdepth_luma) die("4:4:4 independent mode with different luma / chroma bit depths");
    p.search_range = input->search_range;
    p.max_num_references = img->max_num_references;
    p.rdopt = input->rdopt;
    p.full_search = input->full_search;
    p.me_metric[0] = input->MEErrorMetric[F_PEL]; p.me_metric[1] = input->MEErrorMetric[H_PEL]; p.me_metric[2] = input->MEErrorMetric[Q_PEL];
    p.transform8x8 = input->Transform8x8Mode != 0;
    p.mv_limit_min_y = LEVELMVLIMIT[img->LevelIndex][0];
    p.mv_limit_max_y = LEVELMVLIMIT[img->LevelIndex][1];
    p.chroma_me_enable = input->ChromaMEEnable;
    p.chroma_me_weight = input->ChromaMEWeight;
    p.yuv_format = img->yuv_format;
    p.weighted_me = apply_wp;
    p.device = -1;
    if (L.init(&p)) die("me_cuda_init");
    active = 1;
    verify = getenv("ME_CUDA_VERIFY") && atoi(getenv("ME_CUDA_VERIFY"));
    use_pms = !(getenv("ME_CUDA_PMS") && !atoi(getenv("ME_CUDA_PMS")));
    use_pms_mbaff = !(getenv("ME_CUDA_PMS_MBAFF") && !atoi(getenv("ME_CUDA_PMS_MBAFF")));
    use_mc = !(getenv("ME_CUDA_MC") && !atoi(getenv("ME_CUDA_MC")));
    lazy_planes = !(getenv("ME_CUDA_HOST_PLANES") && atoi(getenv("ME_CUDA_HOST_PLANES")));
    poison_planes = getenv("ME_CUDA_POISON_PLANES") && atoi(getenv("ME_CUDA_POISON_PLANES"));
    use_deblock = !(getenv("ME_CUDA_DEBLOCK") && !atoi(getenv("ME_CUDA_DEBLOCK")));
    deblocked_rows = NULL;
  }
}

void __real_Clear_Motion_Search_Module(void);
void __wrap_Clear_Motion_Search_Module(void)
{
  if (active && verify) fprintf(stderr, "me_cuda shim: verify mode: %ld calls checked, %ld mismatches\n", n_checked, n_mismatch);
  if (active) {
    int64_t pl = 0, ps = 0, pm = 0;
    L.pms_stats(&pl, &ps, &pm);
    fprintf(stderr, "me_cuda shim: served on the device: %ld plane sets, %ld ffs, %ld fullpel, %ld subpel, %ld bi_full, %ld bi_sub calls\n",
            served_planes, served_ffs, served_fullpel, served_subpel, served_bi_full, served_bi_sub);
    fprintf(stderr, "me_cuda shim: sub-pel stages of the pattern searches served on the device: %ld EPZS, %ld EPZS bi-pred, %ld UMHEX, %ld smpUMHEX, %ld smpUMHEX full calls\n",
            served_pattern[1], served_pattern_bi, served_pattern[2], served_pattern[3], served_pattern[4]);
    fprintf(stderr, "me_cuda shim: interlace: %ld macroblock visits searched on a field view of an MBAFF frame, %ld on the fields of field pictures\n",
            served_field_mbs, served_field_pictures_mbs);
    fprintf(stderr, "me_cuda shim: partition-search launches %lld, block searches answered from them %lld, argument mismatches %lld\n",
            (long long)pl, (long long)ps, (long long)pm);
    fprintf(stderr, "me_cuda shim: mode-decision costs on the device: %ld BIDPartitionCost, %ld GetDirectCost8x8 calls; plane tiles paged to the host: %ld (%ld bytes%s)\n",
            served_bid_cost, served_direct_cost, paged_tiles, paged_bytes, lazy_planes ? "" : "; all planes copied back at creation");
    fprintf(stderr, "me_cuda shim: pictures deblocked on the device: %ld, reference plane sets built from them without an upload: %ld\n",
            served_deblock, reused_deblocked);
  }
  else
    fprintf(stderr, "me_cuda shim: pass-through (ME_CUDA_DISABLE): nothing served on the device\n");
  if (active) { L.shutdown(); active = 0; n_map = 0; }
  __real_Clear_Motion_Search_Module();
}

/* ---------------------------------------------------------------- pictures */
int __real_encode_one_frame(void);
int __wrap_encode_one_frame(void)
{
  pic_counter++;
  return __real_encode_one_frame();
}

void __real_getSubImagesLuma(StorablePicture *s);
void __wrap_getSubImagesLuma(StorablePicture *s)
{
  me_ref_handle h = 0;
  int i, plane = 0;
  if (!active) { __real_getSubImagesLuma(s); return; }
  if (IS_INDEPENDENT(input)) plane = s->colour_plane_id;      /* set by UnifiedOneForthPix_JV (image.c:1706) right before this call */
  /* 4:4:4 common mode runs this routine on the U and V planes too (image.c:1647-1657).  Those plane sets are the chroma sub-images
   * of ChromaMEEnable; without it they only serve the host's motion compensation and stay on the CPU */
  else if (s->p_curr_img != s->imgY) {
    if (!input->ChromaMEEnable) { __real_getSubImagesLuma(s); return; }
    plane = (s->p_curr_img == s->imgUV[0]) ? 1 : 2;
  }
  /* a picture structure can be reallocated at the same address: drop a stale mapping, and forget pictures that
   * left the decoded picture buffer through a call path that bypasses the wrapped free (same translation unit) */
  for (i = n_map - 1; i >= 0; i--)
    if (map[i].p == s ? map[i].plane == plane : !in_dpb(map[i].p)) unmap_index(i);   /* s itself enters the buffer after this call */
  /* the 16 planes are computed on the device and copied into the picture's plane set for the encoder's host-side consumers */
  /* the 16 planes stay on the device; the host's own consumers (OneComponentLumaPrediction, macroblock.c:816, and whatever host
   * search code a SearchMode keeps) reach them through UMVLine4X / FastLine4X, which page in the tiles they are about to read */
  if (lazy_planes && deblocked_rows && deblocked_pic == pic_counter && s->p_curr_img == deblocked_rows && plane == 0 && s->p_curr_img == s->imgY) {
    /* this is the picture the device has just deblocked (image.c:331 -> store_picture_in_dpb -> UnifiedOneForthPix): its filtered luma
     * is still there */
    if (L.ref_create_deblocked(s->size_x, s->size_y, NULL, &h)) die("me_cuda_ref_create_deblocked");
    reused_deblocked++;
  }
  else if (L.ref_create((const me_imgpel *const *)s->p_curr_img, s->size_x, s->size_y, lazy_planes ? NULL : (me_imgpel ****)s->p_curr_img_sub, &h)) die("me_cuda_ref_create");
  deblocked_rows = NULL;
  if (n_map == MAX_MAP) die("more than 64 reference pictures with device plane sets (decoded picture buffer bookkeeping out of step)");
  served_planes++;
  map[n_map].p = s; map[n_map].plane = plane; map[n_map].h = h; map[n_map].sub = s->p_curr_img_sub;
  map[n_map].wp = s->size_x + 2 * IMG_PAD_SIZE; map[n_map].hp = s->size_y + 2 * IMG_PAD_SIZE;
  map[n_map].ntx = (map[n_map].wp + TILE_W - 1) / TILE_W; map[n_map].nty = (map[n_map].hp + TILE_H - 1) / TILE_H;
  map[n_map].valid = NULL;
  if (lazy_planes) {
    map[n_map].valid = calloc((size_t)16 * map[n_map].ntx * map[n_map].nty, 1);
    if (!map[n_map].valid) die("out of memory (tile flags)");
    if (poison_planes) {
      int fy, fx, y;
      for (fy = 0; fy < 4; fy++) for (fx = 0; fx < 4; fx++) for (y = 0; y < map[n_map].hp; y++)
        memset(s->p_curr_img_sub[fy][fx][y], 0xa5, (size_t)map[n_map].wp * sizeof(imgpel));
    }
    /* the weight estimators read the integer plane of every listed picture directly, not through the line accessors
     * (weighted_prediction.c:84, :280, :427, :623: p_curr_img_sub[0][0]): with weighted prediction that plane comes over whole */
    if (input->WeightedPrediction || input->WeightedBiprediction) {
      if (L.ref_download_region(h, 0, 0, 0, 0, map[n_map].wp, map[n_map].hp, (me_imgpel *)&s->p_curr_img_sub[0][0][0][0],
                                (int)(s->p_curr_img_sub[0][0][1] - s->p_curr_img_sub[0][0][0])))
        die("me_cuda_ref_download_region");
      memset(map[n_map].valid, 1, (size_t)map[n_map].ntx * map[n_map].nty);
      paged_tiles += (long)map[n_map].ntx * map[n_map].nty; paged_bytes += (long)map[n_map].wp * map[n_map].hp * (long)sizeof(imgpel);
    }
  }
  n_map++;
  /* ChromaMEEnable: the searches read the chroma sub-images of this picture (getSubImagesChroma, image.c:1659, still runs on the
   * host for motion compensation); the library evaluates them from the reconstructed chroma planes */
  if (input->ChromaMEEnable && img->yuv_format != YUV444 &&
      L.ref_set_chroma(h, (const me_imgpel *const *)s->imgUV[0], (const me_imgpel *const *)s->imgUV[1], s->size_x_cr, s->size_y_cr))
    die("me_cuda_ref_set_chroma");
  if (input->ChromaMEEnable && img->yuv_format == YUV444 && !IS_INDEPENDENT(input) && plane == 2) {
    me_ref_handle hy = 0, hu = 0;
    for (i = 0; i < n_map; i++) if (map[i].p == s) { if (map[i].plane == 0) hy = map[i].h; else if (map[i].plane == 1) hu = map[i].h; }
    if (!hy || !hu || L.ref_set_chroma_planes(hy, hu, h)) die("me_cuda_ref_set_chroma_planes");
  }
}

void __real_free_storable_picture(StorablePicture *p);
void __wrap_free_storable_picture(StorablePicture *p)
{
  int i;
  if (active && p)
    for (i = n_map - 1; i >= 0; i--)
      if (map[i].p == p) unmap_index(i);
  __real_free_storable_picture(p);
}

/* ---------------------------------------------------------------- per macroblock (macroblock.c:509) */
/* Once per macroblock, before its first search: current picture (first macroblock of a picture), reference lists, slice type,
 * macroblock origin.  With SearchMode = 0 the reference's own per-macroblock reset (macroblock.c:509) triggers it; with
 * SearchMode = -1 there is no such call and the first search of the macroblock does. */
static int mb_synced_pic = -1, mb_synced_mb = -1;
static imgpel **cur_uploaded_rows = NULL, **mb_synced_rows = NULL;   /* an MBAFF pair is visited as frame and as field macroblocks under the same numbers */
static void mb_sync(void)
{
  /* Interlace: the macroblock's original samples are the frame (pCurImg == imgY_org_frm), or the top / bottom field -- row pointers
   * into that frame (lencod.c:1970) -- for the two pictures of field coding (image.c:2983, :3001) and for field macroblocks of an MBAFF
   * frame (macroblock.c:165-190).  The frame goes to the device once per coded picture; each macroblock selects its view. */
  int view = (pCurImg == imgY_org_frm) ? 0 : (input->PicInterlace || input->MbInterlace) ? ((pCurImg == imgY_org_top) ? 1 : (pCurImg == imgY_org_bot) ? 2 : -1) : -1;
  if (IS_INDEPENDENT(input)) view = -1;                  /* the colour planes are uploaded one by one (below) */
  if (mb_synced_pic == pic_counter && mb_synced_mb == img->current_mb_nr && mb_synced_rows == pCurImg) return;
  mb_synced_pic = pic_counter; mb_synced_mb = img->current_mb_nr; mb_synced_rows = pCurImg;
  if (view >= 0) {
    if (pic_uploaded != pic_counter || cur_uploaded_rows != imgY_org_frm) {
      int frame_h = img->height << (img->structure != FRAME);
      if (L.set_current((const me_imgpel *const *)imgY_org_frm, img->width, frame_h)) die("me_cuda_set_current");
      if (input->ChromaMEEnable &&
          L.set_current_chroma((const me_imgpel *const *)imgUV_org_frm[0], (const me_imgpel *const *)imgUV_org_frm[1], img->width_cr,
                               img->height_cr << (img->structure != FRAME)))
        die("me_cuda_set_current_chroma");
      pic_uploaded = pic_counter; plane_uploaded = current_plane(); cur_uploaded_rows = imgY_org_frm;
    }
    if (L.select_view(view)) die("me_cuda_select_current_view");
    if (view) { if (img->structure == FRAME) served_field_mbs++; else served_field_pictures_mbs++; }
  }
  else if (pic_uploaded != pic_counter || plane_uploaded != current_plane() || cur_uploaded_rows != pCurImg) {
    if (L.set_current((const me_imgpel *const *)pCurImg, img->width, img->height)) die("me_cuda_set_current");
    if (input->ChromaMEEnable &&
        L.set_current_chroma((const me_imgpel *const *)imgUV_org[0], (const me_imgpel *const *)imgUV_org[1], img->width_cr, img->height_cr))
      die("me_cuda_set_current_chroma");
    pic_uploaded = pic_counter; plane_uploaded = current_plane(); cur_uploaded_rows = pCurImg;
  }
  {
    int list_offset = img->mb_data[img->current_mb_nr].list_offset, l, i;
    int nlists = (img->type == B_SLICE) ? 2 : 1;
    me_ref_handle hs[32];
    for (l = 0; l < 2; l++) {
      int n = (l < nlists) ? listXsize[l + list_offset] : 0;
      if (n > 32) n = 32;
      for (i = 0; i < n; i++) {
        hs[i] = handle_of(listX[l + list_offset][i]);
        /* a listed picture without device planes would silently shorten the list the searches index: fail loudly instead
         * (the fields of a stored frame have their own plane sets, dpb_split_field; list offsets 2 / 4 of MBAFF field macroblocks
         * select lists made of them) */
        if (!hs[i]) die("a reference-list entry has no plane set on the device");
      }
      if (L.set_ref_list(l, n, hs)) die("me_cuda_set_ref_list");
    }
    L.set_slice_type(img->type == B_SLICE ? ME_B_SLICE : ME_P_SLICE);
    if (L.mb_begin(img->opix_x, img->opix_y)) die("me_cuda_mb_begin");
  }
}

void __real_ResetFastFullIntegerSearch(void);
void __wrap_ResetFastFullIntegerSearch(void)
{
  __real_ResetFastFullIntegerSearch();
  if (!active) return;
  mb_synced_mb = -1;                                     /* a macroblock coded twice (e.g. rate control) is set up twice */
  mb_sync();
}

/* The reference search functions leave values in the globals of refbuf.h / me_distortion.c that later host code relies
 * on: OneComponentLumaPrediction (macroblock.c:816-819) calls UMVLine4X BEFORE it assigns width_pad / height_pad, so its
 * first clamp uses whatever the last motion search stored there (me_fullfast.c:524-525, me_fullsearch.c:380-381,
 * :215-216, :570-571).  A drop-in replacement has to keep those side effects or the very first prediction of a run is
 * fetched with a clamp range of 0 and one mode decision changes. */
static void keep_reference_side_effects(int list, short ref)
{
  int list_offset = img->mb_data[img->current_mb_nr].list_offset;
  StorablePicture *rp = listX[list + list_offset][ref];
  ref_pic_sub.luma = rp->p_curr_img_sub;
  img_width  = rp->size_x;
  img_height = rp->size_y;
  width_pad  = rp->size_x_pad;
  height_pad = rp->size_y_pad;
  /* the global ChromaMEEnable (me_distortion.c:56) as the reference function about to be replaced would see it:
   * BlockMotionSearch changes it between the integer-pel and the later stages (mv-search.c:612, :779) */
  if (input->ChromaMEEnable && L.set_chroma_state(ChromaMEEnable)) die("me_cuda_set_chroma_state");
}

/* Weighted-prediction ME: what SetupFastFullPelSearch (me_fullfast.c:507-508, :530-534) and SubPelBlockMotionSearch
 * (me_fullsearch.c:368-388) read from the picture parameter set and the wp_* tables before a uni-predictive search. */
static void pass_weights(int list, short ref)
{
  int list_offset = img->mb_data[img->current_mb_nr].list_offset;
  int apply = ((active_pps->weighted_pred_flag && (img->type == P_SLICE || img->type == SP_SLICE)) ||
               (active_pps->weighted_bipred_idc && (img->type == B_SLICE))) && input->UseWeightedReferenceME;
  /* the library does not weight the chroma samples of uni-predictive searches: the encoder's estimators never produce
   * chroma weights other than the default (weighted_prediction.c:98-103, 308-313); anything else is refused loudly */
  if (apply && input->ChromaMEEnable &&
      (wp_weight[list + list_offset][ref][1] != (1 << chroma_log_weight_denom) || wp_weight[list + list_offset][ref][2] != (1 << chroma_log_weight_denom) ||
       wp_offset[list + list_offset][ref][1] || wp_offset[list + list_offset][ref][2]))
    die("weighted chroma samples in a uni-predictive search (ChromaMEEnable with non-default weight_cr / offset_cr) are not implemented");
  if (L.set_ref_weights(list, ref, apply, apply ? wp_weight[list + list_offset][ref][0] : 0,
                        apply ? wp_offset[list + list_offset][ref][0] : 0, apply ? luma_log_weight_denom : 0))
    die("me_cuda_set_ref_weights");
}

/* Weighted bi-prediction: the weights the two bi-predictive searches derive from their arguments -- FullPelBlockMotionBiPred
 * at me_fullsearch.c:202-227, SubPelBlockSearchBiPred at :541-582 (their wp_offset index expressions differ for list 1). */
static void pass_bipred_weights(int list, short ref, int subpel)
{
  int list_offset = img->mb_data[img->current_mb_nr].list_offset;
  if (active_pps->weighted_bipred_idc > 0) {
    int o1, o2, w1, w2;
    if (subpel) {
      o1 = list == 0 ? wp_offset[list_offset][ref][0] : wp_offset[list_offset + 1][0][0];
      o2 = list == 0 ? wp_offset[list_offset + 1][ref][0] : wp_offset[list_offset][0][0];
    } else {
      o1 = list == 0 ? wp_offset[list_offset][ref][0] : wp_offset[list_offset + 1][0][ref];
      o2 = list == 0 ? wp_offset[list_offset + 1][ref][0] : wp_offset[list_offset][0][ref];
    }
    w1 = list == 0 ? wbp_weight[list_offset][ref][0][0] : wbp_weight[list_offset + LIST_1][0][ref][0];
    w2 = list == 0 ? wbp_weight[list_offset + LIST_1][ref][0][0] : wbp_weight[list_offset][0][ref][0];
    if (L.set_bipred_weights(list, 1, w1, w2, (o1 + o2 + 1) >> 1, luma_log_weight_denom)) die("me_cuda_set_bipred_weights");
    if (input->ChromaMEEnable) {
      /* weight1_cr / weight2_cr / offsetBi_cr, me_fullsearch.c:245-262 (the same expressions in both functions) */
      int k;
      for (k = 0; k < 2; k++) {
        int c1 = list == 0 ? wbp_weight[list_offset][ref][0][k + 1] : wbp_weight[list_offset + LIST_1][0][ref][k + 1];
        int c2 = list == 0 ? wbp_weight[list_offset + LIST_1][ref][0][k + 1] : wbp_weight[list_offset][0][ref][k + 1];
        int ob = list == 0 ? (wp_offset[list_offset][ref][k + 1] + wp_offset[list_offset + LIST_1][ref][k + 1] + 1) >> 1
                           : (wp_offset[list_offset + LIST_1][0][k + 1] + wp_offset[list_offset][0][k + 1] + 1) >> 1;
        if (L.set_bipred_chroma_weights(list, k, c1, c2, ob)) die("me_cuda_set_bipred_chroma_weights");
      }
    }
  } else if (L.set_bipred_weights(list, 0, 0, 0, 0, 0)) die("me_cuda_set_bipred_weights");
}

/* ---------------------------------------------------------------- one launch per PartitionMotionSearch (mv-search.c:1378) */
/* PartitionMotionSearch(blocktype, block8x8) is a loop over lists, references and sub-blocks around BlockMotionSearch; every
 * block search in it would be one device round trip.  Instead, on entry, the motion state SetMotionVectorPredictor reads
 * (enc_picture->ref_idx / ->mv around and inside the macroblock) goes to the device and one launch runs the whole call -- and,
 * speculatively, the calls that will read the same state: the first 16x8 and 8x16 partitions together with the 16x16 block (all
 * their neighbours lie outside the macroblock), and block types 5..7 of an 8x8 block together with type 4 (mode_decision.c:
 * 575-674 searches them one after the other without touching the state outside the block).  The reference function then runs
 * unchanged; its block searches are answered from the launch when their arguments equal what the device used (checked in the
 * library), so a wrong guess costs time, never a different result. */
/* one cell of the motion state from the neighbour at sample (x, y) relative to the macroblock.  MBAFF frames: reference index and
 * vertical vector component of a neighbour in another macroblock are converted between frame and field units as
 * SetMotionVectorPredictor does when it reads them (mv-search.c:140-172, :200-236) -- the conversion depends only on the field
 * flags of the two macroblocks, both known now */
static void snapshot_cell(int x, int y, int ok, int nlists, int8_t *ref_idx[2], int16_t *mv[2])
{
  PixelPos pix;
  int l, cur_field = img->MbaffFrameFlag && img->mb_data[img->current_mb_nr].mb_field;
  if (ok) { getLuma4x4Neighbour(img->current_mb_nr, x, y, &pix); ok = pix.available; }
  for (l = 0; l < 2; l++) {
    int ref, mvy;
    if (!ok || l >= nlists) { *ref_idx[l] = ME_PMS_UNAVAILABLE; continue; }
    ref = enc_picture->ref_idx[l][pix.pos_y][pix.pos_x];
    mvy = enc_picture->mv[l][pix.pos_y][pix.pos_x][1];
    if (img->MbaffFrameFlag && pix.mb_addr != img->current_mb_nr) {
      int nb_field = img->mb_data[pix.mb_addr].mb_field;
      if (cur_field && !nb_field) { ref = ref * 2; mvy = mvy / 2; }
      else if (!cur_field && nb_field) { ref = ref >> 1; mvy = mvy * 2; }
    }
    *ref_idx[l] = (int8_t)ref;
    mv[l][0] = enc_picture->mv[l][pix.pos_y][pix.pos_x][0];
    mv[l][1] = (int16_t)mvy;
  }
}

static void snapshot_motion(me_mb_motion *st)
{
  int r, c, k, nlists = (img->type == B_SLICE) ? 2 : 1;
  int mbaff = img->MbaffFrameFlag != 0;
  memset(st, 0, sizeof *st);
  st->mbaff = (int8_t)mbaff;
  for (r = -1; r < 4; r++)
    for (c = -1; c < 5; c++) {
      int8_t *ri[2] = { &st->ref_idx[0][r + 1][c + 1], &st->ref_idx[1][r + 1][c + 1] };
      int16_t *mv[2] = { st->mv[0][r + 1][c + 1], st->mv[1][r + 1][c + 1] };
      int ok = !(c == 4 && r >= 0);                       /* right of the macroblock: not available (mb_access.c:143) */
      /* MBAFF: exactly the sample positions SetMotionVectorPredictor asks getLuma4x4Neighbour for (left column x = -1, row above y = -1) */
      snapshot_cell(mbaff ? (c < 0 ? -1 : 4 * c) : 4 * c, mbaff ? (r < 0 ? -1 : 4 * r) : 4 * r, ok, nlists, ri, mv);
    }
  for (k = 0; k < 4; k++) {                                /* neighbour D of the blocks in column 0: sample (-1, 4k - 1) */
    int8_t *ri[2] = { &st->d_left_ref_idx[0][k], &st->d_left_ref_idx[1][k] };
    int16_t *mv[2] = { st->d_left_mv[0][k], st->d_left_mv[1][k] };
    snapshot_cell(-1, 4 * k - 1, mbaff, nlists, ri, mv);
  }
}

void __real_PartitionMotionSearch(int blocktype, int block8x8, int *lambda_factor);
void __wrap_PartitionMotionSearch(int blocktype, int block8x8, int *lambda_factor)
{
  /* frame pictures, field pictures and MBAFF frames (the motion state goes over with the frame / field conversions of
   * SetMotionVectorPredictor already applied, snapshot_cell) */
  int on = active && use_pms && input->SearchMode == FAST_FULL_SEARCH && !input->DisableSubpelME &&
           blocktype >= 1 && blocktype <= 7 && (use_pms_mbaff || !img->MbaffFrameFlag);
  if (on && !L.pms_cached(blocktype, block8x8)) {
    me_mb_motion st;
    me_pms_item items[ME_PMS_MAX_ITEMS];
    int n = 0, l, r, m, bslice = (img->type == B_SLICE), nlists = bslice ? 2 : 1;
    int list_offset = img->mb_data[img->current_mb_nr].list_offset;
    mb_sync();
    /* SetupFastFullPelSearch of every (list, ref) the call will search (me_fullfast.c:550: the 16x16 predictor, whose neighbours
     * lie outside the macroblock, is the same whenever it is taken during this macroblock) */
    if (input->ChromaMEEnable && L.set_chroma_state(input->ChromaMEEnable)) die("me_cuda_set_chroma_state");
    for (l = 0; l < nlists; l++)
      for (r = 0; r < listXsize[l + list_offset]; r++) {
        pass_weights(l, (short)r);
        if (!L.ffs_is_set_up(l, r)) {
          short pmv[2];
          SetMotionVectorPredictor(pmv, enc_picture->ref_idx[l], enc_picture->mv[l], (short)r, l, 0, 0, 16, 16);
          if (L.ffs_setup(l, r, pmv[0], pmv[1])) die("me_cuda_ffs_setup");
        }
      }
    items[n].blocktype = (int8_t)blocktype; items[n].block8x8 = (int8_t)block8x8; n++;
    if (input->CtxAdptLagrangeMult == 0) {               /* same lambda factors for the calls that follow (md_high.c:115, mode_decision.c:663-672) */
      if (blocktype == 1) {
        for (m = 2; m <= 3; m++)
          if (input->InterSearch[bslice][m]) { items[n].blocktype = (int8_t)m; items[n].block8x8 = 0; n++; }
      } else if (blocktype == 4 && !img->mb_data[img->current_mb_nr].luma_transform_size_8x8_flag && input->Transform8x8Mode != 2) {
        for (m = 5; m <= 7; m++)
          if (input->InterSearch[bslice][m]) { items[n].blocktype = (int8_t)m; items[n].block8x8 = (int8_t)block8x8; n++; }
      }
    }
    snapshot_motion(&st);
    r = L.pms_begin(&st, n, items, lambda_factor);
    if (r > 0) die("me_cuda_pms_begin");                 /* r < 0: configuration not covered, the block searches go one by one */
  }
  __real_PartitionMotionSearch(blocktype, block8x8, lambda_factor);
  if (on) L.pms_end(blocktype, block8x8);
}

/* ---------------------------------------------------------------- block searches */
int __real_FastFullPelBlockMotionSearch(imgpel *, short, int, int, int, int, short, short, short *, short *, int, int, int);
int __wrap_FastFullPelBlockMotionSearch(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                        short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y,
                                        int search_range, int min_mcost, int lambda_factor)
{
  if (!active)
    return __real_FastFullPelBlockMotionSearch(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y,
                                               mv_x, mv_y, search_range, min_mcost, lambda_factor);
  if (!L.ffs_is_set_up(list, ref)) {
    /* SetupFastFullPelSearch (me_fullfast.c:550): the search centre is the 16x16 predictor, computed by host code */
    short pmv[2];
    keep_reference_side_effects(list, ref);
    pass_weights(list, ref);
    SetMotionVectorPredictor(pmv, enc_picture->ref_idx[list], enc_picture->mv[list], ref, list, 0, 0, 16, 16);
    if (L.ffs_setup(list, ref, pmv[0], pmv[1])) die("me_cuda_ffs_setup");
  }
  served_ffs++;
  if (verify) {
    short gx = *mv_x, gy = *mv_y;
    int gc = L.ffs(ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y, &gx, &gy, search_range, min_mcost, lambda_factor);
    int rc = __real_FastFullPelBlockMotionSearch(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y,
                                                 mv_x, mv_y, search_range, min_mcost, lambda_factor);
    n_checked++;
    if (gc != rc || gx != *mv_x || gy != *mv_y) {
      n_mismatch++;
      fprintf(stderr, "VERIFY ffs pic %d mb (%d,%d) list %d ref %d bt %d blk (%d,%d) pmv (%d,%d) sr %d lam %d: gpu (%d,%d,%d) ref (%d,%d,%d)\n",
              pic_counter, img->opix_x, img->opix_y, list, ref, blocktype, pic_pix_x, pic_pix_y, pred_mv_x, pred_mv_y, search_range,
              lambda_factor, gx, gy, gc, *mv_x, *mv_y, rc);
    }
    return rc;
  }
  return L.ffs(ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y, mv_x, mv_y, search_range, min_mcost, lambda_factor);
}

int __real_FullPelBlockMotionSearch(imgpel *, short, int, int, int, int, short, short, short *, short *, int, int, int);
int __wrap_FullPelBlockMotionSearch(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                    short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y,
                                    int search_range, int min_mcost, int lambda_factor)
{
  if (!active)
    return __real_FullPelBlockMotionSearch(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y,
                                           mv_x, mv_y, search_range, min_mcost, lambda_factor);
  mb_sync();
  keep_reference_side_effects(list, ref);
  pass_weights(list, ref);
  served_fullpel++;
  if (verify) {
    short gx = *mv_x, gy = *mv_y;
    int gc = L.fullpel(ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y, &gx, &gy, search_range, min_mcost, lambda_factor);
    int rc = __real_FullPelBlockMotionSearch(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y,
                                             mv_x, mv_y, search_range, min_mcost, lambda_factor);
    n_checked++;
    if (gc != rc || gx != *mv_x || gy != *mv_y) {
      n_mismatch++;
      fprintf(stderr, "VERIFY fullpel pic %d mb (%d,%d) list %d ref %d bt %d blk (%d,%d) pmv (%d,%d) sr %d min %d: gpu (%d,%d,%d) ref (%d,%d,%d)\n",
              pic_counter, img->opix_x, img->opix_y, list, ref, blocktype, pic_pix_x, pic_pix_y, pred_mv_x, pred_mv_y, search_range,
              min_mcost, gx, gy, gc, *mv_x, *mv_y, rc);
    }
    return rc;
  }
  return L.fullpel(ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y, mv_x, mv_y, search_range, min_mcost, lambda_factor);
}

int __real_SubPelBlockMotionSearch(imgpel *, short, int, int, int, int, short, short, short *, short *, int, int, int, int *);
int __wrap_SubPelBlockMotionSearch(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                   short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y,
                                   int search_pos2, int search_pos4, int min_mcost, int *lambda_factor)
{
  if (!active)
    return __real_SubPelBlockMotionSearch(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y,
                                          mv_x, mv_y, search_pos2, search_pos4, min_mcost, lambda_factor);
  mb_sync();
  keep_reference_side_effects(list, ref);
  pass_weights(list, ref);
  served_subpel++;
  if (verify) {
    short ix = *mv_x, iy = *mv_y, gx = *mv_x, gy = *mv_y;
    int gc = L.subpel(ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y, &gx, &gy, search_pos2, search_pos4, min_mcost, lambda_factor);
    int rc = __real_SubPelBlockMotionSearch(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y,
                                            mv_x, mv_y, search_pos2, search_pos4, min_mcost, lambda_factor);
    n_checked++;
    if (gc != rc || gx != *mv_x || gy != *mv_y) {
      n_mismatch++;
      fprintf(stderr, "VERIFY subpel pic %d mb (%d,%d) list %d ref %d bt %d blk (%d,%d) pmv (%d,%d) in (%d,%d) min %d: gpu (%d,%d,%d) ref (%d,%d,%d)\n",
              pic_counter, img->opix_x, img->opix_y, list, ref, blocktype, pic_pix_x, pic_pix_y, pred_mv_x, pred_mv_y, ix, iy, min_mcost,
              gx, gy, gc, *mv_x, *mv_y, rc);
    }
    return rc;
  }
  return L.subpel(ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y, mv_x, mv_y, search_pos2, search_pos4,
                  min_mcost, lambda_factor);
}

int __real_FullPelBlockMotionBiPred(imgpel *, short, int, int, int, int, short, short, short, short,
                                    short *, short *, short *, short *, int, int, int);
int __wrap_FullPelBlockMotionBiPred(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                    short pred_mv_x1, short pred_mv_y1, short pred_mv_x2, short pred_mv_y2,
                                    short *mv_x, short *mv_y, short *s_mv_x, short *s_mv_y,
                                    int search_range, int min_mcost, int lambda_factor)
{
  if (!active)
    return __real_FullPelBlockMotionBiPred(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x1, pred_mv_y1,
                                           pred_mv_x2, pred_mv_y2, mv_x, mv_y, s_mv_x, s_mv_y, search_range, min_mcost, lambda_factor);
  mb_sync();
  keep_reference_side_effects(list, ref);
  pass_bipred_weights(list, ref, 0);
  served_bi_full++;
  return L.bi_full(ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x1, pred_mv_y1, pred_mv_x2, pred_mv_y2,
                   mv_x, mv_y, s_mv_x, s_mv_y, search_range, min_mcost, lambda_factor);
}

int __real_SubPelBlockSearchBiPred(imgpel *, short, int, int, int, int, short, short, short *, short *, short *, short *,
                                   int, int, int, int *);
int __wrap_SubPelBlockSearchBiPred(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                   short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y, short *s_mv_x, short *s_mv_y,
                                   int search_pos2, int search_pos4, int min_mcost, int *lambda_factor)
{
  if (!active)
    return __real_SubPelBlockSearchBiPred(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y,
                                          mv_x, mv_y, s_mv_x, s_mv_y, search_pos2, search_pos4, min_mcost, lambda_factor);
  mb_sync();
  keep_reference_side_effects(list, ref);
  pass_bipred_weights(list, ref, 1);
  served_bi_sub++;
  return L.bi_sub(ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y, mv_x, mv_y, s_mv_x, s_mv_y,
                  search_pos2, search_pos4, min_mcost, lambda_factor);
}

/* ---------------------------------------------------------------- sub-pel stages of the pattern searches (SearchMode 1, 2, 3) */
/* subthres[blocktype] as EPZSInit derives it (me_epzs.c:335-354); the array itself is a file static of me_epzs.c */
static int epzs_subthres(int blocktype)
{
  static const int medthres_base[8] = { 0, 256, 128, 128, 64, 32, 32, 16 };       /* me_epzs.c:41 */
  int pel_error_me = 1 << (img->bitdepth_luma - 8), pel_error_me_cr = 1 << (img->bitdepth_chroma - 8);
  double chroma_weight = input->ChromaMEEnable ? pel_error_me_cr * input->ChromaMEWeight * (double)(img->width_cr * img->height_cr) / (double)(img->width * img->height) : 0;
  return input->EPZSSubPelThresScale * (medthres_base[blocktype] * pel_error_me + (int)(medthres_base[blocktype] * chroma_weight + 0.5));
}

static void pattern_verify(const char *what, int list, short ref, int blocktype, int pic_pix_x, int pic_pix_y, int ix, int iy, int min_mcost,
                           int gx, int gy, int gc, int rx, int ry, int rc)
{
  n_checked++;
  if (gc != rc || gx != rx || gy != ry) {
    n_mismatch++;
    fprintf(stderr, "VERIFY %s pic %d mb (%d,%d) list %d ref %d bt %d blk (%d,%d) in (%d,%d) min %d: gpu (%d,%d,%d) ref (%d,%d,%d)\n",
            what, pic_counter, img->opix_x, img->opix_y, list, ref, blocktype, pic_pix_x, pic_pix_y, ix, iy, min_mcost, gx, gy, gc, rx, ry, rc);
  }
}

int __real_EPZSSubPelBlockMotionSearch(imgpel *, short, int, int, int, int, short *, short *, int, int, int, int *);
int __wrap_EPZSSubPelBlockMotionSearch(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                       short pred_mv[2], short mv[2], int search_pos2, int search_pos4, int min_mcost, int *lambda)
{
  if (!active)
    return __real_EPZSSubPelBlockMotionSearch(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv, mv, search_pos2, search_pos4, min_mcost, lambda);
  mb_sync();
  keep_reference_side_effects(list, ref);
  pass_weights(list, ref);
  served_pattern[1]++;
  if (verify) {
    short g[2] = { mv[0], mv[1] }, in[2] = { mv[0], mv[1] };
    int gc = L.epzs_sub(ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv, g, search_pos2, search_pos4, min_mcost, lambda, epzs_subthres(blocktype));
    int rc = __real_EPZSSubPelBlockMotionSearch(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv, mv, search_pos2, search_pos4, min_mcost, lambda);
    pattern_verify("epzs_subpel", list, ref, blocktype, pic_pix_x, pic_pix_y, in[0], in[1], min_mcost, g[0], g[1], gc, mv[0], mv[1], rc);
    return rc;
  }
  return L.epzs_sub(ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv, mv, search_pos2, search_pos4, min_mcost, lambda, epzs_subthres(blocktype));
}

int __real_EPZSSubPelBlockSearchBiPred(imgpel *, short, int, int, int, int, short *, short *, short *, short *, int, int, int, int *);
int __wrap_EPZSSubPelBlockSearchBiPred(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                       short *pred_mv1, short *pred_mv2, short mv[2], short s_mv[2],
                                       int search_pos2, int search_pos4, int min_mcost, int *lambda)
{
  if (!active)
    return __real_EPZSSubPelBlockSearchBiPred(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv1, pred_mv2, mv, s_mv,
                                              search_pos2, search_pos4, min_mcost, lambda);
  mb_sync();
  keep_reference_side_effects(list, ref);
  pass_bipred_weights(list, ref, 1);                      /* the index expressions of me_epzs.c:2746-2747 are those of SubPelBlockSearchBiPred */
  served_pattern_bi++;
  if (verify) {
    short g[2] = { mv[0], mv[1] }, in[2] = { mv[0], mv[1] };
    int gc = L.epzs_bi_sub(ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv1, pred_mv2, g, s_mv, search_pos2, search_pos4, min_mcost, lambda);
    int rc = __real_EPZSSubPelBlockSearchBiPred(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv1, pred_mv2, mv, s_mv,
                                                search_pos2, search_pos4, min_mcost, lambda);
    pattern_verify("epzs_bi_subpel", list, ref, blocktype, pic_pix_x, pic_pix_y, in[0], in[1], min_mcost, g[0], g[1], gc, mv[0], mv[1], rc);
    return rc;
  }
  return L.epzs_bi_sub(ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv1, pred_mv2, mv, s_mv, search_pos2, search_pos4, min_mcost, lambda);
}

/* the three UMHexagonS sub-pel stages share a signature (me_umhex.h, me_umhexsmp.h); which = 2 UMHEX, 3 smpUMHEX, 4 smpUMHEX full */
static int umhex_device(int which, short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype, short pred_mv_x, short pred_mv_y,
                        short *mv_x, short *mv_y, int search_pos2, int search_pos4, int min_mcost, int lambda_factor)
{
  if (which == 2)
    return L.umhex_sub(ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y, mv_x, mv_y, search_pos2, search_pos4, min_mcost, lambda_factor);
  if (which == 3)
    return L.smp_sub(ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y, mv_x, mv_y, search_pos2, search_pos4, min_mcost, lambda_factor,
                     smpUMHEX_pred_MV_uplayer_X, smpUMHEX_pred_MV_uplayer_Y, SubPelThreshold1, SubPelThreshold3);
  return L.smp_full_sub(ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y, mv_x, mv_y, search_pos2, search_pos4, min_mcost, lambda_factor,
                        SubPelThreshold1, SubPelThreshold3);
}

#define UMHEX_WRAP(NAME, WHICH, LABEL)                                                                                              \
int __real_##NAME(imgpel *, short, int, int, int, int, short, short, short *, short *, int, int, int, int);                         \
int __wrap_##NAME(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype, short pred_mv_x, short pred_mv_y, \
                  short *mv_x, short *mv_y, int search_pos2, int search_pos4, int min_mcost, int lambda_factor)                     \
{                                                                                                                                   \
  if (!active)                                                                                                                      \
    return __real_##NAME(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y, mv_x, mv_y, search_pos2,       \
                         search_pos4, min_mcost, lambda_factor);                                                                    \
  mb_sync();                                                                                                                        \
  keep_reference_side_effects(list, ref);                                                                                           \
  pass_weights(list, ref);                                                                                                          \
  served_pattern[WHICH]++;                                                                                                          \
  if (verify) {                                                                                                                     \
    short ix = *mv_x, iy = *mv_y, gx = *mv_x, gy = *mv_y;                                                                           \
    int gc = umhex_device(WHICH, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y, &gx, &gy, search_pos2,           \
                          search_pos4, min_mcost, lambda_factor);                                                                   \
    int rc = __real_##NAME(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y, mv_x, mv_y, search_pos2,     \
                           search_pos4, min_mcost, lambda_factor);                                                                  \
    pattern_verify(LABEL, list, ref, blocktype, pic_pix_x, pic_pix_y, ix, iy, min_mcost, gx, gy, gc, *mv_x, *mv_y, rc);             \
    return rc;                                                                                                                      \
  }                                                                                                                                 \
  return umhex_device(WHICH, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y, mv_x, mv_y, search_pos2,             \
                      search_pos4, min_mcost, lambda_factor);                                                                       \
}
UMHEX_WRAP(UMHEXSubPelBlockMotionSearch, 2, "umhex_subpel")
UMHEX_WRAP(smpUMHEXSubPelBlockMotionSearch, 3, "smpumhex_subpel")
UMHEX_WRAP(smpUMHEXFullSubPelBlockMotionSearch, 4, "smpumhex_full_subpel")

/* ---------------------------------------------------------------- the host's reads of the quarter-pel planes (refbuf.c:25, :37) */
/* FastLine4X / UMVLine4X return a pointer into plane Pic[y & 3][x & 3]; every caller then reads at most 16 rows of at most 16
 * samples from it (OneComponentLumaPrediction, macroblock.c:821-826; the distortion functions of me_distortion.c).  The planes
 * live on the device: the tiles that rectangle touches are copied into the picture's host plane set first, once. */
static void page_in(imgpel ****Pic, int fy, int fx, int ypos, int xpos)
{
  int i, tx, ty, x1, y1;
  if (!active || !lazy_planes) return;
  for (i = n_map - 1; i >= 0; i--) if (map[i].sub == Pic) break;
  if (i < 0) return;                                   /* a plane set the host built itself (4:4:4 chroma without ChromaMEEnable) */
  if (xpos < 0) xpos = 0;
  if (ypos < 0) ypos = 0;
  x1 = xpos + 15 < map[i].wp - 1 ? xpos + 15 : map[i].wp - 1;
  y1 = ypos + 15 < map[i].hp - 1 ? ypos + 15 : map[i].hp - 1;
  for (ty = ypos / TILE_H; ty <= y1 / TILE_H; ty++)
    for (tx = xpos / TILE_W; tx <= x1 / TILE_W; tx++) {
      unsigned char *v = &map[i].valid[((size_t)(fy * 4 + fx) * map[i].nty + ty) * map[i].ntx + tx];
      int x0 = tx * TILE_W, y0 = ty * TILE_H, w, h;
      if (*v) continue;
      w = map[i].wp - x0 < TILE_W ? map[i].wp - x0 : TILE_W;
      h = map[i].hp - y0 < TILE_H ? map[i].hp - y0 : TILE_H;
      /* rows of a plane are contiguous (get_mem2Dpel, memalloc.c:69-82): row pitch = padded width */
      if (L.ref_download_region(map[i].h, fy, fx, x0, y0, w, h, (me_imgpel *)&Pic[fy][fx][y0][x0], (int)(Pic[fy][fx][1] - Pic[fy][fx][0])))
        die("me_cuda_ref_download_region");
      *v = 1;
      paged_tiles++; paged_bytes += (long)w * h * (long)sizeof(imgpel);
    }
}

imgpel *__real_FastLine4X(imgpel ****Pic, int y, int x);
imgpel *__wrap_FastLine4X(imgpel ****Pic, int y, int x)
{
  page_in(Pic, y & 3, x & 3, y >> 2, x >> 2);
  return __real_FastLine4X(Pic, y, x);
}

imgpel *__real_UMVLine4X(imgpel ****Pic, int y, int x);
imgpel *__wrap_UMVLine4X(imgpel ****Pic, int y, int x)
{
  page_in(Pic, y & 3, x & 3, iClip3(0, height_pad, y >> 2), iClip3(0, width_pad, x >> 2));     /* the clamp of refbuf.c:39-40 */
  return __real_UMVLine4X(Pic, y, x);
}

/* ---------------------------------------------------------------- mode-decision costs (mv-search.c:1052, :1273) */
/* What LumaPrediction (macroblock.c:832-930) would do for the 4x4 block (h, v) of the macroblock, as a record of the library:
 * direction, references, the vectors it picks (all_mv, or the bi-predictive search results when currMB->bi_pred_me selects
 * them, macroblock.c:863-864) and the weights of macroblock.c:876-913. */
static void fill_pred_block(me_pred_block *b, Macroblock *currMB, int h, int v, int p_dir, int l0_mode, int l1_mode, short l0_ref, short l1_ref)
{
  short ****mv_array = img->all_mv[v][h];
  int apply = ((active_pps->weighted_pred_flag && (img->type == P_SLICE || img->type == SP_SLICE)) ||
               (active_pps->weighted_bipred_idc && (img->type == B_SLICE)));
  memset(b, 0, sizeof *b);
  b->pix_x = (int16_t)(img->opix_x + 4 * h); b->pix_y = (int16_t)(img->opix_y + 4 * v);
  b->p_dir = (int8_t)p_dir;
  if (p_dir < 0) return;
  if (currMB->bi_pred_me && l0_ref == 0 && l1_ref == 0 && p_dir == 2 && l0_mode == 1 && l1_mode == 1)
    mv_array = currMB->bi_pred_me == 1 ? img->bipred_mv1[v][h] : img->bipred_mv2[v][h];
  b->ref[0] = (int8_t)l0_ref; b->ref[1] = (int8_t)l1_ref;
  if (p_dir == 0 || p_dir == 2) { b->mv[0][0] = mv_array[LIST_0][l0_ref][l0_mode][0]; b->mv[0][1] = mv_array[LIST_0][l0_ref][l0_mode][1]; }
  if (p_dir == 1 || p_dir == 2) { b->mv[1][0] = mv_array[LIST_1][l1_ref][l1_mode][0]; b->mv[1][1] = mv_array[LIST_1][l1_ref][l1_mode][1]; }
  b->apply_weights = (int8_t)apply;
  if (apply) {
    b->log_denom = (int16_t)luma_log_weight_denom;
    if (p_dir == 2) {
      b->weight[0] = (int16_t)wbp_weight[0][l0_ref][l1_ref][0]; b->weight[1] = (int16_t)wbp_weight[1][l0_ref][l1_ref][0];
      b->offset = (int16_t)((wp_offset[0][l0_ref][0] + wp_offset[1][l1_ref][0] + 1) >> 1);
    } else if (p_dir == 0) { b->weight[0] = (int16_t)wp_weight[0][l0_ref][0]; b->offset = (int16_t)wp_offset[0][l0_ref][0]; }
    else { b->weight[0] = (int16_t)wp_weight[1][l1_ref][0]; b->offset = (int16_t)wp_offset[1][l1_ref][0]; }
  }
}

/* OneComponentLumaPrediction leaves the clamp range of the last picture it read in width_pad / height_pad (macroblock.c:818-819) */
static void keep_prediction_side_effects(Macroblock *currMB, int p_dir, short l0_ref, short l1_ref)
{
  StorablePicture *rp = NULL;
  if (p_dir == 0 || p_dir == 2) rp = listX[0 + currMB->list_offset][l0_ref];
  if (p_dir == 1 || p_dir == 2) rp = listX[1 + currMB->list_offset][l1_ref];
  if (rp) { width_pad = rp->size_x_pad; height_pad = rp->size_y_pad; }
}

extern int *mvbits;

int __real_BIDPartitionCost(Macroblock *currMB, int blocktype, int block8x8, short ref_l0, short ref_l1, int lambda_factor);
int __wrap_BIDPartitionCost(Macroblock *currMB, int blocktype, int block8x8, short ref_l0, short ref_l1, int lambda_factor)
{
  static const int bx0[5][4] = {{0,0,0,0}, {0,0,0,0}, {0,0,0,0}, {0,2,0,0}, {0,2,0,2}};
  static const int by0[5][4] = {{0,0,0,0}, {0,0,0,0}, {0,2,0,0}, {0,0,0,0}, {0,0,2,2}};
  me_pred_block blocks[16];
  int32_t c4[4], c8[4];
  int parttype = blocktype < 4 ? blocktype : 4;
  int step_h0 = input->part_size[parttype][0], step_v0 = input->part_size[parttype][1];
  int step_h = input->part_size[blocktype][0], step_v = input->part_size[blocktype][1];
  int bx = bx0[parttype][block8x8], by = by0[parttype][block8x8];
  int v, h, rx, ry, k, n = 0, mvd_bits = 0, mcost, use8 = input->Transform8x8Mode && blocktype <= 4;
  /* MBAFF frames too: LumaPrediction takes its pictures from listX[list + currMB->list_offset] (macroblock.c:862-871) -- the lists
   * mb_sync has set for this macroblock -- and the reference does not combine weighted prediction with MBAFF (configfile.c:1114) */
  if (!active || !use_mc) return __real_BIDPartitionCost(currMB, blocktype, block8x8, ref_l0, ref_l1, lambda_factor);
  mb_sync();
  /* motion-vector rate (mv-search.c:1083-1096): host arithmetic on the encoder's own arrays */
  for (v = by; v < by + step_v0; v += step_v)
    for (h = bx; h < bx + step_h0; h += step_h) {
      mvd_bits += mvbits[img->all_mv[v][h][LIST_0][ref_l0][blocktype][0] - img->pred_mv[v][h][LIST_0][ref_l0][blocktype][0]];
      mvd_bits += mvbits[img->all_mv[v][h][LIST_0][ref_l0][blocktype][1] - img->pred_mv[v][h][LIST_0][ref_l0][blocktype][1]];
      mvd_bits += mvbits[img->all_mv[v][h][LIST_1][ref_l1][blocktype][0] - img->pred_mv[v][h][LIST_1][ref_l1][blocktype][0]];
      mvd_bits += mvbits[img->all_mv[v][h][LIST_1][ref_l1][blocktype][1] - img->pred_mv[v][h][LIST_1][ref_l1][blocktype][1]];
    }
  mcost = WEIGHTED_COST(lambda_factor, mvd_bits);
  /* residual (mv-search.c:1100-1127): the partition's 8x8 regions, four 4x4 predictions each */
  for (ry = by; ry < by + step_v0; ry += 2)
    for (rx = bx; rx < bx + step_h0; rx += 2, n++)
      for (k = 0; k < 4; k++) fill_pred_block(&blocks[4 * n + k], currMB, rx + (k & 1), ry + (k >> 1), 2, blocktype, blocktype, ref_l0, ref_l1);
  if (L.pred_cost8x8(blocks, n, input->ModeDecisionMetric, c4, c8)) die("me_cuda_pred_cost8x8");
  for (k = 0; k < n; k++) mcost += use8 ? c8[k] : c4[k];
  keep_prediction_side_effects(currMB, 2, ref_l0, ref_l1);
  served_bid_cost++;
  if (verify) {
    int rc = __real_BIDPartitionCost(currMB, blocktype, block8x8, ref_l0, ref_l1, lambda_factor);
    n_checked++;
    if (rc != mcost) {
      n_mismatch++;
      fprintf(stderr, "VERIFY BIDPartitionCost pic %d mb %d bt %d blk %d refs (%d,%d): gpu %d ref %d\n", pic_counter, img->current_mb_nr, blocktype, block8x8,
              (int)ref_l0, (int)ref_l1, mcost, rc);
    }
    return rc;
  }
  return mcost;
}

int __real_GetDirectCost8x8(Macroblock *currMB, int block, int *cost8x8);
int __wrap_GetDirectCost8x8(Macroblock *currMB, int block, int *cost8x8)
{
  me_pred_block blocks[4];
  int32_t c4, c8;
  int k, bx = (block % 2) * 2, by = (block / 2) * 2, cost, in8 = *cost8x8;
  if (!active || !use_mc) return __real_GetDirectCost8x8(currMB, block, cost8x8);
  mb_sync();
  for (k = 0; k < 4; k++) {
    int h = bx + (k & 1), v = by + (k >> 1);
    int py = (img->opix_y >> 2) + v, px = (img->opix_x >> 2) + h;                    /* pic_pix_y >> 2, pic_pix_x >> 2 (mv-search.c:1290) */
    fill_pred_block(&blocks[k], currMB, h, v, direct_pdir[py][px], 0, 0, direct_ref_idx[LIST_0][py][px], direct_ref_idx[LIST_1][py][px]);
  }
  if (L.pred_cost8x8(blocks, 1, input->ModeDecisionMetric, &c4, &c8)) die("me_cuda_pred_cost8x8");
  if (c4 == INT_MAX) { *cost8x8 = INT_MAX; cost = INT_MAX; }                          /* mode not allowed (mv-search.c:1290-1294) */
  else {
    cost = c4;
    if ((input->rdopt == 0) && input->Transform8x8Mode) *cost8x8 += c8;
    keep_prediction_side_effects(currMB, blocks[3].p_dir, blocks[3].ref[0], blocks[3].ref[1]);
  }
  served_direct_cost++;
  if (verify) {
    int r8 = in8, rc = __real_GetDirectCost8x8(currMB, block, &r8);
    n_checked++;
    if (rc != cost || r8 != *cost8x8) {
      n_mismatch++;
      fprintf(stderr, "VERIFY GetDirectCost8x8 pic %d mb %d block %d: gpu (%d,%d) ref (%d,%d)\n", pic_counter, img->current_mb_nr, block, cost, *cost8x8, rc, r8);
    }
    *cost8x8 = r8;
    return rc;
  }
  return cost;
}

/* ---------------------------------------------------------------- in-loop deblocking (loopFilter.c:87, called image.c:331) */
/* The boundary strengths stay the reference's own: GetStrengthNormal (loopFilter.c:262) reads macroblock modes, coded-block flags,
 * reference indices and vectors.  This wrapper walks the macroblocks exactly as DeblockFrame / DeblockMb do (loopFilter.c:87-253),
 * records per macroblock which edges the reference would filter and with which strengths, and hands the edge loops
 * (EdgeLoopLumaNormal :529, EdgeLoopChromaNormal :814) to the device.  Frame pictures without MBAFF, 8 bit, 4:0:0 / 4:2:0; anything
 * else runs the reference's host filter. */
extern void GetStrengthNormal(byte Strength[16], ImageParameters *img, int MbQAddr, int dir, int edge, int mvlimit);
void __real_DeblockFrame(ImageParameters *p_img, imgpel **imgY, imgpel ***imgUV);
void __wrap_DeblockFrame(ImageParameters *p_img, imgpel **imgY, imgpel ***imgUV)
{
  static me_deblock_mb *mbs = NULL;
  static unsigned mbs_cap = 0;
  unsigned i;
  int chroma = imgUV != NULL && p_img->yuv_format == YUV420;
  /* frame pictures and field pictures (a field is filtered as a picture of its own: same edge loops, GetStrengthNormal with
   * mvlimit 2, loopFilter.c:162-164); MBAFF frames (GetStrengthMBAff, mixed frame / field edges) stay the reference's host filter */
  if (!active || !use_deblock || p_img->MbaffFrameFlag || p_img->max_imgpel_value != 255 ||
      (p_img->yuv_format != YUV420 && p_img->yuv_format != YUV400) || (p_img->width & 15) || (p_img->height & 15)) {
    __real_DeblockFrame(p_img, imgY, imgUV);
    return;
  }
  if (mbs_cap < p_img->PicSizeInMbs) {
    free(mbs);
    mbs = malloc(sizeof(me_deblock_mb) * p_img->PicSizeInMbs);
    if (!mbs) die("out of memory (deblocking records)");
    mbs_cap = p_img->PicSizeInMbs;
  }
  memset(mbs, 0, sizeof(me_deblock_mb) * p_img->PicSizeInMbs);
  for (i = 0; i < p_img->PicSizeInMbs; i++)                    /* loopFilter.c:105-113 */
    if (p_img->mb_data[i].mb_type == IPCM) { p_img->mb_data[i].qp = 0; p_img->mb_data[i].qpc[0] = 0; p_img->mb_data[i].qpc[1] = 0; }
  for (i = 0; i < p_img->PicSizeInMbs; i++) {                  /* DeblockMb, loopFilter.c:128-253, up to the edge loops */
    Macroblock *MbQ = &p_img->mb_data[i];
    me_deblock_mb *r = &mbs[i];
    byte Strength[16];
    int mb_x, mb_y, dir, edge, k, left, top, non8x8;
    r->qp = (int8_t)MbQ->qp; r->qpc[0] = (int8_t)MbQ->qpc[0]; r->qpc[1] = (int8_t)MbQ->qpc[1];
    r->alpha_off = (int8_t)MbQ->LFAlphaC0Offset; r->beta_off = (int8_t)MbQ->LFBetaOffset;
    p_img->DeblockCall = 1;
    get_mb_pos((int)i, &mb_x, &mb_y, IS_LUMA);
    left = mb_x != 0; top = mb_y != 0;
    non8x8 = !MbQ->luma_transform_size_8x8_flag;
    if (MbQ->LFDisableIdc == 1) { p_img->DeblockCall = 0; continue; }
    if (MbQ->LFDisableIdc == 2) { left = MbQ->mbAvailA; top = MbQ->mbAvailB; }
    p_img->current_mb_nr = (int)i;
    CheckAvailabilityOfNeighbors(MbQ);
    for (dir = 0; dir < 2; dir++) {
      int EdgeCondition = (dir && top) || (!dir && left);
      for (edge = 0; edge < 4; edge++) {
        int sum = 0;
        if (!(edge || EdgeCondition)) continue;
        GetStrengthNormal(Strength, p_img, (int)i, dir, edge << 2, p_img->structure != FRAME ? 2 : 4);     /* mvlimit, loopFilter.c:162-164 */
        for (k = 0; k < 16; k++) sum += Strength[k];
        if (!sum) continue;
        memcpy(r->strength[dir][edge], Strength, 16);
        if ((edge & 1) && !non8x8) r->luma_skip |= (uint8_t)(1u << (4 * dir + edge));
      }
    }
    p_img->DeblockCall = 0;
  }
  if (verify) {
    /* the reference filter on copies, compared sample by sample with what the device returns; the reference result is kept */
    int w = p_img->width, h = p_img->height, cw = p_img->width_cr, ch = p_img->height_cr, y, c, bad = 0;
    imgpel *cy = malloc(sizeof(imgpel) * (size_t)w * h), *cc = chroma ? malloc(sizeof(imgpel) * (size_t)cw * ch * 2) : NULL;
    if (!cy || (chroma && !cc)) die("out of memory (verify)");
    for (y = 0; y < h; y++) memcpy(cy + (size_t)y * w, imgY[y], sizeof(imgpel) * (size_t)w);
    for (c = 0; c < (chroma ? 2 : 0); c++) for (y = 0; y < ch; y++) memcpy(cc + ((size_t)c * ch + y) * cw, imgUV[c][y], sizeof(imgpel) * (size_t)cw);
    __real_DeblockFrame(p_img, imgY, imgUV);
    {
      imgpel **ry = malloc(sizeof(imgpel *) * (size_t)h), **ru = NULL, **rv = NULL;
      for (y = 0; y < h; y++) ry[y] = cy + (size_t)y * w;
      if (chroma) {
        ru = malloc(sizeof(imgpel *) * (size_t)ch); rv = malloc(sizeof(imgpel *) * (size_t)ch);
        for (y = 0; y < ch; y++) { ru[y] = cc + (size_t)y * cw; rv[y] = cc + ((size_t)ch + y) * cw; }
      }
      if (L.deblock_frame((me_imgpel *const *)ry, (me_imgpel *const *)ru, (me_imgpel *const *)rv, w, h, cw, ch, mbs)) die("me_cuda_deblock_frame");
      for (y = 0; y < h; y++) if (memcmp(ry[y], imgY[y], sizeof(imgpel) * (size_t)w)) bad++;
      for (c = 0; c < (chroma ? 2 : 0); c++) for (y = 0; y < ch; y++) if (memcmp((c ? rv : ru)[y], imgUV[c][y], sizeof(imgpel) * (size_t)cw)) bad++;
      free(ry); free(ru); free(rv);
    }
    n_checked++;
    if (bad) { n_mismatch++; fprintf(stderr, "VERIFY DeblockFrame pic %d: %d sample rows differ\n", pic_counter, bad); }
    free(cy); free(cc);
    served_deblock++;
    deblocked_rows = bad ? NULL : imgY; deblocked_pic = pic_counter;
    return;
  }
  if (L.deblock_frame((me_imgpel *const *)imgY, chroma ? (me_imgpel *const *)imgUV[0] : NULL, chroma ? (me_imgpel *const *)imgUV[1] : NULL,
                      p_img->width, p_img->height, p_img->width_cr, p_img->height_cr, mbs))
    die("me_cuda_deblock_frame");
  served_deblock++;
  deblocked_rows = imgY; deblocked_pic = pic_counter;
}
