/*
 * me_cuda.h -- C ABI of libme_cuda.so: the B200 (sm_100a) motion-estimation path for the JM 12.4 H.264
 * reference encoder `lencod`.  Plain C, plain pointers and sizes; no CUDA or torch types cross this boundary.
 *
 * What it replaces in the reference (paths relative to the reference tree, file:line):
 *   - integer-pel fast full search      lencod/src/me_fullfast.c:491 (SetupFastFullPelSearch), :210 (SetupLargerBlocks),
 *                                       :833 (FastFullPelBlockMotionSearch)
 *   - quarter-pel plane generation      lencod/src/img_luma.c:45 (getSubImagesLuma), lencod/src/image.c:1601
 *   - per-block full search             lencod/src/me_fullsearch.c:47 (FullPelBlockMotionSearch, SearchMode = -1)
 *   - sub-pel refinement (SAD / SATD)   lencod/src/me_fullsearch.c:341 (SubPelBlockMotionSearch),
 *                                       lencod/src/me_distortion.c:351,657,182,272
 *   - bi-predictive search              lencod/src/me_fullsearch.c:164 (FullPelBlockMotionBiPred), :520 (SubPelBlockSearchBiPred),
 *                                       lencod/src/me_distortion.c:482,824
 *   - reference-plane addressing        lencod/src/refbuf.c:25,37 (FastLine4X / UMVLine4X)
 *
 * Two ways in:
 *   (1) drop-in calls with the reference's own signatures (me_cuda_FastFullPelBlockMotionSearch, ...), one call per
 *       block search, synchronous -- what the SearchMode dispatch of mv-search.c:741-748,790-826,958-1021 binds to;
 *   (2) the frame-batch call me_cuda_frame_search(): every (macroblock, list, ref) search of a picture in one
 *       launch sequence -- the throughput path bench.py measures.
 * Both run the same device code.  There is no CPU fallback: every entry point fails (non-zero / error handler) when
 * no CUDA device is usable.
 *
 * All calls must come from one host thread (the reference is single-threaded and non-reentrant).
 */
#ifndef ME_CUDA_H
#define ME_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ME_CUDA_ABI_VERSION 2          /* 2: me_mb_motion grew (MBAFF), bit_depth 8..14, pattern sub-pel entry points, field views */

/* JM's pixel type in the FRExt build: `typedef unsigned short imgpel` (lencod/inc/global.h:50). */
typedef unsigned short me_imgpel;

/* Opaque handle of a reference picture whose 16 quarter-pel planes live in device memory (0 = invalid). */
typedef int32_t me_ref_handle;

#define ME_NUM_PARTITIONS 41   /* 1 + 2 + 2 + 4 + 8 + 8 + 16 blocks of types 1..7 (SetupLargerBlocks, me_fullfast.c:210-273) */
#define ME_IMG_PAD 20          /* IMG_PAD_SIZE, lencod/inc/defines.h:107 */

/* Slice types as in lencod/inc/global.h (P_SLICE = 0, B_SLICE = 1). */
#define ME_P_SLICE 0
#define ME_B_SLICE 1

/* Encoder-wide parameters, read once (mirrors what Init_Motion_Search_Module, mv-search.c:290-433, and
 * InitializeFastFullIntegerSearch, me_fullfast.c:71-141, take from `input` and `img`). */
typedef struct me_cuda_params {
  int32_t abi_version;          /* ME_CUDA_ABI_VERSION */
  int32_t width, height;        /* coded luma size (img->width, img->height): multiples of 16 */
  int32_t bit_depth;            /* img->bitdepth_luma, 8..14.  More than 8 (imgpel is u16, global.h:50): plane sets and
                                   the current picture hold 16-bit samples, the clips of the interpolation are at (1 << bit_depth) - 1;
                                   served by the row-pointer entry points (me_cuda_ref_create, me_cuda_set_current,
                                   me_cuda_ref_download_region), every drop-in search (me_cuda_ffs_setup, me_cuda_pms_*, the block
                                   searches and refinements of all search modes, field views included) and me_cuda_frame_search (general
                                   kernels), weighted-prediction ME included, and me_cuda_pred_cost8x8.  Everything else -- the *_u8 /
                                   *_device picture entry points, me_cuda_frame_bipred, chroma-aware ME, me_cuda_luma_prediction,
                                   deblocking -- is for 8-bit pictures and fails with a message instead of reading 16-bit samples as
                                   bytes */
  int32_t search_range;         /* input->search_range */
  int32_t max_num_references;   /* img->max_num_references */
  int32_t rdopt;                /* input->rdopt */
  int32_t full_search;          /* input->full_search (cfg RestrictSearchRange): 2 = same range for every reference */
  int32_t me_metric[3];         /* input->MEErrorMetric[F_PEL,H_PEL,Q_PEL]: 0 SAD, 1 SSE, 2 Hadamard SAD */
  int32_t transform8x8;         /* input->Transform8x8Mode != 0 */
  int32_t mv_limit_min_y;       /* LEVELMVLIMIT[img->LevelIndex][0], mv-search.h:36 */
  int32_t mv_limit_max_y;       /* LEVELMVLIMIT[img->LevelIndex][1] */
  int32_t chroma_me_enable;     /* input->ChromaMEEnable: 0, 1 (chroma in the integer-pel searches) or 2 (also later stages); see below */
  int32_t weighted_me;          /* informational: weighted-prediction ME in use (weights arrive through me_cuda_set_ref_weights) */
  int32_t device;               /* CUDA device ordinal; -1: env ME_CUDA_DEVICE, else 0 */
  int32_t chroma_me_weight;     /* input->ChromaMEWeight (used when chroma_me_enable != 0) */
  int32_t yuv_format;           /* img->yuv_format: 1 = 4:2:0, 2 = 4:2:2, 3 = 4:4:4 (needed when chroma_me_enable != 0); 0 = not given */
  int32_t reserved[6];
} me_cuda_params;

/* error handling: the drop-in search calls return a cost like the reference functions do, so failures are reported
 * through a handler (default: print to stderr and abort()).  A JM build installs error(text, code) (filehandle.c:38). */
typedef void (*me_cuda_error_fn)(const char *text, int code);
void        me_cuda_set_error_handler(me_cuda_error_fn fn);
const char *me_cuda_last_error(void);

/* ---- life cycle: Init_Motion_Search_Module (mv-search.c:431) / Clear_Motion_Search_Module (mv-search.c:461) ---- */
int  me_cuda_init(const me_cuda_params *p);            /* 0 on success */
void me_cuda_shutdown(void);

/* ---- reference pictures: body of getSubImagesLuma (img_luma.c:45), called from UnifiedOneForthPix (image.c:1642) ----
 * imgY_rows: h row pointers of w reconstructed luma samples (StorablePicture::imgY, mbuffer.h:52).
 * host_sub : NULL, or JM's imgY_sub[4][4][h+40][w+40] (mbuffer.h:53) to be filled with the 16 planes, because
 *            the encoder's motion compensation (macroblock.c:816) still reads them on the host. */
int me_cuda_ref_create(const me_imgpel *const *imgY_rows, int w, int h, me_imgpel ****host_sub, me_ref_handle *out);
/* same, from a contiguous 8-bit plane (bench / tests) */
int me_cuda_ref_create_u8(const uint8_t *luma, int stride, int w, int h, me_ref_handle *out);
/* free_storable_picture (mbuffer.c:556-600) */
int me_cuda_ref_destroy(me_ref_handle h);
/* rebuild the planes of an existing handle from a new picture, in place (a decoded-picture-buffer frame store being reused).
 * Buffer lifetime: this call and me_cuda_set_current_u8 queue the copy straight from the caller's buffer and return; with
 * pinned memory the DMA may still be reading it.  The buffer must stay untouched until the next call that drains the
 * library stream (me_cuda_frame_search, me_cuda_frame_bipred, me_cuda_ref_download_plane, me_cuda_sync, any drop-in search). */
int me_cuda_ref_update_u8(me_ref_handle h, const uint8_t *luma, int stride, int w, int hgt);
/* copy one of the 16 planes back: dst[(h+40)][dst_stride], plane index as Pic[fy][fx] (refbuf.c:27) */
int me_cuda_ref_download_plane(me_ref_handle h, int fy, int fx, uint8_t *dst, int dst_stride);

/* ---- current picture: pCurImg after ReadOneFrame (image.c:409-421) ---- */
int me_cuda_set_current(const me_imgpel *const *rows, int w, int h);
int me_cuda_set_current_u8(const uint8_t *luma, int stride, int w, int h);

/* ---- reference lists: mirror of listX[list + list_offset][0..n-1] (mbuffer.h:174) for the slice being coded ---- */
int me_cuda_set_ref_list(int list, int n, const me_ref_handle *handles);
/* img->type of the slice being coded (ME_P_SLICE / ME_B_SLICE); only the !rdopt zero-vector bonus depends on it */
int me_cuda_set_slice_type(int slice_type);

/* ---- weighted-prediction ME (cfg UseWeightedReferenceME with WeightedPrediction / WeightedBiprediction): the reference's
 * uni-predictive searches replace every reference sample p of list entry (list, ref) by
 *     iClip1(255, ((weight * p + (1 << (log_denom - 1))) >> log_denom) + offset)
 * with weight = wp_weight[list][ref][0], offset = wp_offset[list][ref][0], log_denom = luma_log_weight_denom
 * (me_fullfast.c:507-508,530-534,640-737; me_fullsearch.c:368-388; computeSADWP / computeSSEWP / computeSATDWP,
 * me_distortion.c:413,1107,734; weighted_prediction.c:50).  apply = 0 switches the entry back to plain samples.  The weights
 * stay with the list entry until changed; set them after me_cuda_set_ref_list, before the searches of the slice.
 * The bi-predictive searches never read these weighted samples (see me_cuda_set_bipred_weights). ---- */
int me_cuda_set_ref_weights(int list, int ref, int apply, int weight, int offset, int log_denom);
/* Weighted bi-prediction (active_pps->weighted_bipred_idc > 0): FullPelBlockMotionBiPred / SubPelBlockSearchBiPred called with
 * list argument `list` predict iClip1(255, ((weight1 * p1 + weight2 * p2 + 2 * round) >> (log_denom + 1)) + offset_bi), p1 = sample
 * of reference 0 of `list` at the fixed vector, p2 = sample of reference 0 of the other list at the candidate, with the values
 * those functions derive at me_fullsearch.c:202-227 / :541-582 (wbp_weight, wp_offset, luma_log_weight_denom), distortion =
 * computeBiPredSAD2 / SSE2 / SATD2 (me_distortion.c:556, :1254, :907 -- including the source-pointer slip of the 8x8 branch
 * of computeBiPredSATD2, :1015, which the encoder's results depend on).  apply = 0: plain (p1 + p2 + 1) >> 1.  Used by the
 * drop-in calls and by me_cuda_frame_bipred. */
int me_cuda_set_bipred_weights(int list, int apply, int weight1, int weight2, int offset_bi, int log_denom);
/* the same for chroma component 0 / 1 when ChromaMEEnable puts chroma terms into the bi-predictive distortions: weight1_cr[k],
 * weight2_cr[k], offsetBi_cr[k] (me_fullsearch.c:245-262; the reference applies them with the luma denominator,
 * me_distortion.c:565-566, 634).  After me_cuda_set_bipred_weights(list, 1, ...), which resets chroma to the default weights.
 * Chroma samples of the UNI-predictive searches are never weighted here: the reference's estimators leave weight_cr / offset_cr
 * at their defaults (weighted_prediction.c:98-103, 308-313); a caller with other values must not use the chroma terms. */
int me_cuda_set_bipred_chroma_weights(int list, int component, int weight1, int weight2, int offset_bi);

/* ---- chroma terms of the searches (cfg ChromaMEEnable = 1 or 2): the fast full search adds the
 * distortion of the co-located chroma block of every 4x4 (me_fullfast.c:776-814), computeSAD / computeSSE and their
 * bi-predictive forms add ChromaMEWeight times the chroma block distortion of each component (me_distortion.c:376-402,
 * :516-543, :1069-1091); the Hadamard functions have no chroma term.  Reference samples are the 1/8-sample sub-images of
 * getSubImagesChroma (img_chroma.c:374), read through UMVLine8X_chroma (refbuf.c:53); the library evaluates them on the fly
 * from the reconstructed chroma planes (4:2:0, 4:2:2; for 4:4:4 see me_cuda_ref_set_chroma_planes).  Supported with every metric combination and together
 * with weighted prediction (see me_cuda_set_bipred_chroma_weights for what is weighted).
 *   me_cuda_set_current_chroma : imgUV_org[0..1] of the picture being coded (cw x ch samples each)
 *   me_cuda_ref_set_chroma     : StorablePicture::imgUV[0..1] of a reference picture, after me_cuda_ref_create
 *   me_cuda_set_chroma_state   : the value of the global ChromaMEEnable (me_distortion.c:56) the NEXT search call sees -- the
 *                                reference sets it to input->ChromaMEEnable before the integer-pel search (mv-search.c:612)
 *                                and to (input->ChromaMEEnable == 2) from the sub-pel stage on (mv-search.c:779), which is
 *                                what the bi-predictive searches of the same BlockMotionSearch call run with ---- */
int me_cuda_set_current_chroma(const me_imgpel *const *u_rows, const me_imgpel *const *v_rows, int cw, int ch);
int me_cuda_set_current_chroma_u8(const uint8_t *u, const uint8_t *v, int stride, int cw, int ch);
int me_cuda_ref_set_chroma(me_ref_handle h, const me_imgpel *const *u_rows, const me_imgpel *const *v_rows, int cw, int ch);
int me_cuda_ref_set_chroma_u8(me_ref_handle h, const uint8_t *u, const uint8_t *v, int stride, int cw, int ch);
/* 4:4:4 (common mode): the chroma sub-images are made by the luma interpolator from the U and V planes (image.c:1647-1657), so
 * they are ordinary pictures of this library: create them with me_cuda_ref_create from StorablePicture::imgUV[0] / [1] and
 * attach them to the luma picture.  The three handles are destroyed separately. */
int me_cuda_ref_set_chroma_planes(me_ref_handle h, me_ref_handle h_u, me_ref_handle h_v);
int me_cuda_set_chroma_state(int enable);

/* ---- interlace (PicInterlace / MbInterlace).  Reference pictures and the current picture may be frames of the
 * initialised size or fields of such frames (same width, half the rows: dpb_split_field, mbuffer.c:2971-3024; field_picture,
 * image.c:1015): me_cuda_ref_create* / me_cuda_set_current* accept both, every picture keeps its own size.  When the whole
 * original frame is on the device, the field a macroblock belongs to need not be uploaded again: view 0 = the picture as uploaded,
 * 1 / 2 = its top / bottom field (rows 2y / 2y + 1; imgY_org_top / imgY_org_bot are row pointers into imgY_org_frm, lencod.c:1970,
 * memalloc.c:27), selected per macroblock as the reference sets pCurImg (macroblock.c:165-190 for MBAFF frames, image.c:2983, :3001
 * for field pictures).  Covers the original chroma planes too.  Read by the drop-in block searches, me_cuda_pms_begin and
 * me_cuda_pred_cost8x8; the frame batch wants view 0.  me_cuda_set_current* resets the view to 0. ---- */
int me_cuda_select_current_view(int view);

/* ---- per macroblock: ResetFastFullIntegerSearch (me_fullfast.c:197, called macroblock.c:509) ---- */
int me_cuda_mb_begin(int opix_x, int opix_y);

/* ---- per (list, ref) of the macroblock: SetupFastFullPelSearch (me_fullfast.c:491).  The caller supplies the 16x16
 * motion-vector predictor (SetMotionVectorPredictor(..., 0, 0, 16, 16), me_fullfast.c:550) which stays host code. ---- */
int me_cuda_ffs_setup(int list, int ref, int pmv16_x, int pmv16_y);
int me_cuda_ffs_is_set_up(int list, int ref);

/* ---- drop-in block searches; arguments and return value as the reference functions, minus the unused orig_pic
 * (the library reads the original block from the current picture).  mv in/out as in the reference. ---- */
int me_cuda_FastFullPelBlockMotionSearch(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                         short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y,
                                         int search_range, int min_mcost, int lambda_factor);   /* me_fullfast.h:21 */
/* the per-block spiral full search of SearchMode = -1 (and of the `else` branch of the dispatch, mv-search.c:749-768); mv in:
 * search centre in integer pel as set at mv-search.c:751-762.  Needs no me_cuda_ffs_setup. */
int me_cuda_FullPelBlockMotionSearch(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                     short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y,
                                     int search_range, int min_mcost, int lambda_factor);       /* me_fullsearch.h:21, me_fullsearch.c:47 */
int me_cuda_SubPelBlockMotionSearch(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                    short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y,
                                    int search_pos2, int search_pos4, int min_mcost,
                                    const int *lambda_factor);                                  /* me_fullsearch.h:126 */
int me_cuda_FullPelBlockMotionBiPred(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                     short pred_mv_x1, short pred_mv_y1, short pred_mv_x2, short pred_mv_y2,
                                     short *mv_x, short *mv_y, short *s_mv_x, short *s_mv_y,
                                     int search_range, int min_mcost, int lambda_factor);       /* me_fullsearch.h:122 */
int me_cuda_SubPelBlockSearchBiPred(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                    short pred_mv_x, short pred_mv_y,
                                    short *mv_x, short *mv_y, short *s_mv_x, short *s_mv_y,
                                    int search_pos2, int search_pos4, int min_mcost,
                                    const int *lambda_factor);                                  /* me_fullsearch.h:129 */

/* ---- sub-pel stages of the pattern searches (SearchMode 1 UMHexagonS, 2 simplified UMHexagonS, 3 EPZS).
 * The integer-pel stages of these modes stay the reference's host code (data-dependent walks of a few positions); what the dispatch
 * at mv-search.c:790-826 and :984-1022 calls afterwards is served here.  Arguments and return value as the reference functions minus
 * orig_pic; state the reference keeps in file statics / globals of the mode arrives as trailing arguments:
 *   subthres      subthres[blocktype] of EPZSInit (me_epzs.c:335-354; a file static there)
 *   uplayer_x/y   smpUMHEX_pred_MV_uplayer_X / _Y (me_umhexsmp.h:44-45, set by smpUMHEX_setup, me_umhexsmp.c:1203-1214)
 *   threshold1/3  SubPelThreshold1 / SubPelThreshold3 (me_umhexsmp.h:34-35, me_umhexsmp.c:106-107); shifted by the block type inside */
int me_cuda_EPZSSubPelBlockMotionSearch(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                        const short pred_mv[2], short mv[2], int search_pos2, int search_pos4, int min_mcost,
                                        const int *lambda, int subthres);                         /* me_epzs.h, me_epzs.c:2390 */
int me_cuda_EPZSSubPelBlockSearchBiPred(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                        const short *pred_mv1, const short *pred_mv2, short mv[2], const short s_mv[2],
                                        int search_pos2, int search_pos4, int min_mcost, const int *lambda);   /* me_epzs.c:2728 */
int me_cuda_UMHEXSubPelBlockMotionSearch(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                         short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y,
                                         int search_pos2, int search_pos4, int min_mcost, int lambda_factor);   /* me_umhex.c:562 */
int me_cuda_smpUMHEXSubPelBlockMotionSearch(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                            short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y,
                                            int search_pos2, int search_pos4, int min_mcost, int lambda_factor,
                                            int uplayer_x, int uplayer_y, int threshold1, int threshold3);       /* me_umhexsmp.c:616 */
int me_cuda_smpUMHEXFullSubPelBlockMotionSearch(short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                                short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y,
                                                int search_pos2, int search_pos4, int min_mcost, int lambda_factor,
                                                int threshold1, int threshold3);                                   /* me_umhexsmp.c:422 */

/* ---- one launch per PartitionMotionSearch (mv-search.c:1378-1454) instead of one per block search ----
 * PartitionMotionSearch(blocktype, block8x8) runs, for every list and reference, BlockMotionSearch on each sub-block of the
 * partition: SetMotionVectorPredictor (mv-search.c:87-281) from the motion state around the macroblock, the integer-pel search,
 * the sub-pel refinement, and writes the result back into that state for the predictors of the next sub-block
 * (mv-search.c:1441-1449).  me_cuda_pms_begin() runs all of that on the device for up to ME_PMS_MAX_ITEMS such calls that read the
 * SAME motion state (the caller's speculation: e.g. 16x16 with the first 16x8 and 8x16 partitions, whose neighbours all lie
 * outside the macroblock, or block types 4..7 of one 8x8 block) in ONE launch, and keeps the results.  The drop-in calls
 * me_cuda_FastFullPelBlockMotionSearch / me_cuda_SubPelBlockMotionSearch that the reference's own PartitionMotionSearch then
 * makes are answered from those results when -- and only when -- their arguments (block, predictor, lambda factors, search
 * range, start vector, start cost) equal what the device used; anything else takes the per-block path as before.  So the
 * speculation can cost time but never changes a result.  Needs me_cuda_ffs_setup() for every listed (list, ref) of the
 * macroblock first.  Returns 0 on success, -1 when the configuration is not covered (SearchMode other than the fast full
 * search, more than ME_PMS_MAX_REFS list entries): the caller simply goes on without it. */
#define ME_PMS_MAX_ITEMS 4
#define ME_PMS_MAX_REFS 16
#define ME_PMS_UNAVAILABLE (-128)      /* me_mb_motion::ref_idx of a 4x4 block outside the picture / slice */
typedef struct me_mb_motion {          /* enc_picture->ref_idx[list] / ->mv[list] of the 4x4 blocks getLuma4x4Neighbour can reach */
  int8_t  ref_idx[2][5][6];            /* [list][row + 1][col + 1], rows -1..3, columns -1..4 of the macroblock's 4x4 grid (column 4: row -1) */
  int16_t mv[2][5][6][2];
  /* MBAFF frames (mbaff != 0).  The neighbour a sample position maps to depends on the frame / field state of both macroblock pairs
   * (getAffNeighbour, mb_access.c), so the cells above are taken at exactly the positions SetMotionVectorPredictor asks for -- left
   * column at (-1, 4 * row), row above at (4 * col, -1) -- and the neighbour D of a block in column 0, (-1, 4 * row - 1), which is no
   * longer the cell of the row above, travels separately in d_left[list][row].  Reference indices and vertical vector components of
   * cells outside the macroblock arrive already converted between frame and field units (mv-search.c:140-172, :200-236: index * 2 or
   * >> 1, vector / 2 or * 2): the device-side predictor is the same for both kinds of picture. */
  int8_t  mbaff;
  int8_t  reserved[3];
  int8_t  d_left_ref_idx[2][4];
  int16_t d_left_mv[2][4][2];
} me_mb_motion;                        /* 344 bytes */
typedef struct me_pms_item { int8_t blocktype, block8x8; } me_pms_item;
int me_cuda_pms_begin(const me_mb_motion *state, int n_items, const me_pms_item *items, const int *lambda_factor /* [3] */);
int me_cuda_pms_cached(int blocktype, int block8x8);     /* 1: results of this PartitionMotionSearch call are waiting */
int me_cuda_pms_end(int blocktype, int block8x8);        /* the call has returned: drop them */
/* counters since init: group launches, drop-in calls answered from their results, calls whose arguments did not match */
int me_cuda_pms_stats(int64_t *launches, int64_t *served, int64_t *missed);

/* ---- frame batch (no single reference call site: it is PartitionMotionSearch, mv-search.c:1378, for every
 * macroblock x list x ref of a picture at once, with the predictors supplied by the caller) ----
 * Partition order p = 0..40: type 1 {0}; type 2 {0,8}; type 3 {0,2}; type 4 {0,2,8,10}; type 5 {0,2,..,14};
 * type 6 {0,1,2,3,8,9,10,11}; type 7 {0..15}; numbers are the top-left 4x4 block index 4*row+col inside the
 * macroblock, i.e. BlockSAD[list][ref][type][index] of me_fullfast.c:855. */
typedef struct me_mb_job {
  int16_t opix_x, opix_y;               /* macroblock origin in the current picture (img->opix_x/y) */
  int8_t  list, ref;                    /* reference = handle set by me_cuda_set_ref_list(list)[ref] */
  int8_t  slice_type;                   /* ME_P_SLICE / ME_B_SLICE */
  int8_t  reserved;
  int16_t pmv16[2];                     /* 16x16 predictor -> search centre (me_fullfast.c:550-566) */
  int16_t pmv[ME_NUM_PARTITIONS][2];    /* predictor of every partition, quarter-pel (x, y) */
  int32_t lambda[3];                    /* lambda_factor[F_PEL,H_PEL,Q_PEL] (defines.h:125) */
} me_mb_job;                            /* 188 bytes */

typedef struct me_mb_result {
  int16_t mv[ME_NUM_PARTITIONS][2];     /* final quarter-pel vector (what BlockMotionSearch writes to all_mv, mv-search.c:855) */
  int32_t cost[ME_NUM_PARTITIONS];      /* min_mcost returned by BlockMotionSearch (mv-search.c:1042) */
} me_mb_result;                         /* 328 bytes */

typedef struct me_mb_result_int {       /* optional: the integer-pel stage on its own (mv in integer pels, as *mv_x of me_fullfast.c:900) */
  int16_t mv[ME_NUM_PARTITIONS][2];
  int32_t cost[ME_NUM_PARTITIONS];
} me_mb_result_int;

/* host buffers in, host buffers out; synchronous.  out_int may be NULL. */
int me_cuda_frame_search(const me_mb_job *jobs, int n_jobs, me_mb_result *out, me_mb_result_int *out_int);

/* ---- frame batch, bi-predictive refinement of B slices: the block of BlockMotionSearch (mv-search.c:864-1034) that follows
 * the uni-predictive search of block type 1 against reference 0 of `list` -- (1 + refinements) integer-pel searches
 * (FullPelBlockMotionBiPred, me_fullsearch.c:164) with the range halved each time and the two lists taking turns, then up
 * to two sub-pel refinements (SubPelBlockSearchBiPred, me_fullsearch.c:520).  One job per (macroblock, list); the pictures
 * are reference 0 of both lists as set by me_cuda_set_ref_list. */
typedef struct me_bipred_job {
  int16_t opix_x, opix_y;               /* macroblock origin */
  int8_t  list;                         /* the list whose 16x16 / reference-0 search has just finished */
  int8_t  reserved[3];
  int16_t pmv[2];                       /* pred_mv: predictor of `list`, reference 0 (quarter pel) */
  int16_t pmv_bi[2];                    /* pred_mv_bi: predictor of the other list, reference 0 (mv-search.c:880) */
  int16_t mv[2];                        /* quarter-pel vector that uni-predictive search returned (mv-search.c:855) */
  int32_t lambda[3];                    /* lambda_factor[F_PEL,H_PEL,Q_PEL] */
} me_bipred_job;                        /* 32 bytes */

typedef struct me_bipred_result {
  int16_t mv[2];                        /* bipred_mv[..][iterlist][0][1]     (mv-search.c:1028-1029), quarter pel */
  int16_t bimv[2];                      /* bipred_mv[..][iterlist ^ 1][0][1] (mv-search.c:1030-1031) */
  int32_t cost;                         /* min_mcostbi after the last stage */
  int32_t iterlist;                     /* list ^ (refinements & 1) */
} me_bipred_result;                     /* 16 bytes */

/* refinements = input->BiPredMERefinements, search_range = input->BiPredMESearchRange (<= 64), subpel = input->BiPredMESubPel
 * (0, 1 or 2).  Host buffers in, host buffers out; synchronous. */
int me_cuda_frame_bipred(const me_bipred_job *jobs, int n_jobs, int refinements, int search_range, int subpel,
                         me_bipred_result *out);

/* ---- motion-compensated luma prediction and the mode-decision costs built on it: the host consumers
 * of the quarter-pel planes other than the searches.  With these served by the library the planes need not be copied back into
 * StorablePicture::imgY_sub (me_cuda_ref_create with host_sub = NULL). ----
 *
 * LumaPrediction (macroblock.c:832-930) of one block: p_dir 0 / 1 fetch the block from the plane set of one list entry with
 * OneComponentLumaPrediction (macroblock.c:807-827: UMVLine4X on the block origin, refbuf.c:37, then plain rows), p_dir 2 from
 * both and average ((l0 + l1 + 1) >> 1); with apply_weights the explicit / implicit weighted forms of macroblock.c:876-909. */
typedef struct me_pred_block {
  int16_t pix_x, pix_y;                 /* top-left sample of the block in the picture (img->opix_x + block_x, img->opix_y + block_y) */
  int8_t  p_dir;                        /* 0: list 0, 1: list 1, 2: both; < 0: no prediction (GetDirectCost8x8: mode not allowed) */
  int8_t  ref[2];                       /* l0_ref_idx, l1_ref_idx: indices into the lists of me_cuda_set_ref_list */
  int8_t  apply_weights;                /* active_pps->weighted_pred_flag (P) / weighted_bipred_idc (B) */
  int16_t mv[2][2];                     /* quarter-pel vectors, [list][x, y] (all_mv / bipred_mv of the block) */
  int16_t weight[2];                    /* p_dir 0 / 1: wp_weight[list][ref][0] in weight[0]; p_dir 2: wbp_weight[0 / 1][l0][l1][0] */
  int16_t offset;                       /* p_dir 0 / 1: wp_offset[list][ref][0]; 2: (wp_offset[0][l0][0] + wp_offset[1][l1][0] + 1) >> 1 */
  int16_t log_denom;                    /* luma_log_weight_denom */
} me_pred_block;                        /* 24 bytes */

/* n blocks of bsx x bsy samples (4, 8 or 16 each way): pred receives n * bsx * bsy bytes, every block row by row -- the rows
 * img->mpr[block_y ..][block_x ..] LumaPrediction writes.  Synchronous. */
int me_cuda_luma_prediction(const me_pred_block *blocks, int n, int bsx, int bsy, uint8_t *pred);

/* The residual cost all three mode-decision functions are made of, per 8x8 luma region: blocks[4 * r + k], k = 0..3 in raster
 * order, are the 4x4 blocks of region r, each predicted by itself (LumaPrediction with block size 4x4, as the reference does);
 * cost4[r] = sum over the four blocks of distortion4x4(original - prediction) (me_distortion.c:76), cost8[r] = distortion8x8
 * of the whole region (me_distortion.c:110).  metric = input->ModeDecisionMetric: 0 SAD, 1 squared error, 2 Hadamard SAD.
 * A region with a block whose p_dir < 0 gets cost4 = cost8 = INT_MAX.
 *   GetSkipCostMB    (mv-search.c:1136-1181): four regions, cost4 each -- or cost8 each when rdopt == 0 with Transform8x8Mode
 *   GetDirectCost8x8 (mv-search.c:1273-1325): one region, returns cost4; *cost8x8 += cost8 when rdopt == 0 with Transform8x8Mode
 *   BIDPartitionCost (mv-search.c:1052-1128): the regions of the partition (p_dir 2), cost4 -- or cost8 with Transform8x8Mode
 *                    and blocktype <= 4 -- plus the motion-vector rate
 * Original samples: the picture of me_cuda_set_current.  Synchronous. */
int me_cuda_pred_cost8x8(const me_pred_block *blocks, int n_regions, int metric, int32_t *cost4, int32_t *cost8);

/* a rectangle of one of the 16 planes as imgpel rows: dst[y * dst_stride + x] = Pic[fy][fx][y0 + y][x0 + x] (refbuf.c:27);
 * the binding pages the planes into StorablePicture::imgY_sub on demand with it instead of copying them all back */
int me_cuda_ref_download_region(me_ref_handle h, int fy, int fx, int x0, int y0, int w, int hgt, me_imgpel *dst, int dst_stride);

/* ---- in-loop deblocking of the coded picture: DeblockFrame (loopFilter.c:87-253) for frame pictures and
 * field pictures (each field a picture of its own) without MBAFF, 8-bit, 4:0:0 / 4:2:0.  The boundary strengths are the encoder's own (GetStrength, loopFilter.c:262): per macroblock
 * and (direction 0 = vertical edges / 1 = horizontal, edge 0..3) the 16 values DeblockMb got for the edges it filters, zeros for the
 * edges it does not (picture / slice boundary, LFDisableIdc, all-zero strengths). ---- */
typedef struct me_deblock_mb {
  uint8_t strength[2][4][16];
  int8_t  qp, qpc[2];                   /* Macroblock::qp, ::qpc[0..1] (0 for IPCM, loopFilter.c:105-113) */
  int8_t  alpha_off, beta_off;          /* LFAlphaC0Offset, LFBetaOffset */
  uint8_t luma_skip;                    /* bit 4 * dir + edge: the luma edge is not filtered (8x8 transform, filterNon8x8LumaEdgesFlag,
                                           loopFilter.c:154-155) although the strengths are set for the chroma edge */
  uint8_t reserved[10];
} me_deblock_mb;                        /* 144 bytes */

/* filters the picture in place (rows of imgpel; u_rows / v_rows NULL for 4:0:0): mbs holds (w / 16) * (h / 16) records in raster order.
 * The filtered luma also stays on the device: me_cuda_ref_create_deblocked builds the plane set of the stored reference picture
 * from it without another upload (the caller guarantees the picture has not changed in between).  Synchronous. */
int me_cuda_deblock_frame(me_imgpel *const *y_rows, me_imgpel *const *u_rows, me_imgpel *const *v_rows, int w, int h, int cw, int ch,
                          const me_deblock_mb *mbs);
int me_cuda_ref_create_deblocked(int w, int h, me_imgpel ****host_sub, me_ref_handle *out);

/* ---- device-resident variants for measurement (bench.py): pointers are CUDA device pointers, `stream` is a
 * cudaStream_t passed as void* (NULL = the library's own stream).  Asynchronous.  A caller's stream is ordered against the
 * library stream by events on both sides of the call, so the program order of the calls is the order on the device even
 * when they name different streams; the caller's buffers must stay valid until the stream the call named has drained. ---- */
int me_cuda_frame_search_device(const void *d_jobs, int n_jobs, void *d_out, void *d_out_int, void *stream);
int me_cuda_ref_create_device(const void *d_luma_u8, int stride, int w, int h, void *stream, me_ref_handle *out);
int me_cuda_ref_update_device(me_ref_handle hd, const void *d_luma_u8, int stride, int w, int h, void *stream);
int me_cuda_set_current_device(const void *d_luma_u8, int stride, int w, int h, void *stream);
/* per-kernel-class device timing (CUDA events around every launch; class 0 plane generation, 1 integer-pel search,
 * 2 sub-pel refinement).  me_cuda_set_timing(1) starts / resets collection, (0) stops it. */
int me_cuda_set_timing(int on);
int me_cuda_kernel_time_ms(int which, double *total_ms, int *n_launches);
void *me_cuda_stream(void);                 /* the library's stream (cudaStream_t) */
int   me_cuda_sync(void);                   /* cudaStreamSynchronize on it */
/* number of kernels this library has launched since init (bench.py reports it as gpu_launches) */
int64_t me_cuda_launch_count(void);
/* diagnostics: share of the partitions the throughput sub-pel kernel serves from its TMA-staged 48 x 33 x 16-plane region (the
 * others read their tiles from global memory with the clamp of UMVLine4X, refbuf.c:37).  enable != 0 starts and zeroes the
 * count, 0 stops it; the counts so far go to served / total (may be null). */
int me_cuda_subpel_region_stats(int enable, int64_t *served, int64_t *total);
/* diagnostics: measured issue rate (lane-operations per second, whole GPU) of the integer instructions the search
 * kernels are bound by; kind 0 VABSDIFF4, 1 IADD3, 2 IMAD, 3 VIADDMNMX, 4 LOP3, 5 SHF, 6 VIMNMX, 7 PRMT, 8..14 the kinds 1..7
 * interleaved 1:1 with VABSDIFF4, 15 LDS + VABSDIFF4, 16 / 17 VABSDIFF4 + IMAD / VIADDMNMX with the operand pattern of the search
 * kernel (one shared reference word per group, immediate multiplier: at most two register-file reads per instruction), 18 the
 * arithmetic of one window row of the sweep (32 VABSDIFF4 : 21 add+min : 33 IMAD), dependency-free: the practical ceiling of that mix.
 * This is the denominator of bench.py's INT-ALU roofline (MEASURED_PEAKS.json has no integer figure). */
int me_cuda_measure_int_peak(int kind, double *lane_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* ME_CUDA_H */
