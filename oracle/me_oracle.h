/*
 * me_oracle.h -- TEST INFRASTRUCTURE.  CPU restatement (plain C) of the reference algorithm for the
 * motion-estimation hot path of JM 12.4 lencod.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may use it; the product (libme_cuda.so) never links or calls it.
 *
 * Parity pin: every function here is checked against call traces and plane dumps of the UNMODIFIED reference
 * encoder (oracle/_ref/lencod_trace, built by oracle/Makefile from the reference sources) and against the golden
 * fixtures committed in tests/golden/ that were produced from those traces (tests/golden/make_golden.py).
 */
#ifndef ME_ORACLE_H
#define ME_ORACLE_H

#include <stdint.h>
#include "../include/me_cuda.h"   /* struct layouts of the batch interface only */

#ifdef __cplusplus
extern "C" {
#endif

/* sample type: 8-bit pictures, or -- the same source compiled with -DORA_PEL16 into liboracle_me16.so -- pictures of more than 8 bits
 * per sample (imgpel is u16 in the reference, global.h:50); strides count samples.  ora_set_max_pel: img->max_imgpel_value */
#ifdef ORA_PEL16
typedef uint16_t ora_pel;
#else
typedef uint8_t ora_pel;
#endif
void ora_set_max_pel(int max_pel_value);

typedef struct ora_cfg {
  int search_range;        /* input->search_range */
  int rdopt;               /* input->rdopt */
  int full_search;         /* input->full_search */
  int metric[3];           /* input->MEErrorMetric[] */
  int transform8x8;        /* input->Transform8x8Mode */
  int mv_limit_min_y, mv_limit_max_y;  /* LEVELMVLIMIT[LevelIndex][0..1] */
  int chroma_me;           /* input->ChromaMEEnable (0, 1 = integer pel only, 2 = also sub-pel): enters start_me_refinement_hp/qp */
} ora_cfg;

typedef struct ora_ref {
  int w, h, wp, hp;        /* picture and padded (+40) sizes */
  ora_pel *plane[4][4];    /* [vertical fraction][horizontal fraction], hp rows of wp bytes */
  /* weighted-prediction ME (UseWeightedReferenceME): every reference sample of a UNI-predictive distortion becomes
   * clip1(((wp_weight * p + wp_round) >> wp_denom) + wp_offset) (me_fullfast.c:653, me_distortion.c:413-480, :734-822) */
  int wp_apply, wp_weight, wp_offset, wp_denom;
  /* reconstructed chroma (StorablePicture::imgUV), for ChromaMEEnable; NULL when not set */
  ora_pel *crec[2];
  int cw, ch;
  /* 4:4:4: the chroma sub-images are made by the luma interpolator (image.c:1647-1657): the plane sets of U and V (not owned) */
  const struct ora_ref *c444[2];
} ora_ref;

/* tables */
int  ora_num_positions(int range);
void ora_spiral(int range, int16_t *sx, int16_t *sy);   /* (2*range+1)^2 entries */
int  ora_mvbits(int d);
int  ora_mv_cost(int lambda_factor, int cx, int cy, int px, int py);
extern const int8_t ora_part_type[ME_NUM_PARTITIONS];    /* block type 1..7 of partition p */
extern const int8_t ora_part_blk[ME_NUM_PARTITIONS];     /* top-left 4x4 block index (4*row+col) of partition p */
void ora_block_size(int blocktype, int *bsx, int *bsy);

/* quarter-pel planes (getSubImagesLuma, img_luma.c:45-639) */
ora_ref *ora_ref_create(const ora_pel *luma, int stride, int w, int h);
void     ora_ref_destroy(ora_ref *r);
const ora_pel *ora_ref_plane(const ora_ref *r, int fy, int fx);
/* weights of the (list, ref) entry this picture is searched as: wp_weight / wp_offset [list][ref][0], luma_log_weight_denom */
void     ora_ref_set_weights(ora_ref *r, int apply, int weight, int offset, int log_denom);

/* ChromaMEEnable: reconstructed chroma planes of a reference picture (copied) */
void ora_ref_set_chroma(ora_ref *r, const ora_pel *u, const ora_pel *v, int stride, int cw, int ch);
/* 4:4:4 (chroma_shift 2, 2): the quarter-pel plane sets built from the picture's U and V planes */
void ora_ref_set_chroma444(ora_ref *r, const ora_ref *u, const ora_ref *v);
/* chroma state of the NEXT search call: the global ChromaMEEnable as that call sees it (me_distortion.c:56; mv-search.c:612, :779),
 * input->ChromaMEWeight, the original chroma planes (imgUV_org) and chroma_shift_x / chroma_shift_y (lencod.c:2856-2879:
 * 3,3 for 4:2:0; 3,2 for 4:2:2) */
void ora_set_chroma(int enable, int weight, const ora_pel *cur_u, const ora_pel *cur_v, int stride, int shift_x, int shift_y);

/* plane access with JM's origin clamp (UMVLine4X, refbuf.c:37): pointer to the first pixel of a block row */
const ora_pel *ora_umv_line(const ora_ref *r, int y4, int x4);

/* integer-pel fast full search */
int  ora_setup_range(const ora_cfg *c, int ref);
void ora_search_centre(const ora_cfg *c, int ref, int opix_x, int opix_y, int pmv16_x, int pmv16_y,
                       int *centre_x, int *centre_y);
/* surf: 41 surfaces of ora_num_positions(range) ints each, partition order of me_cuda.h */
void ora_ffs_surfaces(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r,
                      int opix_x, int opix_y, int centre_x, int centre_y, int range, int32_t *surf);
int  ora_ffs_block(const ora_cfg *c, const int32_t *surf_p, int setup_range, int opix_x, int opix_y,
                   int centre_x, int centre_y, int pmv_x, int pmv_y, int search_range, int min_mcost,
                   int lambda_factor, short *mv_x, short *mv_y);
/* search range PartitionMotionSearch passes down (mv-search.c:1411-1416) */
int  ora_block_search_range(const ora_cfg *c, int ref, int blocktype);

/* distortion */
int ora_sad_block(const ora_pel *src, int src_stride, int bsx, int bsy, const ora_ref *r, int x4, int y4);
int ora_satd_block(const ora_pel *src, int src_stride, int bsx, int bsy, const ora_ref *r, int x4, int y4, int use8x8);
int ora_hadamard4x4(const int *d);
int ora_hadamard8x8(const int *d);

/* sub-pel refinement (SubPelBlockMotionSearch, me_fullsearch.c:341) */
int ora_subpel(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r,
               int pic_pix_x, int pic_pix_y, int blocktype, int ref, int slice_type,
               int pmv_x, int pmv_y, short *mv_x, short *mv_y, int search_pos2, int search_pos4,
               int min_mcost, const int *lambda);

/* per-block full search (FullPelBlockMotionSearch, me_fullsearch.c:47): SearchMode = -1 */
int ora_fullpel_block(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r,
                      int pic_pix_x, int pic_pix_y, int blocktype, int ref, int slice_type,
                      int pmv_x, int pmv_y, short *mv_x, short *mv_y, int search_range, int min_mcost, int lambda_factor);

/* weighted bi-prediction (computeBiPred*2): one weight set per value of the searches' `list` argument; ora_bipred_fullpel /
 * ora_bipred_subpel use the set chosen by ora_select_bipred_list, ora_mb_bipred switches between the two as the reference does */
void ora_set_bipred_weights(int list, int apply, int weight1, int weight2, int offset_bi, int log_denom);
void ora_select_bipred_list(int list);
void ora_set_bipred_chroma_weights(int list, int component, int weight1, int weight2, int offset_bi);   /* after ora_set_bipred_weights */

/* bi-predictive searches (me_fullsearch.c:164, :520): r1 is read at the fixed vector, r2 at the candidate */
int ora_bipred_fullpel(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r1, const ora_ref *r2,
                       int pic_pix_x, int pic_pix_y, int blocktype,
                       int pmv_x1, int pmv_y1, int pmv_x2, int pmv_y2,
                       short *mv_x, short *mv_y, const short *s_mv_x, const short *s_mv_y,
                       int search_range, int min_mcost, int lambda_factor);
int ora_bipred_subpel(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r1, const ora_ref *r2,
                      int pic_pix_x, int pic_pix_y, int blocktype, int pmv_x, int pmv_y,
                      short *mv_x, short *mv_y, const short *s_mv_x, const short *s_mv_y,
                      int search_pos2, int search_pos4, int min_mcost, const int *lambda);

/* ---- sub-pel stages of the pattern searches (SearchMode 1 UMHexagonS, 2 simplified UMHexagonS, 3 EPZS).
 * Vectors are quarter pel.  State the reference keeps in file statics / globals arrives as arguments:
 * subthres = subthres[blocktype] (me_epzs.c:354); uplayer_x / y = smpUMHEX_pred_MV_uplayer_X / _Y (me_umhexsmp.c:1203-1214);
 * threshold1 / threshold3 = SubPelThreshold1 / SubPelThreshold3 (me_umhexsmp.c:106-107), shifted inside by the block type */
int ora_epzs_subpel(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r,
                    int pic_pix_x, int pic_pix_y, int blocktype, int pmv_x, int pmv_y, short *mv_x, short *mv_y,
                    int search_pos2, int search_pos4, int min_mcost, const int *lambda, int subthres);
int ora_epzs_subpel_bipred(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r1, const ora_ref *r2,
                           int pic_pix_x, int pic_pix_y, int blocktype, int pmv1_x, int pmv1_y, int pmv2_x, int pmv2_y,
                           short *mv_x, short *mv_y, const short *s_mv_x, const short *s_mv_y,
                           int search_pos2, int search_pos4, int min_mcost, const int *lambda);
int ora_umhex_subpel(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r,
                     int pic_pix_x, int pic_pix_y, int blocktype, int pmv_x, int pmv_y, short *mv_x, short *mv_y,
                     int min_mcost, int lambda_factor);
int ora_smpumhex_subpel(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r,
                        int pic_pix_x, int pic_pix_y, int blocktype, int pmv_x, int pmv_y, short *mv_x, short *mv_y,
                        int min_mcost, int lambda_factor, int uplayer_x, int uplayer_y, int threshold1, int threshold3);
int ora_smpumhex_full_subpel(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r,
                             int pic_pix_x, int pic_pix_y, int blocktype, int ref, int slice_type, int pmv_x, int pmv_y,
                             short *mv_x, short *mv_y, int search_pos2, int search_pos4, int min_mcost, int lambda_factor,
                             int threshold1, int threshold3);

/* one batch job = PartitionMotionSearch for all 41 partitions of one (macroblock, list, ref) */
void ora_mb_search(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r,
                   const me_mb_job *job, me_mb_result *out, me_mb_result_int *out_int);

/* one bi-predictive batch job = the block of BlockMotionSearch at mv-search.c:864-1034 (block type 1, reference 0):
 * l0 / l1 = reference 0 of list 0 / list 1 */
void ora_mb_bipred(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *l0, const ora_ref *l1,
                   const me_bipred_job *job, int refinements, int search_range, int subpel, me_bipred_result *out);

#ifndef ORA_PEL16
/* ---- motion-compensated prediction and the mode-decision costs built on it (include/me_cuda.h: me_pred_block).
 * list0 / list1: the reference lists as arrays of pictures; blocks[4 * by + bx]: what LumaPrediction does for the 4x4 block
 * (bx, by) of the macroblock.  Reference: macroblock.c:807-930, me_distortion.c:76-137, mv-search.c:1052-1128, :1136-1181, :1273-1325 */
void ora_luma_prediction(const ora_ref *const *list0, const ora_ref *const *list1, const me_pred_block *b, int bsx, int bsy, uint8_t *pred);
int  ora_distortion4x4(int metric, const int *diff);
int  ora_distortion8x8(int metric, const int *diff);
int  ora_get_skip_cost_mb(const uint8_t *cur, int cur_stride, const ora_ref *const *l0, const ora_ref *const *l1,
                          const me_pred_block *blocks, int rdopt, int transform8x8, int metric);
int  ora_get_direct_cost8x8(const uint8_t *cur, int cur_stride, const ora_ref *const *l0, const ora_ref *const *l1,
                            const me_pred_block *blocks, int block, int rdopt, int transform8x8, int metric, int *cost8x8);
int  ora_bid_partition_cost(const uint8_t *cur, int cur_stride, const ora_ref *const *l0, const ora_ref *const *l1,
                            const me_pred_block *blocks, const int16_t (*pmv)[2][2], int blocktype, int block8x8,
                            int lambda_factor, int transform8x8, int metric);

/* ---- DeblockFrame (loopFilter.c:87-253) with the boundary strengths given per macroblock (include/me_cuda.h: me_deblock_mb);
 * frame pictures, no MBAFF, 8 bit, 4:2:0 (u, v non-NULL) or 4:0:0.  In place. */
void ora_deblock_frame(uint8_t *y, int ystride, uint8_t *u, uint8_t *v, int cstride, int w, int h, const me_deblock_mb *mbs);

#endif /* !ORA_PEL16 */

#ifdef __cplusplus
}
#endif
#endif
