/*
 * trace_wrap.c -- TEST INFRASTRUCTURE (oracle side).  Not part of the product path.
 *
 * A call tracer for the UNMODIFIED reference encoder.  The reference objects are compiled
 * from /root/reference/lencod/src where they lie (oracle/Makefile) and linked with
 *     -Wl,--wrap=FastFullPelBlockMotionSearch,--wrap=SubPelBlockMotionSearch,...
 * so that every cross-translation-unit call of a hot-path function lands here first.
 * Each wrapper records the arguments, forwards to the real function (__real_X) and
 * records the results.  Nothing in the encoder's behaviour changes: the bitstream of the
 * traced binary stays byte-identical to the plain one (tests/test_oracle_ref.py checks it).
 *
 * Functions intercepted (reference file:line of the definition):
 *   FastFullPelBlockMotionSearch   lencod/src/me_fullfast.c:833   (called mv-search.c:745)
 *   SubPelBlockMotionSearch        lencod/src/me_fullsearch.c:341 (called mv-search.c:799,823)
 *   FullPelBlockMotionSearch       lencod/src/me_fullsearch.c:47  (called mv-search.c:765; SearchMode = -1)
 *   FullPelBlockMotionBiPred       lencod/src/me_fullsearch.c:164 (called mv-search.c:960)
 *   SubPelBlockSearchBiPred        lencod/src/me_fullsearch.c:520 (called mv-search.c:998,1018)
 *   EPZSSubPelBlockMotionSearch    lencod/src/me_epzs.c:2390      (called mv-search.c:818; SearchMode 3 with EPZSSubPelME)
 *   EPZSSubPelBlockSearchBiPred    lencod/src/me_epzs.c:2728      (called mv-search.c:993,1013; EPZSSubPelMEBiPred)
 *   UMHEXSubPelBlockMotionSearch   lencod/src/me_umhex.c:562      (called mv-search.c:794; SearchMode 1, block types 4..7)
 *   smpUMHEXSubPelBlockMotionSearch / smpUMHEXFullSubPelBlockMotionSearch  lencod/src/me_umhexsmp.c:616, :422 (called mv-search.c:807,812; SearchMode 2)
 *   ResetFastFullIntegerSearch     lencod/src/me_fullfast.c:197   (called macroblock.c:509)
 *   getSubImagesLuma               lencod/src/img_luma.c:45       (called image.c:1642)
 *   encode_one_frame               lencod/src/image.c:346         (called lencod.c:341)
 *
 * Output (directory named by env ME_TRACE_DIR, default "."):
 *   trace.bin          fixed-size records, 32 little-endian int32 each (layout: REC_* below)
 *   cur_<n>.y          original luma of coded picture n (W*H bytes; 16-bit little-endian samples when the bit depth exceeds 8, as all dumps then); with ChromaMEEnable also cur_<n>.u / .v
 *   ref_<id>.y         reconstructed luma a set of quarter-pel planes was built from; with ChromaMEEnable also ref_<id>.u / .v
 *   ref_<id>.sub       the 16 planes as the reference produced them, [4][4][H+40][W+40] bytes
 *                      (only when ME_TRACE_PLANES=1)
 * ME_TRACE_FROM_PIC=n keeps the search records of coded pictures >= n only (picture records and dumps are always written):
 * a 1080p run is traced for its last picture without gigabytes of records of the pictures that only build up the reference list.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "global.h"
#include "image.h"
#include "mbuffer.h"
#include "me_distortion.h"
#include "me_umhexsmp.h"

#define REC_INTS 32
enum { REC_FFS = 1, REC_SUBPEL = 2, REC_BI_FULL = 3, REC_BI_SUB = 4, REC_REFPIC = 5, REC_CURPIC = 6, REC_FULLPEL = 7,
       REC_SUBPEL_PAT = 8, REC_BI_SUB_PAT = 9 };

extern const int LEVELMVLIMIT[17][6];
extern void SetMotionVectorPredictor (short pmv[2], char **refPic, short ***tmp_mv, short ref_frame,
                                      int list, int block_x, int block_y, int blockshape_x, int blockshape_y);

static FILE *trace_fp = NULL;
static const char *trace_dir = ".";
static int dump_planes = 0;
static int from_pic = 0;
static int coded_pic_no = -1;      /* incremented per encode_one_frame */
static int cur_dumped_for = -2;

/* picture-pointer -> id table for reference pictures */
#define MAX_TRACKED 64
static struct { StorablePicture *p; int id; } tracked[MAX_TRACKED];
static int n_tracked = 0, next_ref_id = 0;

/* private mirror of the static search_setup_done/search_center of me_fullfast.c:41-46 */
static int setup_done[2][32];
static int centre_x[2][32], centre_y[2][32], setup_range[2][32];

static void trace_open(void)
{
  char path[1024];
  const char *d;
  if (trace_fp) return;
  d = getenv("ME_TRACE_DIR");
  if (d && *d) trace_dir = d;
  dump_planes = getenv("ME_TRACE_PLANES") && atoi(getenv("ME_TRACE_PLANES"));
  if (getenv("ME_TRACE_FROM_PIC")) from_pic = atoi(getenv("ME_TRACE_FROM_PIC"));
  snprintf(path, sizeof path, "%s/trace.bin", trace_dir);
  trace_fp = fopen(path, "wb");
  if (!trace_fp) { perror(path); exit(97); }
}

static void put(const int *rec)
{
  trace_open();
  if (rec[0] != REC_REFPIC && rec[0] != REC_CURPIC && rec[1] < from_pic) return;
  fwrite(rec, sizeof(int), REC_INTS, trace_fp);
}

static int ref_id_of(StorablePicture *p)
{
  int i;
  for (i = n_tracked - 1; i >= 0; i--) if (tracked[i].p == p) return tracked[i].id;
  return -1;
}

static void dump_luma(const char *name, imgpel **rows, int w, int h)
{
  char path[1024];
  FILE *f;
  int x, y;
  unsigned char *line = (unsigned char *)malloc(w);
  snprintf(path, sizeof path, "%s/%s", trace_dir, name);
  f = fopen(path, "wb");
  if (!f) { perror(path); exit(97); }
  for (y = 0; y < h; y++) {
    if (img->bitdepth_luma > 8) { fwrite(rows[y], sizeof(imgpel), w, f); continue; }   /* more than 8 bits: the 16-bit samples as they are */
    for (x = 0; x < w; x++) line[x] = (unsigned char)rows[y][x];
    fwrite(line, 1, w, f);
  }
  fclose(f);
  free(line);
}

static void dump_current_if_needed(void)
{
  char name[64];
  int rec[REC_INTS] = {0};
  if (cur_dumped_for == coded_pic_no) return;
  cur_dumped_for = coded_pic_no;
  snprintf(name, sizeof name, "cur_%d.y", coded_pic_no);
  dump_luma(name, pCurImg, img->width, img->height);
  if (input->ChromaMEEnable) {
    snprintf(name, sizeof name, "cur_%d.u", coded_pic_no); dump_luma(name, imgUV_org[0], img->width_cr, img->height_cr);
    snprintf(name, sizeof name, "cur_%d.v", coded_pic_no); dump_luma(name, imgUV_org[1], img->width_cr, img->height_cr);
  }
  rec[0] = REC_CURPIC; rec[1] = coded_pic_no; rec[2] = img->type; rec[3] = img->width; rec[4] = img->height;
  rec[5] = img->framepoc; rec[6] = img->qp;
  rec[7] = input->ChromaMEEnable; rec[8] = input->ChromaMEWeight; rec[9] = img->yuv_format; rec[10] = img->width_cr; rec[11] = img->height_cr;
  rec[12] = img->bitdepth_luma;               /* > 8: the picture dumps hold 16-bit samples */
  put(rec);
}

/* ---------------------------------------------------------------- frame counter */
int __real_encode_one_frame(void);
int __wrap_encode_one_frame(void)
{
  trace_open();
  coded_pic_no++;
  return __real_encode_one_frame();
}

/* ---------------------------------------------------------------- plane generation */
void __real_getSubImagesLuma(StorablePicture *s);
void __wrap_getSubImagesLuma(StorablePicture *s)
{
  int rec[REC_INTS] = {0};
  char name[64];
  int id;
  trace_open();
  __real_getSubImagesLuma(s);
  /* 4:4:4 calls this three times per picture (Y,U,V); only the luma pass is traced */
  if (s->p_curr_img != s->imgY) return;
  id = next_ref_id++;
  if (n_tracked == MAX_TRACKED) { memmove(&tracked[0], &tracked[1], sizeof(tracked[0]) * (MAX_TRACKED - 1)); n_tracked--; }
  tracked[n_tracked].p = s; tracked[n_tracked].id = id; n_tracked++;
  snprintf(name, sizeof name, "ref_%d.y", id);
  dump_luma(name, s->imgY, s->size_x, s->size_y);
  if (input->ChromaMEEnable) {
    snprintf(name, sizeof name, "ref_%d.u", id); dump_luma(name, s->imgUV[0], s->size_x_cr, s->size_y_cr);
    snprintf(name, sizeof name, "ref_%d.v", id); dump_luma(name, s->imgUV[1], s->size_x_cr, s->size_y_cr);
  }
  if (dump_planes) {
    char path[1024];
    FILE *f;
    int fy, fx, y, x;
    int wp = s->size_x_padded, hp = s->size_y_padded;
    unsigned char *line = (unsigned char *)malloc(wp);
    snprintf(path, sizeof path, "%s/ref_%d.sub", trace_dir, id);
    f = fopen(path, "wb");
    if (!f) { perror(path); exit(97); }
    for (fy = 0; fy < 4; fy++)
      for (fx = 0; fx < 4; fx++)
        for (y = 0; y < hp; y++) {
          if (img->bitdepth_luma > 8) { fwrite(s->imgY_sub[fy][fx][y], sizeof(imgpel), wp, f); continue; }
          for (x = 0; x < wp; x++) line[x] = (unsigned char)s->imgY_sub[fy][fx][y][x];
          fwrite(line, 1, wp, f);
        }
    fclose(f);
    free(line);
  }
  rec[0] = REC_REFPIC; rec[1] = id; rec[2] = s->size_x; rec[3] = s->size_y; rec[4] = coded_pic_no;
  rec[5] = s->poc;
  put(rec);
}

/* ---------------------------------------------------------------- per-MB reset */
void __real_ResetFastFullIntegerSearch(void);
void __wrap_ResetFastFullIntegerSearch(void)
{
  memset(setup_done, 0, sizeof setup_done);
  __real_ResetFastFullIntegerSearch();
}

static int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

/* weighted-prediction state of a uni-predictive search as the reference derives it (me_fullfast.c:507-508,530-534;
 * me_fullsearch.c:368-388): rec[28] apply_weights, rec[29] weight, rec[30] offset, rec[31] luma_log_weight_denom */
static void put_weights(int *rec, int list, int list_offset, int ref)
{
  int apply = ((active_pps->weighted_pred_flag && (img->type == P_SLICE || img->type == SP_SLICE)) ||
               (active_pps->weighted_bipred_idc && (img->type == B_SLICE))) && input->UseWeightedReferenceME;
  rec[28] = apply;
  if (apply) {
    rec[29] = wp_weight[list + list_offset][ref][0];
    rec[30] = wp_offset[list + list_offset][ref][0];
    rec[31] = luma_log_weight_denom;
  }
}

/* ---------------------------------------------------------------- integer-pel fast full search */
int __real_FastFullPelBlockMotionSearch(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y,
                                        int blocktype, short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y,
                                        int search_range, int min_mcost, int lambda_factor);
int __wrap_FastFullPelBlockMotionSearch(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y,
                                        int blocktype, short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y,
                                        int search_range, int min_mcost, int lambda_factor)
{
  int rec[REC_INTS] = {0};
  int list_offset = img->mb_data[img->current_mb_nr].list_offset;
  int cost;
  dump_current_if_needed();
  if (!setup_done[list][ref]) {
    /* what SetupFastFullPelSearch is about to derive (me_fullfast.c:505,550-566) */
    short pmv[2];
    int r = (input->full_search == 2 || ref == 0) ? input->search_range : input->search_range / 2;
    int cx, cy;
    SetMotionVectorPredictor(pmv, enc_picture->ref_idx[list], enc_picture->mv[list], ref, list, 0, 0, 16, 16);
    cx = pmv[0] / 4; cy = pmv[1] / 4;
    if (!input->rdopt) { cx = clip3(-r, r, cx); cy = clip3(-r, r, cy); }
    cx = clip3(-2047 + r, 2047 - r, cx);
    cy = clip3(LEVELMVLIMIT[img->LevelIndex][0] + r, LEVELMVLIMIT[img->LevelIndex][1] - r, cy);
    centre_x[list][ref] = cx + img->opix_x;
    centre_y[list][ref] = cy + img->opix_y;
    setup_range[list][ref] = r;
    setup_done[list][ref] = 1;
    rec[22] = 1;                      /* this call triggered the setup */
    rec[23] = pmv[0]; rec[24] = pmv[1];
  }
  rec[0] = REC_FFS; rec[1] = coded_pic_no; rec[2] = img->type; rec[3] = img->opix_x; rec[4] = img->opix_y;
  rec[5] = list; rec[6] = ref; rec[7] = blocktype; rec[8] = pic_pix_x; rec[9] = pic_pix_y;
  rec[10] = pred_mv_x; rec[11] = pred_mv_y; rec[12] = search_range; rec[13] = lambda_factor; rec[14] = min_mcost;
  rec[15] = centre_x[list][ref]; rec[16] = centre_y[list][ref]; rec[17] = setup_range[list][ref];
  rec[21] = ref_id_of(listX[list + list_offset][ref]);
  rec[25] = input->rdopt;
  rec[27] = ChromaMEEnable;                    /* the global of me_distortion.c:56 as this call sees it */
  put_weights(rec, list, list_offset, ref);
  cost = __real_FastFullPelBlockMotionSearch(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype,
                                             pred_mv_x, pred_mv_y, mv_x, mv_y, search_range, min_mcost, lambda_factor);
  rec[18] = *mv_x; rec[19] = *mv_y; rec[20] = cost;
  put(rec);
  return cost;
}

/* ---------------------------------------------------------------- per-block full search (SearchMode = -1) */
int __real_FullPelBlockMotionSearch(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y,
                                    int blocktype, short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y,
                                    int search_range, int min_mcost, int lambda_factor);
int __wrap_FullPelBlockMotionSearch(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y,
                                    int blocktype, short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y,
                                    int search_range, int min_mcost, int lambda_factor)
{
  int rec[REC_INTS] = {0};
  int list_offset = img->mb_data[img->current_mb_nr].list_offset;
  int cost;
  dump_current_if_needed();
  rec[0] = REC_FULLPEL; rec[1] = coded_pic_no; rec[2] = img->type; rec[3] = img->opix_x; rec[4] = img->opix_y;
  rec[5] = list; rec[6] = ref; rec[7] = blocktype; rec[8] = pic_pix_x; rec[9] = pic_pix_y;
  rec[10] = pred_mv_x; rec[11] = pred_mv_y; rec[12] = search_range; rec[13] = lambda_factor; rec[14] = min_mcost;
  rec[15] = *mv_x; rec[16] = *mv_y;                              /* search centre, integer pel */
  rec[21] = ref_id_of(listX[list + list_offset][ref]);
  rec[22] = test8x8transform; rec[25] = input->rdopt; rec[26] = ChromaMEEnable;
  put_weights(rec, list, list_offset, ref);
  cost = __real_FullPelBlockMotionSearch(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype,
                                         pred_mv_x, pred_mv_y, mv_x, mv_y, search_range, min_mcost, lambda_factor);
  rec[18] = *mv_x; rec[19] = *mv_y; rec[20] = cost;
  put(rec);
  return cost;
}

/* ---------------------------------------------------------------- sub-pel refinement */
int __real_SubPelBlockMotionSearch(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y,
                                   int blocktype, short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y,
                                   int search_pos2, int search_pos4, int min_mcost, int *lambda_factor);
int __wrap_SubPelBlockMotionSearch(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y,
                                   int blocktype, short pred_mv_x, short pred_mv_y, short *mv_x, short *mv_y,
                                   int search_pos2, int search_pos4, int min_mcost, int *lambda_factor)
{
  int rec[REC_INTS] = {0};
  int list_offset = img->mb_data[img->current_mb_nr].list_offset;
  int cost;
  dump_current_if_needed();
  rec[0] = REC_SUBPEL; rec[1] = coded_pic_no; rec[2] = img->type | (ChromaMEEnable << 8);   /* bits 8..: the global ChromaMEEnable */ rec[3] = img->opix_x; rec[4] = img->opix_y;
  rec[5] = list; rec[6] = ref; rec[7] = blocktype; rec[8] = pic_pix_x; rec[9] = pic_pix_y;
  rec[10] = pred_mv_x; rec[11] = pred_mv_y; rec[12] = *mv_x; rec[13] = *mv_y; rec[14] = min_mcost;
  rec[15] = lambda_factor[0]; rec[16] = lambda_factor[1]; rec[17] = lambda_factor[2];
  rec[21] = ref_id_of(listX[list + list_offset][ref]);
  rec[22] = test8x8transform; rec[23] = start_me_refinement_hp; rec[24] = start_me_refinement_qp;
  rec[25] = input->rdopt; rec[26] = search_pos2; rec[27] = search_pos4;
  put_weights(rec, list, list_offset, ref);
  cost = __real_SubPelBlockMotionSearch(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype,
                                        pred_mv_x, pred_mv_y, mv_x, mv_y, search_pos2, search_pos4, min_mcost, lambda_factor);
  rec[18] = *mv_x; rec[19] = *mv_y; rec[20] = cost;
  put(rec);
  return cost;
}

/* ---------------------------------------------------------------- bi-predictive integer-pel search */
int __real_FullPelBlockMotionBiPred(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y,
                                    int blocktype, short pred_mv_x1, short pred_mv_y1, short pred_mv_x2, short pred_mv_y2,
                                    short *mv_x, short *mv_y, short *s_mv_x, short *s_mv_y,
                                    int search_range, int min_mcost, int lambda_factor);
int __wrap_FullPelBlockMotionBiPred(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y,
                                    int blocktype, short pred_mv_x1, short pred_mv_y1, short pred_mv_x2, short pred_mv_y2,
                                    short *mv_x, short *mv_y, short *s_mv_x, short *s_mv_y,
                                    int search_range, int min_mcost, int lambda_factor)
{
  int rec[REC_INTS] = {0};
  int list_offset = img->mb_data[img->current_mb_nr].list_offset;
  int cost;
  dump_current_if_needed();
  rec[0] = REC_BI_FULL; rec[1] = coded_pic_no; rec[2] = img->type; rec[3] = img->opix_x; rec[4] = img->opix_y;
  rec[5] = list; rec[6] = ref; rec[7] = blocktype; rec[8] = pic_pix_x; rec[9] = pic_pix_y;
  rec[10] = pred_mv_x1; rec[11] = pred_mv_y1; rec[12] = pred_mv_x2; rec[13] = pred_mv_y2;
  rec[14] = min_mcost; rec[15] = lambda_factor; rec[16] = search_range;
  rec[17] = *mv_x; rec[18] = *mv_y; rec[19] = *s_mv_x; rec[20] = *s_mv_y;
  rec[21] = ref_id_of(listX[list + list_offset][ref]);          /* picture read at the fixed vector s_mv */
  rec[22] = ref_id_of(listX[(list ^ 1) + list_offset][0]);      /* picture read at the searched candidate */
  rec[26] = ChromaMEEnable;
  if (active_pps->weighted_bipred_idc > 0) {
    /* what the function is about to derive (me_fullsearch.c:202-227): rec[29] = 1 + log denominator (0 = no weights),
     * rec[30] = weight1 | weight2 << 16, rec[31] = offsetBi */
    int o1 = list == 0 ? wp_offset[list_offset][ref][0] : wp_offset[list_offset + 1][0][ref];
    int o2 = list == 0 ? wp_offset[list_offset + 1][ref][0] : wp_offset[list_offset][0][ref];
    int w1 = list == 0 ? wbp_weight[list_offset][ref][0][0] : wbp_weight[list_offset + LIST_1][0][ref][0];
    int w2 = list == 0 ? wbp_weight[list_offset + LIST_1][ref][0][0] : wbp_weight[list_offset][0][ref][0];
    rec[29] = 1 + luma_log_weight_denom; rec[30] = (w1 & 0xffff) | (w2 << 16); rec[31] = (o1 + o2 + 1) >> 1;
    if (ChromaMEEnable) {
      /* weight1_cr / weight2_cr / offsetBi_cr (me_fullsearch.c:245-262), 8 bits each:
       * rec[27] = w1_cb | w2_cb << 8 | obi_cb << 16 | 1 << 24 (present), rec[28] = w1_cr | w2_cr << 8 | obi_cr << 16 */
      int k;
      for (k = 0; k < 2; k++) {
        int c1 = list == 0 ? wbp_weight[list_offset][ref][0][k + 1] : wbp_weight[list_offset + LIST_1][0][ref][k + 1];
        int c2 = list == 0 ? wbp_weight[list_offset + LIST_1][ref][0][k + 1] : wbp_weight[list_offset][0][ref][k + 1];
        int ob = list == 0 ? (wp_offset[list_offset][ref][k + 1] + wp_offset[list_offset + LIST_1][ref][k + 1] + 1) >> 1
                           : (wp_offset[list_offset + LIST_1][0][k + 1] + wp_offset[list_offset][0][k + 1] + 1) >> 1;
        rec[27 + k] = (c1 & 0xff) | ((c2 & 0xff) << 8) | ((ob & 0xff) << 16) | ((k == 0) << 24);
      }
    }
  }
  cost = __real_FullPelBlockMotionBiPred(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype,
                                         pred_mv_x1, pred_mv_y1, pred_mv_x2, pred_mv_y2,
                                         mv_x, mv_y, s_mv_x, s_mv_y, search_range, min_mcost, lambda_factor);
  rec[23] = *mv_x; rec[24] = *mv_y; rec[25] = cost;
  put(rec);
  return cost;
}

/* ---------------------------------------------------------------- bi-predictive sub-pel refinement */
int __real_SubPelBlockSearchBiPred(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y,
                                   int blocktype, short pred_mv_x, short pred_mv_y,
                                   short *mv_x, short *mv_y, short *s_mv_x, short *s_mv_y,
                                   int search_pos2, int search_pos4, int min_mcost, int *lambda_factor);
int __wrap_SubPelBlockSearchBiPred(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y,
                                   int blocktype, short pred_mv_x, short pred_mv_y,
                                   short *mv_x, short *mv_y, short *s_mv_x, short *s_mv_y,
                                   int search_pos2, int search_pos4, int min_mcost, int *lambda_factor)
{
  int rec[REC_INTS] = {0};
  int list_offset = img->mb_data[img->current_mb_nr].list_offset;
  int cost;
  dump_current_if_needed();
  rec[0] = REC_BI_SUB; rec[1] = coded_pic_no; rec[2] = img->type | (ChromaMEEnable << 8); rec[3] = img->opix_x; rec[4] = img->opix_y;
  rec[5] = list; rec[6] = ref; rec[7] = blocktype; rec[8] = pic_pix_x; rec[9] = pic_pix_y;
  rec[10] = pred_mv_x; rec[11] = pred_mv_y;
  rec[12] = *mv_x; rec[13] = *mv_y; rec[14] = *s_mv_x; rec[15] = *s_mv_y;
  rec[16] = min_mcost; rec[17] = lambda_factor[0]; rec[18] = lambda_factor[1]; rec[19] = lambda_factor[2];
  rec[21] = ref_id_of(listX[list + list_offset][ref]);
  rec[22] = ref_id_of(listX[(list ^ 1) + list_offset][0]);
  rec[26] = test8x8transform; rec[27] = start_me_refinement_hp; rec[28] = start_me_refinement_qp;
  if (active_pps->weighted_bipred_idc > 0) {
    /* me_fullsearch.c:541-542, :574-582; same packing as in the integer-pel record */
    int o1 = list == 0 ? wp_offset[list_offset][ref][0] : wp_offset[list_offset + 1][0][0];
    int o2 = list == 0 ? wp_offset[list_offset + 1][ref][0] : wp_offset[list_offset][0][0];
    int w1 = list == 0 ? wbp_weight[list_offset][ref][0][0] : wbp_weight[list_offset + LIST_1][0][ref][0];
    int w2 = list == 0 ? wbp_weight[list_offset + LIST_1][ref][0][0] : wbp_weight[list_offset][0][ref][0];
    rec[29] = 1 + luma_log_weight_denom; rec[30] = (w1 & 0xffff) | (w2 << 16); rec[31] = (o1 + o2 + 1) >> 1;
  }
  cost = __real_SubPelBlockSearchBiPred(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y,
                                        mv_x, mv_y, s_mv_x, s_mv_y, search_pos2, search_pos4, min_mcost, lambda_factor);
  rec[23] = *mv_x; rec[24] = *mv_y; rec[25] = cost;
  put(rec);
  return cost;
}

/* ---------------------------------------------------------------- sub-pel stages of the pattern searches (SearchMode 1, 2, 3)
 * record REC_SUBPEL_PAT: fields 0..21 and 26..31 as REC_SUBPEL (a function with one factor has it in all of 15..17);
 *   rec[22] = test8x8transform | start_me_refinement_hp << 1 | start_me_refinement_qp << 2 | rdopt << 3 | pattern << 4
 *             (pattern: 1 EPZS, 2 UMHexagonS, 3 simplified UMHexagonS, 4 simplified UMHexagonS "full")
 *   rec[23] = SubPelThreshold1 | SubPelThreshold3 << 16 (me_umhexsmp.c:106-107)
 *   rec[24] = subthres[blocktype] as EPZSInit derives it (me_epzs.c:335-354; the array itself is a file static)
 *   rec[25] = smpUMHEX_pred_MV_uplayer_X (low half) | smpUMHEX_pred_MV_uplayer_Y << 16 */
static int epzs_subthres(int blocktype)
{
  static const int medthres_base[8] = { 0, 256, 128, 128, 64, 32, 32, 16 };       /* me_epzs.c:41 */
  int pel_error_me = 1 << (img->bitdepth_luma - 8), pel_error_me_cr = 1 << (img->bitdepth_chroma - 8);
  double chroma_weight = input->ChromaMEEnable ? pel_error_me_cr * input->ChromaMEWeight * (double)(img->width_cr * img->height_cr) / (double)(img->width * img->height) : 0;
  return input->EPZSSubPelThresScale * (medthres_base[blocktype] * pel_error_me + (int)(medthres_base[blocktype] * chroma_weight + 0.5));
}

static void pat_head(int *rec, int pattern, short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype, int pmx, int pmy,
                     int mvx, int mvy, int sp2, int sp4, int min_mcost, int l0, int l1, int l2)
{
  int list_offset = img->mb_data[img->current_mb_nr].list_offset;
  dump_current_if_needed();
  rec[0] = REC_SUBPEL_PAT; rec[1] = coded_pic_no; rec[2] = img->type | (ChromaMEEnable << 8); rec[3] = img->opix_x; rec[4] = img->opix_y;
  rec[5] = list; rec[6] = ref; rec[7] = blocktype; rec[8] = pic_pix_x; rec[9] = pic_pix_y;
  rec[10] = pmx; rec[11] = pmy; rec[12] = mvx; rec[13] = mvy; rec[14] = min_mcost;
  rec[15] = l0; rec[16] = l1; rec[17] = l2;
  rec[21] = ref_id_of(listX[list + list_offset][ref]);
  rec[22] = test8x8transform | (start_me_refinement_hp << 1) | (start_me_refinement_qp << 2) | ((input->rdopt != 0) << 3) | (pattern << 4);
  rec[23] = SubPelThreshold1 | (SubPelThreshold3 << 16);
  rec[24] = epzs_subthres(blocktype);
  rec[25] = (smpUMHEX_pred_MV_uplayer_X & 0xffff) | (smpUMHEX_pred_MV_uplayer_Y << 16);
  rec[26] = sp2; rec[27] = sp4;
  put_weights(rec, list, list_offset, ref);
}

int __real_EPZSSubPelBlockMotionSearch(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                       short pred_mv[2], short mv[2], int search_pos2, int search_pos4, int min_mcost, int *lambda);
int __wrap_EPZSSubPelBlockMotionSearch(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                       short pred_mv[2], short mv[2], int search_pos2, int search_pos4, int min_mcost, int *lambda)
{
  int rec[REC_INTS] = {0}, cost;
  pat_head(rec, 1, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv[0], pred_mv[1], mv[0], mv[1], search_pos2, search_pos4, min_mcost,
           lambda[0], lambda[1], lambda[2]);
  cost = __real_EPZSSubPelBlockMotionSearch(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv, mv, search_pos2, search_pos4, min_mcost, lambda);
  rec[18] = mv[0]; rec[19] = mv[1]; rec[20] = cost;
  put(rec);
  return cost;
}

#define PAT_WRAP(NAME, PATTERN)                                                                                                     \
int __real_##NAME(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype, short pred_mv_x, short pred_mv_y, \
                  short *mv_x, short *mv_y, int search_pos2, int search_pos4, int min_mcost, int lambda_factor);                    \
int __wrap_##NAME(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype, short pred_mv_x, short pred_mv_y, \
                  short *mv_x, short *mv_y, int search_pos2, int search_pos4, int min_mcost, int lambda_factor)                     \
{                                                                                                                                   \
  int rec[REC_INTS] = {0}, cost;                                                                                                    \
  pat_head(rec, PATTERN, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y, *mv_x, *mv_y, search_pos2, search_pos4,  \
           min_mcost, lambda_factor, lambda_factor, lambda_factor);                                                                 \
  cost = __real_##NAME(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y, mv_x, mv_y, search_pos2,         \
                       search_pos4, min_mcost, lambda_factor);                                                                      \
  rec[18] = *mv_x; rec[19] = *mv_y; rec[20] = cost;                                                                                 \
  put(rec);                                                                                                                         \
  return cost;                                                                                                                      \
}
PAT_WRAP(UMHEXSubPelBlockMotionSearch, 2)
PAT_WRAP(smpUMHEXSubPelBlockMotionSearch, 3)
PAT_WRAP(smpUMHEXFullSubPelBlockMotionSearch, 4)

/* record REC_BI_SUB_PAT: as REC_BI_SUB, and rec[20] = second predictor, x (low half) | y << 16 */
int __real_EPZSSubPelBlockSearchBiPred(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                       short *pred_mv1, short *pred_mv2, short mv[2], short s_mv[2],
                                       int search_pos2, int search_pos4, int min_mcost, int *lambda);
int __wrap_EPZSSubPelBlockSearchBiPred(imgpel *orig_pic, short ref, int list, int pic_pix_x, int pic_pix_y, int blocktype,
                                       short *pred_mv1, short *pred_mv2, short mv[2], short s_mv[2],
                                       int search_pos2, int search_pos4, int min_mcost, int *lambda)
{
  int rec[REC_INTS] = {0};
  int list_offset = img->mb_data[img->current_mb_nr].list_offset;
  int cost;
  dump_current_if_needed();
  rec[0] = REC_BI_SUB_PAT; rec[1] = coded_pic_no; rec[2] = img->type | (ChromaMEEnable << 8); rec[3] = img->opix_x; rec[4] = img->opix_y;
  rec[5] = list; rec[6] = ref; rec[7] = blocktype; rec[8] = pic_pix_x; rec[9] = pic_pix_y;
  rec[10] = pred_mv1[0]; rec[11] = pred_mv1[1];
  rec[12] = mv[0]; rec[13] = mv[1]; rec[14] = s_mv[0]; rec[15] = s_mv[1];
  rec[16] = min_mcost; rec[17] = lambda[0]; rec[18] = lambda[1]; rec[19] = lambda[2];
  rec[20] = (pred_mv2[0] & 0xffff) | (pred_mv2[1] << 16);
  rec[21] = ref_id_of(listX[list + list_offset][ref]);
  rec[22] = ref_id_of(listX[(list ^ 1) + list_offset][0]);
  rec[26] = test8x8transform; rec[27] = start_me_refinement_hp; rec[28] = start_me_refinement_qp;
  if (active_pps->weighted_bipred_idc > 0) {
    /* me_epzs.c:2746-2747, :2783-2789; same packing as in the other bi-predictive records */
    int o1 = list == 0 ? wp_offset[list_offset][ref][0] : wp_offset[list_offset + 1][0][0];
    int o2 = list == 0 ? wp_offset[list_offset + 1][ref][0] : wp_offset[list_offset][0][0];
    int w1 = list == 0 ? wbp_weight[list_offset][ref][0][0] : wbp_weight[list_offset + LIST_1][0][ref][0];
    int w2 = list == 0 ? wbp_weight[list_offset + LIST_1][ref][0][0] : wbp_weight[list_offset][0][ref][0];
    rec[29] = 1 + luma_log_weight_denom; rec[30] = (w1 & 0xffff) | (w2 << 16); rec[31] = (o1 + o2 + 1) >> 1;
  }
  cost = __real_EPZSSubPelBlockSearchBiPred(orig_pic, ref, list, pic_pix_x, pic_pix_y, blocktype, pred_mv1, pred_mv2, mv, s_mv,
                                            search_pos2, search_pos4, min_mcost, lambda);
  rec[23] = mv[0]; rec[24] = mv[1]; rec[25] = cost;
  put(rec);
  return cost;
}
