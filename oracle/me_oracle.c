/*
 * me_oracle.c -- TEST INFRASTRUCTURE (see me_oracle.h).  A straight, slow, scalar restatement of what the
 * reference encoder computes on the motion-estimation hot path.  Written from the behaviour of the reference
 * (citations = reference file:line that each function follows); no reference code is reused.
 *
 * Parity pin: tests/test_oracle_golden.py (committed fixtures made from reference traces) and
 * tests/test_oracle_ref.py (live traces of oracle/_ref/lencod_trace when that binary is present).
 */
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include <stddef.h>
#include "me_oracle.h"

#define PAD ME_IMG_PAD

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* img->max_imgpel_value: 255, or (1 << bit depth) - 1 in the 16-bit build (ORA_PEL16; ora_set_max_pel) */
static int g_max_pel = 255;
void ora_set_max_pel(int v) { g_max_pel = v; }
static inline int iabs_(int v) { return v < 0 ? -v : v; }

/* ------------------------------------------------------------------------------------------------ tables */

/* partition p -> (block type, top-left 4x4 block index); order documented in include/me_cuda.h */
const int8_t ora_part_type[ME_NUM_PARTITIONS] = {
  1, 2, 2, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 6, 6,
  7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7 };
const int8_t ora_part_blk[ME_NUM_PARTITIONS] = {
  0, 0, 8, 0, 2, 0, 2, 8, 10, 0, 2, 4, 6, 8, 10, 12, 14, 0, 1, 2, 3, 8, 9, 10, 11,
  0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 };

/* input->blc_size[type][0..1] (configfile.c:805-841) */
void ora_block_size(int blocktype, int *bsx, int *bsy)
{
  static const int sx[8] = { 16, 16, 16, 8, 8, 8, 4, 4 };
  static const int sy[8] = { 16, 16, 8, 16, 8, 4, 8, 4 };
  *bsx = sx[blocktype];
  *bsy = sy[blocktype];
}

int ora_num_positions(int range) { return (2 * range + 1) * (2 * range + 1); }

/* spiral_search_x/y (mv-search.c:366-393): centre, then ring by ring; inside a ring first the top/bottom rows
 * (without corners) left to right alternating top, bottom; then the left/right columns top to bottom. */
void ora_spiral(int range, int16_t *sx, int16_t *sy)
{
  int k = 0, l, i;
  sx[k] = 0; sy[k] = 0; k++;
  for (l = 1; l <= (range > 1 ? range : 1); l++) {
    for (i = -l + 1; i < l; i++) {
      sx[k] = (int16_t)i;  sy[k] = (int16_t)-l; k++;
      sx[k] = (int16_t)i;  sy[k] = (int16_t)l;  k++;
    }
    for (i = -l; i <= l; i++) {
      sx[k] = (int16_t)-l; sy[k] = (int16_t)i;  k++;
      sx[k] = (int16_t)l;  sy[k] = (int16_t)i;  k++;
    }
  }
}

/* mvbits[] (mv-search.c:333-341): length of the signed Exp-Golomb code of d */
int ora_mvbits(int d)
{
  int a = iabs_(d), n = 0;
  if (a == 0) return 1;
  while (a >> (n + 1)) n++;          /* n = floor(log2 a) */
  return 2 * n + 3;
}

/* MV_COST_SMP (defines.h:126-128): 32-bit int product, arithmetic shift */
int ora_mv_cost(int f, int cx, int cy, int px, int py)
{
  return (f * (ora_mvbits(cx - px) + ora_mvbits(cy - py))) >> 16;
}

/* ------------------------------------------------------------------------------------------------ planes */

ora_ref *ora_ref_create(const ora_pel *luma, int stride, int w, int h)
{
  ora_ref *r = (ora_ref *)calloc(1, sizeof *r);
  int wp = w + 2 * PAD, hp = h + 2 * PAD, x, y, fy, fx;
  int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)wp * hp);
  ora_pel *G, *b, *hh, *j;
  r->w = w; r->h = h; r->wp = wp; r->hp = hp;
  for (fy = 0; fy < 4; fy++)
    for (fx = 0; fx < 4; fx++)
      r->plane[fy][fx] = (ora_pel *)malloc((size_t)wp * hp * sizeof(ora_pel));
  G = r->plane[0][0]; b = r->plane[0][2]; hh = r->plane[2][0]; j = r->plane[2][2];

#define AT(P, yy, xx) ((P)[(size_t)(yy) * wp + (xx)])
#define CX(v) clampi((v), 0, wp - 1)
#define CY(v) clampi((v), 0, hp - 1)
  /* integer plane with replicated border (getSubImageInteger, img_luma.c:120-166) */
  for (y = 0; y < hp; y++)
    for (x = 0; x < wp; x++)
      AT(G, y, x) = luma[(size_t)clampi(y - PAD, 0, h - 1) * stride + clampi(x - PAD, 0, w - 1)];
  /* horizontal half-pel: taps (1,-5,20,20,-5,1) over the padded integer plane, columns clamped to the padded
   * plane; the unrounded sum is kept for the centre position (getHorSubImageSixTap, img_luma.c:181-280) */
  for (y = 0; y < hp; y++)
    for (x = 0; x < wp; x++) {
      int s = 20 * (AT(G, y, x) + AT(G, y, CX(x + 1))) - 5 * (AT(G, y, CX(x - 1)) + AT(G, y, CX(x + 2)))
              + (AT(G, y, CX(x - 2)) + AT(G, y, CX(x + 3)));
      AT(tmp, y, x) = s;
      AT(b, y, x) = (ora_pel)clampi((s + 16) >> 5, 0, g_max_pel);
    }
  /* vertical half-pel (getVerSubImageSixTap, img_luma.c:293-380) */
  for (y = 0; y < hp; y++)
    for (x = 0; x < wp; x++) {
      int s = 20 * (AT(G, y, x) + AT(G, CY(y + 1), x)) - 5 * (AT(G, CY(y - 1), x) + AT(G, CY(y + 2), x))
              + (AT(G, CY(y - 2), x) + AT(G, CY(y + 3), x));
      AT(hh, y, x) = (ora_pel)clampi((s + 16) >> 5, 0, g_max_pel);
    }
  /* centre half-pel from the unrounded horizontal sums (getVerSubImageSixTapTmp, img_luma.c:390-470) */
  for (y = 0; y < hp; y++)
    for (x = 0; x < wp; x++) {
      int s = 20 * (AT(tmp, y, x) + AT(tmp, CY(y + 1), x)) - 5 * (AT(tmp, CY(y - 1), x) + AT(tmp, CY(y + 2), x))
              + (AT(tmp, CY(y - 2), x) + AT(tmp, CY(y + 3), x));
      AT(j, y, x) = (ora_pel)clampi((s + 512) >> 10, 0, g_max_pel);
    }
  /* quarter-pel positions: rounded averages of two neighbours (img_luma.c:483-639) */
  for (y = 0; y < hp; y++)
    for (x = 0; x < wp; x++) {
      int x1 = CX(x + 1), y1 = CY(y + 1);
      int g = AT(G, y, x), bb = AT(b, y, x), hv = AT(hh, y, x), jj = AT(j, y, x);
      AT(r->plane[0][1], y, x) = (ora_pel)((g + bb + 1) >> 1);
      AT(r->plane[1][0], y, x) = (ora_pel)((g + hv + 1) >> 1);
      AT(r->plane[1][1], y, x) = (ora_pel)((bb + hv + 1) >> 1);
      AT(r->plane[1][2], y, x) = (ora_pel)((bb + jj + 1) >> 1);
      AT(r->plane[2][1], y, x) = (ora_pel)((hv + jj + 1) >> 1);
      AT(r->plane[0][3], y, x) = (ora_pel)((bb + AT(G, y, x1) + 1) >> 1);
      AT(r->plane[1][3], y, x) = (ora_pel)((bb + AT(hh, y, x1) + 1) >> 1);
      AT(r->plane[2][3], y, x) = (ora_pel)((jj + AT(hh, y, x1) + 1) >> 1);
      AT(r->plane[3][0], y, x) = (ora_pel)((hv + AT(G, y1, x) + 1) >> 1);
      AT(r->plane[3][1], y, x) = (ora_pel)((hv + AT(b, y1, x) + 1) >> 1);
      AT(r->plane[3][2], y, x) = (ora_pel)((jj + AT(b, y1, x) + 1) >> 1);
      AT(r->plane[3][3], y, x) = (ora_pel)((AT(b, y1, x) + AT(hh, y, x1) + 1) >> 1);
    }
#undef AT
#undef CX
#undef CY
  free(tmp);
  return r;
}

void ora_ref_destroy(ora_ref *r)
{
  int fy, fx;
  if (!r) return;
  for (fy = 0; fy < 4; fy++)
    for (fx = 0; fx < 4; fx++) free(r->plane[fy][fx]);
  free(r->crec[0]); free(r->crec[1]);
  free(r);
}

const ora_pel *ora_ref_plane(const ora_ref *r, int fy, int fx) { return r->plane[fy][fx]; }

void ora_ref_set_weights(ora_ref *r, int apply, int weight, int offset, int log_denom)
{
  r->wp_apply = apply; r->wp_weight = weight; r->wp_offset = offset; r->wp_denom = log_denom;
}

/* the reference sample a uni-predictive distortion sees: as stored, or weighted_pel of me_fullfast.c:653 /
 * computeSADWP (me_distortion.c:413) / computeSATDWP (:734) with wp_luma_round = 1 << (denom - 1)
 * (weighted_prediction.c:50) and iClip1(255, .) */
static inline int refpel(const ora_ref *r, int v)
{
  if (!r->wp_apply) return v;
  return clampi(((r->wp_weight * v + (r->wp_denom ? 1 << (r->wp_denom - 1) : 0)) >> r->wp_denom) + r->wp_offset, 0, g_max_pel);
}

/* UMVLine4X (refbuf.c:37-43) with size_x_pad = w + 2*20 - 1 - 16 (mbuffer.c:421-422).  FastLine4X (refbuf.c:25)
 * yields the same pixels whenever the reference chooses it: it is only chosen where the clamp is a no-op or where
 * both addresses lie in the edge-replicated region in which every plane is constant along the row / column. */
const ora_pel *ora_umv_line(const ora_ref *r, int y4, int x4)
{
  int xp = clampi(x4 >> 2, 0, r->w + 2 * PAD - 1 - 16);
  int yp = clampi(y4 >> 2, 0, r->h + 2 * PAD - 1 - 16);
  return r->plane[y4 & 3][x4 & 3] + (size_t)yp * r->wp + xp;
}

/* ------------------------------------------------------------------------------------------------ chroma (ChromaMEEnable) */

static struct {
  int enable, weight, stride, shift_x, shift_y;
  const ora_pel *cur[2];
} g_cr;
/* weighted bi-prediction state (see ora_set_bipred_weights below); c*: the two chroma components */
static struct { int apply, w1, w2, obi, denom, cw1[2], cw2[2], cobi[2]; } g_bw[2], g_bw_cur;

void ora_set_chroma(int enable, int weight, const ora_pel *cur_u, const ora_pel *cur_v, int stride, int shift_x, int shift_y)
{
  g_cr.enable = enable; g_cr.weight = weight; g_cr.cur[0] = cur_u; g_cr.cur[1] = cur_v; g_cr.stride = stride;
  g_cr.shift_x = shift_x; g_cr.shift_y = shift_y;
}

void ora_ref_set_chroma(ora_ref *r, const ora_pel *u, const ora_pel *v, int stride, int cw, int ch)
{
  int k, y;
  for (k = 0; k < 2; k++) {
    const ora_pel *src = k ? v : u;
    free(r->crec[k]);
    r->crec[k] = (ora_pel *)malloc((size_t)cw * ch);
    for (y = 0; y < ch; y++) memcpy(r->crec[k] + (size_t)y * cw, src + (size_t)y * stride, cw);
  }
  r->cw = cw; r->ch = ch;
}

void ora_ref_set_chroma444(ora_ref *r, const ora_ref *u, const ora_ref *v)
{
  r->c444[0] = u; r->c444[1] = v;
  r->cw = r->w; r->ch = r->h;
}
static inline int has_chroma(const ora_ref *r) { return r->crec[0] != NULL || r->c444[0] != NULL; }

/* geometry of the chroma sub-images: padding = luma padding in chroma samples (img_pad_size_uv_x/y, lencod.c:2856-2879), macroblock
 * size in chroma samples (mb_cr_size_x/y, lencod.c:622-623) */
static inline int cr_pad_x(void) { return (4 * PAD) >> g_cr.shift_x; }
static inline int cr_pad_y(void) { return (4 * PAD) >> g_cr.shift_y; }
static inline int cr_mb_x(void) { return 64 >> g_cr.shift_x; }
static inline int cr_mb_y(void) { return 64 >> g_cr.shift_y; }

/* One sample of sub-image (sy, sx) at padded-plane position (ypos, xpos) of component k: getSubImagesChroma (img_chroma.c:374-440) blends
 * the four neighbours with eighth-sample weights, rshift_rnd_sf(., 6), over the edge-replicated reconstruction
 * (generateChroma00/01/10/XX, img_chroma.c:34-372: the padding rows / columns repeat the first / last row / column). */
static int cr_pel(const ora_ref *r, int k, int ypos, int xpos, int sy, int sx)
{
  if (r->c444[k]) return r->c444[k]->plane[sy][sx][(size_t)ypos * r->c444[k]->wp + xpos];     /* 4:4:4: a luma-style plane set */
  int jy = ypos - cr_pad_y(), ix = xpos - cr_pad_x();
  int y0 = clampi(jy, 0, r->ch - 1), y1 = clampi(jy + 1, 0, r->ch - 1), x0 = clampi(ix, 0, r->cw - 1), x1 = clampi(ix + 1, 0, r->cw - 1);
  int kk = sy << (3 - g_cr.shift_y), ll = sx << (3 - g_cr.shift_x);          /* vertical / horizontal offset in eighths */
  int w01 = (8 - kk) * ll, w00 = ((8 - kk) << 3) - w01, w11 = kk * ll, w10 = (kk << 3) - w11;
  const ora_pel *p = r->crec[k];
  return (w00 * p[(size_t)y0 * r->cw + x0] + w01 * p[(size_t)y0 * r->cw + x1] +
          w10 * p[(size_t)y1 * r->cw + x0] + w11 * p[(size_t)y1 * r->cw + x1] + 32) >> 6;
}

/* UMVLine8X_chroma (refbuf.c:53-59): sub-image from the fraction bits, position clamped to [0, size_cr - 1 + 2 * pad - mb_cr_size]
 * (mbuffer.c:425-426).  y, x: luma quarter-pel coordinates including the luma padding, as for ora_umv_line. */
static void cr_line(const ora_ref *r, int y, int x, int *ypos, int *xpos, int *sy, int *sx)
{
  *xpos = clampi(x >> g_cr.shift_x, 0, r->cw - 1 + 2 * cr_pad_x() - cr_mb_x());
  *ypos = clampi(y >> g_cr.shift_y, 0, r->ch - 1 + 2 * cr_pad_y() - cr_mb_y());
  *sy = y & ((1 << g_cr.shift_y) - 1);
  *sx = x & ((1 << g_cr.shift_x) - 1);
}

/* chroma part of computeSAD / computeSSE (me_distortion.c:376-402, :1069-1091) for the luma block of bsx x bsy at (pic_pix_x,
 * pic_pix_y): per component the SAD / SSE of the co-located chroma block, each multiplied by ChromaMEWeight.  r2 != NULL:
 * bi-prediction (p1 + p2 + 1) >> 1 (computeBiPredSAD1 :516-543). */
static int cr_block_dist(int sse, int pic_pix_x, int pic_pix_y, int bsx, int bsy, const ora_ref *r1, int x41, int y41,
                         const ora_ref *r2, int x42, int y42)
{
  int scx = g_cr.shift_x - 2, scy = g_cr.shift_y - 2;                        /* shift_cr_x/y, lencod.c:2881-2882 */
  int bx = bsx >> scx, by = bsy >> scy, cx0 = pic_pix_x >> scx, cy0 = pic_pix_y >> scy, k, x, y, total = 0;
  for (k = 0; k < 2; k++) {
    int yp1, xp1, sy1, sx1, yp2 = 0, xp2 = 0, sy2 = 0, sx2 = 0, m = 0;
    cr_line(r1, y41, x41, &yp1, &xp1, &sy1, &sx1);
    if (r2) cr_line(r2, y42, x42, &yp2, &xp2, &sy2, &sx2);
    for (y = 0; y < by; y++)
      for (x = 0; x < bx; x++) {
        int p = cr_pel(r1, k, yp1 + y, xp1 + x, sy1, sx1), d;
        if (r2) {
          int q = cr_pel(r2, k, yp2 + y, xp2 + x, sy2, sx2);
          /* computeBiPredSAD2 chroma (me_distortion.c:608-643): weight1_cr / weight2_cr / offsetBi_cr with the function's LUMA
           * denominator and rounding (:565-566) */
          p = g_bw_cur.apply ? clampi(((g_bw_cur.cw1[k] * p + g_bw_cur.cw2[k] * q + (g_bw_cur.denom ? 1 << g_bw_cur.denom : 0))
                                       >> (g_bw_cur.denom + 1)) + g_bw_cur.cobi[k], 0, 255)
                             : (p + q + 1) >> 1;
        }
        d = (int)g_cr.cur[k][(size_t)(cy0 + y) * g_cr.stride + cx0 + x] - p;
        m += sse ? d * d : iabs_(d);
      }
    total += g_cr.weight * m;
  }
  return total;
}

/* ------------------------------------------------------------------------------------------------ integer pel */

/* max_search_range[list][ref] (me_fullfast.c:127-140) */
int ora_setup_range(const ora_cfg *c, int ref)
{
  return (c->full_search == 2 || ref == 0) ? c->search_range : c->search_range / 2;
}

/* mv-search.c:1411-1416 */
int ora_block_search_range(const ora_cfg *c, int ref, int blocktype)
{
  int r1 = (ref < 1 ? ref : 1) + 1;
  if (c->full_search == 2) return c->search_range;
  if (c->full_search == 1) return c->search_range / r1;
  return c->search_range / (r1 * (blocktype < 2 ? blocktype : 2));
}

/* search centre (me_fullfast.c:550-566): predictor / 4 with C truncation, clips, plus macroblock origin */
void ora_search_centre(const ora_cfg *c, int ref, int opix_x, int opix_y, int pmv16_x, int pmv16_y,
                       int *centre_x, int *centre_y)
{
  int r = ora_setup_range(c, ref);
  int cx = pmv16_x / 4, cy = pmv16_y / 4;
  if (!c->rdopt) { cx = clampi(cx, -r, r); cy = clampi(cy, -r, r); }
  cx = clampi(cx, -2047 + r, 2047 - r);
  cy = clampi(cy, c->mv_limit_min_y + r, c->mv_limit_max_y - r);
  *centre_x = cx + opix_x;
  *centre_y = cy + opix_y;
}

/* SetupFastFullPelSearch main loop (me_fullfast.c:619-774, unweighted luma branch) + SetupLargerBlocks (:210-273).
 * dist_method is SAD for MEErrorMetric[0]==0 and squared error for anything else (me_fullfast.c:512). */
void ora_ffs_surfaces(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r,
                      int opix_x, int opix_y, int centre_x, int centre_y, int range, int32_t *surf)
{
  int P = ora_num_positions(range), pos, p, by, bx, y, x;
  int16_t *sx = (int16_t *)malloc(sizeof(int16_t) * P), *sy = (int16_t *)malloc(sizeof(int16_t) * P);
  int leaf[16];
  ora_spiral(range, sx, sy);
  for (pos = 0; pos < P; pos++) {
    const ora_pel *rp = ora_umv_line(r, (centre_y + sy[pos] + PAD) << 2, (centre_x + sx[pos] + PAD) << 2);
    memset(leaf, 0, sizeof leaf);
    for (y = 0; y < 16; y++)
      for (x = 0; x < 16; x++) {
        int d = refpel(r, rp[(size_t)y * r->wp + x]) - (int)cur[(size_t)(opix_y + y) * cur_stride + opix_x + x];
        leaf[(y >> 2) * 4 + (x >> 2)] += c->metric[0] ? d * d : iabs_(d);
      }
    if (g_cr.enable && has_chroma(r)) {
      /* me_fullfast.c:776-814: the co-located chroma block of every 4x4 luma block, both components, added as it is
       * (no ChromaMEWeight here) */
      int k, yp, xp, sy_, sx_, cy, cx, mbx = cr_mb_x(), mby = cr_mb_y();
      int ocx = (mbx * opix_x) >> 4, ocy = (mby * opix_y) >> 4;            /* img->opix_c_x/y, macroblock.c:206-207 */
      cr_line(r, (centre_y + sy[pos] + PAD) << 2, (centre_x + sx[pos] + PAD) << 2, &yp, &xp, &sy_, &sx_);
      for (k = 0; k < 2; k++)
        for (cy = 0; cy < mby; cy++)
          for (cx = 0; cx < mbx; cx++) {
            int d = cr_pel(r, k, yp + cy, xp + cx, sy_, sx_) - (int)g_cr.cur[k][(size_t)(ocy + cy) * g_cr.stride + ocx + cx];
            leaf[(cy / (mby / 4)) * 4 + cx / (mbx / 4)] += c->metric[0] ? d * d : iabs_(d);
          }
    }
    for (p = 0; p < ME_NUM_PARTITIONS; p++) {
      int bsx, bsy, s = 0, b = ora_part_blk[p];
      ora_block_size(ora_part_type[p], &bsx, &bsy);
      for (by = 0; by < bsy / 4; by++)
        for (bx = 0; bx < bsx / 4; bx++) s += leaf[b + 4 * by + bx];
      surf[(size_t)p * P + pos] = s;
    }
  }
  free(sx); free(sy);
}

/* FastFullPelBlockMotionSearch (me_fullfast.c:833-903).  surf_p: the surface of this partition.
 * Returns the minimum cost; mv_x, mv_y receive the integer-pel vector relative to the block. */
int ora_ffs_block(const ora_cfg *c, const int32_t *surf_p, int setup_range, int opix_x, int opix_y,
                  int centre_x, int centre_y, int pmv_x, int pmv_y, int search_range, int min_mcost,
                  int lambda_factor, short *mv_x, short *mv_y)
{
  int Pset = ora_num_positions(setup_range);
  int P = ora_num_positions(search_range);
  int offx = centre_x - opix_x, offy = centre_y - opix_y, pos, best = 0;
  int16_t *sx = (int16_t *)malloc(sizeof(int16_t) * Pset), *sy = (int16_t *)malloc(sizeof(int16_t) * Pset);
  ora_spiral(setup_range, sx, sy);
  if (P > Pset) P = Pset;            /* the reference would read past its arrays; no configuration does this */
  if (!c->rdopt) {
    /* the zero vector is scored first and keeps ties (me_fullfast.c:602-616, 867-876) */
    for (pos = 0; pos < Pset; pos++)
      if (sx[pos] == -offx && sy[pos] == -offy) break;
    if (pos < Pset) {
      int m = surf_p[pos] + ora_mv_cost(lambda_factor, 0, 0, pmv_x, pmv_y);
      if (m < min_mcost) { min_mcost = m; best = pos; }
    }
  }
  for (pos = 0; pos < P; pos++) {
    if (surf_p[pos] < min_mcost) {
      int m = surf_p[pos] + ora_mv_cost(lambda_factor, (offx + sx[pos]) << 2, (offy + sy[pos]) << 2, pmv_x, pmv_y);
      if (m < min_mcost) { min_mcost = m; best = pos; }
    }
  }
  *mv_x = (short)(offx + sx[best]);
  *mv_y = (short)(offy + sy[best]);
  free(sx); free(sy);
  return min_mcost;
}

/* ------------------------------------------------------------------------------------------------ distortion */

/* HadamardSAD4x4 (me_distortion.c:182-264): 2-D 4-point Hadamard, sum of magnitudes, (s+1)>>1.
 * The order of the butterflies does not affect the sum of magnitudes, so a plain separable form is used. */
int ora_hadamard4x4(const int *d)
{
  int t[16], i, s = 0;
  for (i = 0; i < 4; i++) {                       /* rows */
    int a = d[4 * i], b = d[4 * i + 1], c = d[4 * i + 2], e = d[4 * i + 3];
    int p = a + e, q = b + c, u = a - e, v = b - c;
    t[4 * i] = p + q; t[4 * i + 1] = p - q; t[4 * i + 2] = u + v; t[4 * i + 3] = u - v;
  }
  for (i = 0; i < 4; i++) {                       /* columns */
    int a = t[i], b = t[4 + i], c = t[8 + i], e = t[12 + i];
    int p = a + e, q = b + c, u = a - e, v = b - c;
    s += iabs_(p + q) + iabs_(p - q) + iabs_(u + v) + iabs_(u - v);
  }
  return (s + 1) >> 1;
}

/* HadamardSAD8x8 (me_distortion.c:272-343): (sum of magnitudes + 2) >> 2 */
int ora_hadamard8x8(const int *d)
{
  int m[64], i, k, s = 0, stage;
  memcpy(m, d, sizeof m);
  for (i = 0; i < 8; i++)                          /* rows: three butterfly stages */
    for (stage = 4; stage >= 1; stage >>= 1)
      for (k = 0; k < 8; k++)
        if (!(k & stage)) {
          int a = m[8 * i + k], b = m[8 * i + k + stage];
          m[8 * i + k] = a + b; m[8 * i + k + stage] = a - b;
        }
  for (i = 0; i < 8; i++)                          /* columns */
    for (stage = 4; stage >= 1; stage >>= 1)
      for (k = 0; k < 8; k++)
        if (!(k & stage)) {
          int a = m[8 * k + i], b = m[8 * (k + stage) + i];
          m[8 * k + i] = a + b; m[8 * (k + stage) + i] = a - b;
        }
  for (i = 0; i < 64; i++) s += iabs_(m[i]);
  return (s + 2) >> 2;
}

/* computeSAD (me_distortion.c:351-405), luma only.  The reference's early exits return partial sums that the
 * callers then discard (they are >= the running minimum), so the full sum is equivalent. */
int ora_sad_block(const ora_pel *src, int src_stride, int bsx, int bsy, const ora_ref *r, int x4, int y4)
{
  const ora_pel *rp = ora_umv_line(r, y4, x4);
  int x, y, s = 0;
  for (y = 0; y < bsy; y++)
    for (x = 0; x < bsx; x++) s += iabs_((int)src[(size_t)y * src_stride + x] - refpel(r, rp[(size_t)y * r->wp + x]));
  return s;
}

static int sse_block(const ora_pel *src, int src_stride, int bsx, int bsy, const ora_ref *r, int x4, int y4)
{
  const ora_pel *rp = ora_umv_line(r, y4, x4);
  int x, y, s = 0;
  for (y = 0; y < bsy; y++)
    for (x = 0; x < bsx; x++) {
      int d = (int)src[(size_t)y * src_stride + x] - refpel(r, rp[(size_t)y * r->wp + x]);
      s += d * d;
    }
  return s;
}

/* computeSATD (me_distortion.c:657-726): the reference fetches every 4x4 (or 8x8) tile with its own get_line call,
 * i.e. the origin clamp applies per tile. */
int ora_satd_block(const ora_pel *src, int src_stride, int bsx, int bsy, const ora_ref *r, int x4, int y4, int use8x8)
{
  int n = use8x8 ? 8 : 4, bx, by, x, y, s = 0, d[64];
  for (by = 0; by < bsy; by += n)
    for (bx = 0; bx < bsx; bx += n) {
      const ora_pel *rp = ora_umv_line(r, y4 + (by << 2), x4 + (bx << 2));
      for (y = 0; y < n; y++)
        for (x = 0; x < n; x++)
          d[y * n + x] = (int)src[(size_t)(by + y) * src_stride + bx + x] - refpel(r, rp[(size_t)y * r->wp + x]);
      s += use8x8 ? ora_hadamard8x8(d) : ora_hadamard4x4(d);
    }
  return s;
}

static int cr_block_dist(int sse, int pic_pix_x, int pic_pix_y, int bsx, int bsy, const ora_ref *r1, int x41, int y41,
                         const ora_ref *r2, int x42, int y42);

/* pic_pix_x / y: position of the block in the picture (for the chroma terms of computeSAD / computeSSE, me_distortion.c:376-402,
 * :1069-1091, when the global ChromaMEEnable is set; computeSATD has none) */
static int unipred(const ora_cfg *c, int level, const ora_pel *src, int stride, int bsx, int bsy,
                   const ora_ref *r, int x4, int y4, int use8x8, int pic_pix_x, int pic_pix_y)
{
  int chroma = (g_cr.enable && has_chroma(r) && c->metric[level] != 2)
               ? cr_block_dist(c->metric[level], pic_pix_x, pic_pix_y, bsx, bsy, r, x4, y4, NULL, 0, 0) : 0;
  switch (c->metric[level]) {       /* computeUniPred[] set-up, mv-search.c:400-424 */
    case 0:  return ora_sad_block(src, stride, bsx, bsy, r, x4, y4) + chroma;
    case 1:  return sse_block(src, stride, bsx, bsy, r, x4, y4) + chroma;
    default: return ora_satd_block(src, stride, bsx, bsy, r, x4, y4, use8x8);
  }
}

/* ------------------------------------------------------------------------------------------------ sub-pel */

/* sub-pel patterns are spiral entries 0..8 (search range 1): (0,0) (0,-1) (0,1) (-1,-1) (1,-1) (-1,0) (1,0) (-1,1) (1,1) */
static void first_spiral(int n, int16_t *sx, int16_t *sy)
{
  int16_t tx[9], ty[9];
  int i;
  ora_spiral(1, tx, ty);
  for (i = 0; i < n && i < 9; i++) { sx[i] = tx[i]; sy[i] = ty[i]; }
}

/* start_me_refinement_hp/qp (mv-search.c:396-397) with ChromaMEEnable == 0 */
/* start_me_refinement_hp / qp (mv-search.c:396-397) */
static int start_hp(const ora_cfg *c) { return (c->chroma_me == 1 || c->metric[0] != c->metric[1]) ? 0 : 1; }
static int start_qp(const ora_cfg *c) { return (c->chroma_me == 1 || c->metric[1] != c->metric[2]) ? 0 : 1; }

/* SubPelBlockMotionSearch (me_fullsearch.c:341-511) */
int ora_subpel(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r,
               int pic_pix_x, int pic_pix_y, int blocktype, int ref, int slice_type,
               int pmv_x, int pmv_y, short *mv_x, short *mv_y, int search_pos2, int search_pos4,
               int min_mcost, const int *lambda)
{
  int bsx, bsy, pos, best, mcost, cx, cy;
  int use8x8 = c->transform8x8 && blocktype <= 4;                 /* test8x8transform, mv-search.c:640 */
  int check0 = (!c->rdopt && slice_type != ME_B_SLICE && ref == 0 && blocktype == 1 && *mv_x == 0 && *mv_y == 0);
  int p4x = (pic_pix_x + PAD) << 2, p4y = (pic_pix_y + PAD) << 2;
  int shp = start_hp(c), sqp = start_qp(c);
  int max_pos2 = (!shp && search_pos2 < 1) ? 1 : search_pos2;
  const ora_pel *src = cur + (size_t)pic_pix_y * cur_stride + pic_pix_x;
  int16_t sx[9], sy[9];
  ora_block_size(blocktype, &bsx, &bsy);
  first_spiral(9, sx, sy);

  for (best = 0, pos = shp; pos < max_pos2; pos++) {              /* half-pel: spiral_hpel_search = 2 * spiral */
    cx = *mv_x + 2 * sx[pos]; cy = *mv_y + 2 * sy[pos];
    mcost = ora_mv_cost(lambda[1], cx, cy, pmv_x, pmv_y);
    if (mcost >= min_mcost) continue;
    mcost += unipred(c, 1, src, cur_stride, bsx, bsy, r, cx + p4x, cy + p4y, use8x8, pic_pix_x, pic_pix_y);
    if (pos == 0 && check0) mcost -= (lambda[1] * 16) >> 16;      /* WEIGHTED_COST(lambda, 16), :439-442 */
    if (mcost < min_mcost) { min_mcost = mcost; best = pos; }
  }
  if (best) { *mv_x += 2 * sx[best]; *mv_y += 2 * sy[best]; }
  if (!sqp) min_mcost = INT_MAX;

  for (best = 0, pos = sqp; pos < search_pos4; pos++) {           /* quarter-pel */
    cx = *mv_x + sx[pos]; cy = *mv_y + sy[pos];
    mcost = ora_mv_cost(lambda[2], cx, cy, pmv_x, pmv_y);
    if (mcost >= min_mcost) continue;
    mcost += unipred(c, 2, src, cur_stride, bsx, bsy, r, cx + p4x, cy + p4y, use8x8, pic_pix_x, pic_pix_y);
    if (mcost < min_mcost) { min_mcost = mcost; best = pos; }
  }
  if (best) { *mv_x += sx[best]; *mv_y += sy[best]; }
  return min_mcost;
}

/* The block distortions with the reference's early exits (computeSAD / SSE: after every row when the sum has reached the
 * threshold, me_distortion.c:366,1066; computeSATD: after every tile when the sum exceeds it, :689).  For a threshold below
 * the full distortion the partial sum comes back.  Everywhere else on this path the caller discards such a candidate, so
 * the full sum is equivalent; FullPelBlockMotionSearch is the exception (see below). */
static int unipred_exit(const ora_cfg *c, const ora_pel *src, int stride, int bsx, int bsy, const ora_ref *r, int x4, int y4,
                        int use8x8, int thr, int pic_pix_x, int pic_pix_y)
{
  int metric = c->metric[0], m = 0, x, y;
  if (metric == 2) {
    int n = use8x8 ? 8 : 4, bx, by;
    for (by = 0; by < bsy; by += n)
      for (bx = 0; bx < bsx; bx += n) {
        m += ora_satd_block(src + (size_t)by * stride + bx, stride, n, n, r, x4 + (bx << 2), y4 + (by << 2), use8x8);
        if (m > thr) return m;
      }
    return m;
  } else {
    const ora_pel *rp = ora_umv_line(r, y4, x4);
    for (y = 0; y < bsy; y++) {
      for (x = 0; x < bsx; x++) {
        int d = (int)src[(size_t)y * stride + x] - refpel(r, rp[(size_t)y * r->wp + x]);
        m += metric ? d * d : iabs_(d);
      }
      if (m >= thr) return m;
    }
    if (g_cr.enable && has_chroma(r)) {
      /* chroma terms (me_distortion.c:376-402): one component after the other, each followed by the same early exit */
      int scx = g_cr.shift_x - 2, scy = g_cr.shift_y - 2, k;
      int bx = bsx >> scx, by = bsy >> scy, cx0 = pic_pix_x >> scx, cy0 = pic_pix_y >> scy;
      for (k = 0; k < 2; k++) {
        int yp, xp, sy_, sx_, mc = 0;
        cr_line(r, y4, x4, &yp, &xp, &sy_, &sx_);
        for (y = 0; y < by; y++)
          for (x = 0; x < bx; x++) {
            int d = (int)g_cr.cur[k][(size_t)(cy0 + y) * g_cr.stride + cx0 + x] - cr_pel(r, k, yp + y, xp + x, sy_, sx_);
            mc += metric ? d * d : iabs_(d);
          }
        m += g_cr.weight * mc;
        if (m >= thr) return m;
      }
    }
    return m;
  }
}

/* FullPelBlockMotionSearch (me_fullsearch.c:47-155): the per-block spiral full search of SearchMode = -1 (and of every
 * SearchMode value the dispatch of mv-search.c:629-768 does not know).  mv in: search centre (integer pel, set by the caller at
 * mv-search.c:751-762), out: best vector.  Distortion = computeUniPred[F_PEL (+3 with weights)]: SAD, SSE or SATD of the whole
 * block.  The zero-vector bonus compares a quarter-pel candidate with integer-pel block coordinates (:129) -- kept as is.
 * The bonus is subtracted BEFORE the distortion call, whose early-exit threshold is min_mcost - mcost (:138): with the
 * caller's min_mcost = INT_MAX and a negative mcost that difference wraps to a negative int, the distortion function leaves
 * after its first row (first tile for SATD) and the candidate is accepted with that partial sum.  This happens for the first
 * spiral position of blocks at picture origin under !rdopt, and the encoder's decisions depend on it -- reproduced. */
int ora_fullpel_block(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r,
                      int pic_pix_x, int pic_pix_y, int blocktype, int ref, int slice_type,
                      int pmv_x, int pmv_y, short *mv_x, short *mv_y, int search_range, int min_mcost, int lambda_factor)
{
  int P = ora_num_positions(search_range), pos, best = 0, bsx, bsy;
  int16_t *sx = (int16_t *)malloc(sizeof(int16_t) * P), *sy = (int16_t *)malloc(sizeof(int16_t) * P);
  int pred_x = (pic_pix_x << 2) + pmv_x, pred_y = (pic_pix_y << 2) + pmv_y;
  int center_x = pic_pix_x + *mv_x, center_y = pic_pix_y + *mv_y;
  int check_for_00 = (blocktype == 1 && !c->rdopt && slice_type != ME_B_SLICE && ref == 0);
  int use8x8 = c->transform8x8 && blocktype <= 4;
  const ora_pel *src = cur + (size_t)pic_pix_y * cur_stride + pic_pix_x;
  ora_block_size(blocktype, &bsx, &bsy);
  ora_spiral(search_range, sx, sy);
  for (pos = 0; pos < P; pos++) {
    int cand_x = (center_x + sx[pos]) << 2, cand_y = (center_y + sy[pos]) << 2;
    int mcost = ora_mv_cost(lambda_factor, cand_x, cand_y, pred_x, pred_y);
    if (check_for_00 && cand_x == pic_pix_x && cand_y == pic_pix_y) mcost -= (lambda_factor * 16) >> 16;
    if (mcost >= min_mcost) continue;
    mcost += unipred_exit(c, src, cur_stride, bsx, bsy, r, cand_x + 4 * PAD, cand_y + 4 * PAD, use8x8,
                          (int)((unsigned)min_mcost - (unsigned)mcost), pic_pix_x, pic_pix_y);
    if (mcost < min_mcost) { best = pos; min_mcost = mcost; }
  }
  if (best) { *mv_x += sx[best]; *mv_y += sy[best]; }
  free(sx); free(sy);
  return min_mcost;
}

/* ------------------------------------------------------------------------------------------------ bi-pred */

/* Weighted bi-prediction: the state FullPelBlockMotionBiPred / SubPelBlockSearchBiPred derive per call from their `list`
 * argument (me_fullsearch.c:202-234, :538-590) -- weight1 (picture of `list`, read at the fixed vector), weight2 (reference 0
 * of the other list, read at the candidates), offsetBi, luma_log_weight_denom.  One set per value of `list`; the set in
 * use is selected before a call (the composite of mv-search.c:864-1034 switches it as the lists take turns). */
/* (g_bw, g_bw_cur: defined with the chroma state above) */

void ora_set_bipred_weights(int list, int apply, int weight1, int weight2, int offset_bi, int log_denom)
{
  g_bw[list & 1].apply = apply; g_bw[list & 1].w1 = weight1; g_bw[list & 1].w2 = weight2;
  g_bw[list & 1].obi = offset_bi; g_bw[list & 1].denom = log_denom;
  /* chroma: the default weights (plain average) until ora_set_bipred_chroma_weights */
  g_bw[list & 1].cw1[0] = g_bw[list & 1].cw1[1] = g_bw[list & 1].cw2[0] = g_bw[list & 1].cw2[1] = 1 << log_denom;
  g_bw[list & 1].cobi[0] = g_bw[list & 1].cobi[1] = 0;
}
/* weight1_cr[k], weight2_cr[k], offsetBi_cr[k] of me_fullsearch.c:245-262 */
void ora_set_bipred_chroma_weights(int list, int k, int weight1, int weight2, int offset_bi)
{
  g_bw[list & 1].cw1[k & 1] = weight1; g_bw[list & 1].cw2[k & 1] = weight2; g_bw[list & 1].cobi[k & 1] = offset_bi;
}
void ora_select_bipred_list(int list) { g_bw_cur = g_bw[list & 1]; }

/* prediction sample: (p1 + p2 + 1) >> 1 (computeBiPredSAD1, me_distortion.c:502), or with weights
 * iClip1(255, ((weight1 * p1 + weight2 * p2 + 2 * wp_luma_round) >> (denom + 1)) + offsetBi) (computeBiPredSAD2 :556-640,
 * computeBiPredSATD2 :907) */
static inline int bipel(int p1, int p2)
{
  if (!g_bw_cur.apply) return (p1 + p2 + 1) >> 1;
  return clampi(((g_bw_cur.w1 * p1 + g_bw_cur.w2 * p2 + 2 * (g_bw_cur.denom ? 1 << (g_bw_cur.denom - 1) : 0)) >> (g_bw_cur.denom + 1))
                + g_bw_cur.obi, 0, g_max_pel);
}

/* computeBiPredSAD1 / computeBiPredSATD1 (me_distortion.c:482-548, 824-899): prediction = (p1 + p2 + 1) >> 1.
 * SAD fetches each picture once per block; SATD once per transform tile. */
static int bipred_dist(const ora_cfg *c, int level, const ora_pel *src, int stride, int bsx, int bsy,
                       const ora_ref *r1, int x41, int y41, const ora_ref *r2, int x42, int y42, int use8x8,
                       int pic_pix_x, int pic_pix_y)
{
  int metric = c->metric[level], s = 0, x, y, bx, by, d[64];
  if (metric == 0 || metric == 1) {
    const ora_pel *a = ora_umv_line(r1, y41, x41), *b = ora_umv_line(r2, y42, x42);
    for (y = 0; y < bsy; y++)
      for (x = 0; x < bsx; x++) {
        int e = (int)src[(size_t)y * stride + x] - bipel(a[(size_t)y * r1->wp + x], b[(size_t)y * r2->wp + x]);
        s += metric ? e * e : iabs_(e);
      }
    if (g_cr.enable && has_chroma(r1) && has_chroma(r2))     /* computeBiPredSAD1 :516-543 (weighted chroma is not restated) */
      s += cr_block_dist(metric, pic_pix_x, pic_pix_y, bsx, bsy, r1, x41, y41, r2, x42, y42);
    return s;
  } else {
    int n = use8x8 ? 8 : 4;
    for (by = 0; by < bsy; by += n)
      for (bx = 0; bx < bsx; bx += n) {
        const ora_pel *a = ora_umv_line(r1, y41 + (by << 2), x41 + (bx << 2));
        const ora_pel *b = ora_umv_line(r2, y42 + (by << 2), x42 + (bx << 2));
        for (y = 0; y < n; y++)
          for (x = 0; x < n; x++)
            if (use8x8 && g_bw_cur.apply) {
              /* computeBiPredSATD2, 8x8 branch (me_distortion.c:966-1023): the eighth sample of a row is read without
               * advancing src_line (:1015), so every row of the ORIGINAL block starts one sample further left than the one
               * above it: linear index by*bsx + bx + y*(bsx-1) + x into the bsx-wide copy of the block (mv-search.c:607-611) */
              int L = by * bsx + bx + y * (bsx - 1) + x;
              d[y * n + x] = (int)src[(size_t)(L / bsx) * stride + (L % bsx)] - bipel(a[(size_t)y * r1->wp + x], b[(size_t)y * r2->wp + x]);
            } else
            d[y * n + x] = (int)src[(size_t)(by + y) * stride + bx + x] - bipel(a[(size_t)y * r1->wp + x], b[(size_t)y * r2->wp + x]);
        s += use8x8 ? ora_hadamard8x8(d) : ora_hadamard4x4(d);
      }
    return s;
  }
}

/* FullPelBlockMotionBiPred (me_fullsearch.c:164-332), unweighted.  r1 = listX[list][ref] read at the fixed vector
 * (s_mv), r2 = listX[list^1][0] read at the candidate.  Vectors are integer pel.  Note the reference stores the
 * centres in `short` (:196-199). */
int ora_bipred_fullpel(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r1, const ora_ref *r2,
                       int pic_pix_x, int pic_pix_y, int blocktype,
                       int pmv_x1, int pmv_y1, int pmv_x2, int pmv_y2,
                       short *mv_x, short *mv_y, const short *s_mv_x, const short *s_mv_y,
                       int search_range, int min_mcost, int lambda_factor)
{
  int P = ora_num_positions(search_range), pos, best = 0, bsx, bsy;
  int16_t *sx = (int16_t *)malloc(sizeof(int16_t) * (P > 9 ? P : 9)), *sy = (int16_t *)malloc(sizeof(int16_t) * (P > 9 ? P : 9));
  int px1 = (pic_pix_x << 2) + pmv_x1, py1 = (pic_pix_y << 2) + pmv_y1;
  int px2 = (pic_pix_x << 2) + pmv_x2, py2 = (pic_pix_y << 2) + pmv_y2;
  short cx0 = (short)(pic_pix_x + *mv_x), cy0 = (short)(pic_pix_y + *mv_y);
  short fx0 = (short)(pic_pix_x + *s_mv_x), fy0 = (short)(pic_pix_y + *s_mv_y);
  const ora_pel *src = cur + (size_t)pic_pix_y * cur_stride + pic_pix_x;
  ora_block_size(blocktype, &bsx, &bsy);
  ora_spiral(search_range, sx, sy);
  for (pos = 0; pos < P; pos++) {
    int cx = (cx0 + sx[pos]) << 2, cy = (cy0 + sy[pos]) << 2;
    int mcost = ora_mv_cost(lambda_factor, fx0 << 2, fy0 << 2, px1, py1) + ora_mv_cost(lambda_factor, cx, cy, px2, py2);
    if (mcost >= min_mcost) continue;
    mcost += bipred_dist(c, 0, src, cur_stride, bsx, bsy, r1, (fx0 << 2) + 4 * PAD, (fy0 << 2) + 4 * PAD,
                         r2, cx + 4 * PAD, cy + 4 * PAD, c->transform8x8 && blocktype <= 4 /* test8x8transform, mv-search.c:640 */,
                         pic_pix_x, pic_pix_y);
    if (mcost < min_mcost) { min_mcost = mcost; best = pos; }
  }
  if (best) { *mv_x += sx[best]; *mv_y += sy[best]; }
  free(sx); free(sy);
  return min_mcost;
}

/* SubPelBlockSearchBiPred (me_fullsearch.c:520-754), unweighted.  Vectors are quarter pel; only the searched
 * vector carries a rate term (:666, :729). */
int ora_bipred_subpel(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r1, const ora_ref *r2,
                      int pic_pix_x, int pic_pix_y, int blocktype, int pmv_x, int pmv_y,
                      short *mv_x, short *mv_y, const short *s_mv_x, const short *s_mv_y,
                      int search_pos2, int search_pos4, int min_mcost, const int *lambda)
{
  int bsx, bsy, pos, best, mcost, cx, cy;
  int use8x8 = c->transform8x8 && blocktype <= 4;
  int p4x = (pic_pix_x + PAD) << 2, p4y = (pic_pix_y + PAD) << 2;
  int smx = *s_mv_x + p4x, smy = *s_mv_y + p4y;
  int shp = start_hp(c), sqp = start_qp(c);
  int max_pos2 = (!shp && search_pos2 < 1) ? 1 : search_pos2;
  const ora_pel *src = cur + (size_t)pic_pix_y * cur_stride + pic_pix_x;
  int16_t sx[9], sy[9];
  ora_block_size(blocktype, &bsx, &bsy);
  first_spiral(9, sx, sy);

  for (best = 0, pos = shp; pos < max_pos2; pos++) {
    cx = *mv_x + 2 * sx[pos]; cy = *mv_y + 2 * sy[pos];
    mcost = ora_mv_cost(lambda[1], cx, cy, pmv_x, pmv_y);
    if (mcost >= min_mcost) continue;
    mcost += bipred_dist(c, 1, src, cur_stride, bsx, bsy, r1, smx, smy, r2, cx + p4x, cy + p4y, use8x8, pic_pix_x, pic_pix_y);
    if (mcost < min_mcost) { min_mcost = mcost; best = pos; }
  }
  if (best) { *mv_x += 2 * sx[best]; *mv_y += 2 * sy[best]; }
  if (!sqp) min_mcost = INT_MAX;
  for (best = 0, pos = sqp; pos < search_pos4; pos++) {
    cx = *mv_x + sx[pos]; cy = *mv_y + sy[pos];
    mcost = ora_mv_cost(lambda[2], cx, cy, pmv_x, pmv_y);
    if (mcost >= min_mcost) continue;
    mcost += bipred_dist(c, 2, src, cur_stride, bsx, bsy, r1, smx, smy, r2, cx + p4x, cy + p4y, use8x8, pic_pix_x, pic_pix_y);
    if (mcost < min_mcost) { min_mcost = mcost; best = pos; }
  }
  if (best) { *mv_x += sx[best]; *mv_y += sy[best]; }
  return min_mcost;
}

/* ------------------------------------------------------------------------------------------------ batch job */

/* PartitionMotionSearch -> BlockMotionSearch (mv-search.c:1378, :560) for the 41 partitions of one
 * (macroblock, list, ref): integer-pel search (:745), mv <<= 2 (:770-774), min_mcost reset when the half-pel
 * metric differs (:785-788), sub-pel refinement with 9/9 positions (:823). */
void ora_mb_search(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r,
                   const me_mb_job *job, me_mb_result *out, me_mb_result_int *out_int)
{
  int range = ora_setup_range(c, job->ref), P = ora_num_positions(range), p, cx, cy;
  int32_t *surf = (int32_t *)malloc(sizeof(int32_t) * (size_t)ME_NUM_PARTITIONS * P);
  ora_search_centre(c, job->ref, job->opix_x, job->opix_y, job->pmv16[0], job->pmv16[1], &cx, &cy);
  ora_ffs_surfaces(c, cur, cur_stride, r, job->opix_x, job->opix_y, cx, cy, range, surf);
  for (p = 0; p < ME_NUM_PARTITIONS; p++) {
    int t = ora_part_type[p], b = ora_part_blk[p];
    int ppx = job->opix_x + 4 * (b & 3), ppy = job->opix_y + 4 * (b >> 2);
    short mx, my;
    int cost = ora_ffs_block(c, surf + (size_t)p * P, range, job->opix_x, job->opix_y, cx, cy,
                             job->pmv[p][0], job->pmv[p][1], ora_block_search_range(c, job->ref, t),
                             INT_MAX, job->lambda[0], &mx, &my);
    if (out_int) { out_int->mv[p][0] = mx; out_int->mv[p][1] = my; out_int->cost[p] = cost; }
    mx = (short)(mx << 2); my = (short)(my << 2);
    if (!start_hp(c)) cost = INT_MAX;
    {
      /* from the sub-pel stage on BlockMotionSearch runs with ChromaMEEnable = (input->ChromaMEEnable == 2) (mv-search.c:779) */
      int chroma_int = g_cr.enable;
      g_cr.enable = chroma_int && c->chroma_me == 2;
      cost = ora_subpel(c, cur, cur_stride, r, ppx, ppy, t, job->ref, job->slice_type,
                        job->pmv[p][0], job->pmv[p][1], &mx, &my, 9, 9, cost, job->lambda);
      g_cr.enable = chroma_int;
    }
    out->mv[p][0] = mx; out->mv[p][1] = my; out->cost[p] = cost;
  }
  free(surf);
}

/* The bi-predictive refinement that BlockMotionSearch runs for block type 1 / reference 0 of a B slice
 * (mv-search.c:864-1034, SearchMode 0): vectors back to integer pel ((mv + 2) >> 2, :884-885); 1 + refinements
 * integer-pel searches in which the two lists take turns (:889-975: odd passes search the vector of `list` with the
 * other one fixed, the cost carries over, the range is BiPredMESearchRange >> i); back to quarter pel (:976-982); one
 * sub-pel refinement of bimv with mv fixed (:984-1001) and, for BiPredMESubPel == 2, one of mv with bimv fixed
 * (:1004-1021).  The vectors are kept in `short` like the reference does. */
void ora_mb_bipred(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *l0, const ora_ref *l1,
                   const me_bipred_job *job, int refinements, int search_range, int subpel, me_bipred_result *out)
{
  const ora_ref *refs[2] = { l0, l1 };
  const int list = job->list;
  short mv[2] = { job->mv[0], job->mv[1] }, bimv[2] = { 0, 0 }, tempmv[2] = { 0, 0 };
  const int16_t *pred_mv = job->pmv, *pred_mv_bi = job->pmv_bi, *pred_mv1 = NULL, *pred_mv2 = NULL;
  int min_mcostbi = INT_MAX, iterlist = list, i;
  mv[0] = (short)((mv[0] + 2) >> 2);
  mv[1] = (short)((mv[1] + 2) >> 2);
  for (i = 0; i <= refinements; i++) {
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
    ora_select_bipred_list(iterlist);
    min_mcostbi = ora_bipred_fullpel(c, cur, cur_stride, refs[iterlist], refs[iterlist ^ 1], job->opix_x, job->opix_y, 1,
                                     pred_mv1[0], pred_mv1[1], pred_mv2[0], pred_mv2[1],
                                     &bimv[0], &bimv[1], &tempmv[0], &tempmv[1], search_range >> i, min_mcostbi, job->lambda[0]);
    mv[0] = tempmv[0]; mv[1] = tempmv[1];
  }
  mv[0] = (short)(tempmv[0] << 2); mv[1] = (short)(tempmv[1] << 2);
  bimv[0] = (short)(bimv[0] << 2); bimv[1] = (short)(bimv[1] << 2);
  if (subpel) {
    if (!start_hp(c)) min_mcostbi = INT_MAX;
    ora_select_bipred_list(iterlist);
    min_mcostbi = ora_bipred_subpel(c, cur, cur_stride, refs[iterlist], refs[iterlist ^ 1], job->opix_x, job->opix_y, 1,
                                    pred_mv2[0], pred_mv2[1], &bimv[0], &bimv[1], &mv[0], &mv[1], 9, 9, min_mcostbi, job->lambda);
  }
  if (subpel == 2) {
    if (!start_hp(c) || !start_qp(c)) min_mcostbi = INT_MAX;
    ora_select_bipred_list(iterlist ^ 1);
    min_mcostbi = ora_bipred_subpel(c, cur, cur_stride, refs[iterlist ^ 1], refs[iterlist], job->opix_x, job->opix_y, 1,
                                    pred_mv1[0], pred_mv1[1], &mv[0], &mv[1], &bimv[0], &bimv[1], 9, 9, min_mcostbi, job->lambda);
  }
  out->mv[0] = mv[0]; out->mv[1] = mv[1];
  out->bimv[0] = bimv[0]; out->bimv[1] = bimv[1];
  out->cost = min_mcostbi;
  out->iterlist = iterlist;
}

#ifndef ORA_PEL16   /* motion-compensated prediction, mode-decision costs and deblocking: restated for 8-bit pictures only */
/* ------------------------------------------------------------------------------------------------ prediction / mode-decision costs */

/* OneComponentLumaPrediction (macroblock.c:807-827): the block's rows start at UMVLine4X(plane set, y, x) (refbuf.c:37) and follow
 * at the padded pitch.  The picture's own samples: the weighted-ME copy (refpel) plays no part in motion compensation. */
static void one_component_luma_prediction(uint8_t *mpred, int pic_pix_x, int pic_pix_y, int bsx, int bsy, const ora_ref *r)
{
  const uint8_t *line = ora_umv_line(r, pic_pix_y, pic_pix_x);
  int j;
  for (j = 0; j < bsy; j++) { memcpy(mpred, line, (size_t)bsx); line += r->wp; mpred += bsx; }
}

/* LumaPrediction (macroblock.c:832-930), block of bsx x bsy at b->pix_x / pix_y; pred: bsy rows of bsx samples (img->mpr) */
void ora_luma_prediction(const ora_ref *const *list0, const ora_ref *const *list1, const me_pred_block *b, int bsx, int bsy, uint8_t *pred)
{
  uint8_t l0[256], l1[256];
  const int pic_opix_x = ((b->pix_x) << 2) + 4 * PAD, pic_opix_y = ((b->pix_y) << 2) + 4 * PAD;      /* macroblock.c:850-851 */
  int k, n = bsx * bsy;
  memset(l0, 0, sizeof l0); memset(l1, 0, sizeof l1);
  if (b->p_dir == 0 || b->p_dir == 2)
    one_component_luma_prediction(l0, pic_opix_x + b->mv[0][0], pic_opix_y + b->mv[0][1], bsx, bsy, list0[b->ref[0]]);
  if (b->p_dir == 1 || b->p_dir == 2)
    one_component_luma_prediction(l1, pic_opix_x + b->mv[1][0], pic_opix_y + b->mv[1][1], bsx, bsy, list1[b->ref[1]]);
  if (b->apply_weights) {
    const int wp_luma_round = 1 << (b->log_denom - 1);                       /* weighted_prediction.c:50 (denominators >= 1) */
    if (b->p_dir == 2) {                                                     /* macroblock.c:884-895 */
      const int wp_round = 2 * wp_luma_round, weight_denom = b->log_denom + 1;
      for (k = 0; k < n; k++) pred[k] = (uint8_t)clampi(((b->weight[0] * l0[k] + b->weight[1] * l1[k] + wp_round) >> weight_denom) + b->offset, 0, 255);
    } else {                                                                 /* :897-913 */
      const uint8_t *l = b->p_dir ? l1 : l0;
      for (k = 0; k < n; k++) pred[k] = (uint8_t)clampi(((b->weight[0] * l[k] + wp_luma_round) >> b->log_denom) + b->offset, 0, 255);
    }
  } else if (b->p_dir == 2) {                                                /* :917-922 */
    for (k = 0; k < n; k++) pred[k] = (uint8_t)((l0[k] + l1[k] + 1) >> 1);
  } else memcpy(pred, b->p_dir ? l1 : l0, (size_t)n);
}

/* distortion4x4 / distortion8x8 (me_distortion.c:76-103, :110-137): metric = input->ModeDecisionMetric */
int ora_distortion4x4(int metric, const int *diff)
{
  int k, s = 0;
  if (metric == 2) return ora_hadamard4x4(diff);
  for (k = 0; k < 16; k++) s += metric == 0 ? iabs_(diff[k]) : diff[k] * diff[k];
  return s;
}
int ora_distortion8x8(int metric, const int *diff)
{
  int k, s = 0;
  if (metric == 2) return ora_hadamard8x8(diff);
  for (k = 0; k < 64; k++) s += metric == 0 ? iabs_(diff[k]) : diff[k] * diff[k];
  return s;
}

/* the state the three cost functions read from the encoder's globals, as plain arrays: blocks[4 * by + bx] describes what
 * LumaPrediction does for the 4x4 block (bx, by) of the macroblock (mode's vectors, direction, references, weights) */
static void diff4x4(const uint8_t *cur, int cur_stride, const ora_ref *const *l0, const ora_ref *const *l1, const me_pred_block *b, int *d /*16*/)
{
  uint8_t mpr[16];
  int i, j;
  ora_luma_prediction(l0, l1, b, 4, 4, mpr);
  for (j = 0; j < 4; j++)
    for (i = 0; i < 4; i++) d[4 * j + i] = (int)cur[(size_t)(b->pix_y + j) * cur_stride + b->pix_x + i] - mpr[4 * j + i];
}

/* GetSkipCostMB (mv-search.c:1136-1181) */
int ora_get_skip_cost_mb(const uint8_t *cur, int cur_stride, const ora_ref *const *l0, const ora_ref *const *l1,
                         const me_pred_block *blocks /* 16 */, int rdopt, int transform8x8, int metric)
{
  int block, block_y, block_x, i, j, cost = 0, d[16], curr_diff[8][8], diff64[64];
  for (block = 0; block < 4; block++) {
    const int mb_y = (block >> 1) << 3, mb_x = (block & 1) << 3;
    for (block_y = mb_y; block_y < mb_y + 8; block_y += 4)
      for (block_x = mb_x; block_x < mb_x + 8; block_x += 4) {
        diff4x4(cur, cur_stride, l0, l1, &blocks[(block_y >> 2) * 4 + (block_x >> 2)], d);
        for (j = 0; j < 4; j++) for (i = 0; i < 4; i++) curr_diff[block_y - mb_y + j][block_x - mb_x + i] = d[4 * j + i];
        if (!((rdopt == 0) && transform8x8)) cost += ora_distortion4x4(metric, d);
      }
    if ((rdopt == 0) && transform8x8) {
      for (j = 0; j < 8; j++) memcpy(&diff64[8 * j], curr_diff[j], 8 * sizeof(int));
      cost += ora_distortion8x8(metric, diff64);
    }
  }
  return cost;
}

/* GetDirectCost8x8 (mv-search.c:1273-1325); a block with p_dir < 0 is direct_pdir < 0 (mode not allowed) */
int ora_get_direct_cost8x8(const uint8_t *cur, int cur_stride, const ora_ref *const *l0, const ora_ref *const *l1,
                           const me_pred_block *blocks /* 16 */, int block, int rdopt, int transform8x8, int metric, int *cost8x8)
{
  const int mb_y = (block / 2) << 3, mb_x = (block % 2) << 3;
  int block_y, block_x, i, j, cost = 0, d[16], curr_diff[8][8], diff64[64];
  for (block_y = mb_y; block_y < mb_y + 8; block_y += 4)
    for (block_x = mb_x; block_x < mb_x + 8; block_x += 4) {
      const me_pred_block *b = &blocks[(block_y >> 2) * 4 + (block_x >> 2)];
      if (b->p_dir < 0) { *cost8x8 = 2147483647; return 2147483647; }
      diff4x4(cur, cur_stride, l0, l1, b, d);
      for (j = 0; j < 4; j++) for (i = 0; i < 4; i++) curr_diff[block_y - mb_y + j][block_x - mb_x + i] = d[4 * j + i];
      cost += ora_distortion4x4(metric, d);
    }
  if ((rdopt == 0) && transform8x8) {
    for (j = 0; j < 8; j++) memcpy(&diff64[8 * j], curr_diff[j], 8 * sizeof(int));
    *cost8x8 += ora_distortion8x8(metric, diff64);
  }
  return cost;
}

/* BIDPartitionCost (mv-search.c:1052-1128): blocks[] carry p_dir 2 and the vectors all_mv[v][h][list][ref][blocktype];
 * pmv[4 * v + h][list] = img->pred_mv[v][h][list][ref][blocktype] */
int ora_bid_partition_cost(const uint8_t *cur, int cur_stride, const ora_ref *const *l0, const ora_ref *const *l1,
                           const me_pred_block *blocks /* 16 */, const int16_t (*pmv)[2][2] /* 16 */, int blocktype, int block8x8,
                           int lambda_factor, int transform8x8, int metric)
{
  static const int bx0[5][4] = { {0,0,0,0}, {0,0,0,0}, {0,0,0,0}, {0,2,0,0}, {0,2,0,2} };
  static const int by0[5][4] = { {0,0,0,0}, {0,0,0,0}, {0,2,0,0}, {0,0,0,0}, {0,0,2,2} };
  static const int blc_size[8][2] = { {0,0}, {16,16}, {16,8}, {8,16}, {8,8}, {8,4}, {4,8}, {4,4} };        /* configfile.c:805-841 */
  int curr_blk[16][16], diff64[64], d[16];
  const int bsx = blc_size[blocktype][0] < 8 ? blc_size[blocktype][0] : 8, bsy = blc_size[blocktype][1] < 8 ? blc_size[blocktype][1] : 8;
  const int parttype = blocktype < 4 ? blocktype : 4;
  const int step_h0 = blc_size[parttype][0] >> 2, step_v0 = blc_size[parttype][1] >> 2;                     /* input->part_size */
  const int step_h = blc_size[blocktype][0] >> 2, step_v = blc_size[blocktype][1] >> 2;
  const int bx = bx0[parttype][block8x8], by = by0[parttype][block8x8];
  int v, h, i, j, k, bxx, byy, mvd_bits = 0, mcost;
  for (v = by; v < by + step_v0; v += step_v)
    for (h = bx; h < bx + step_h0; h += step_h) {
      const me_pred_block *b = &blocks[4 * v + h];
      mvd_bits += ora_mvbits(b->mv[0][0] - pmv[4 * v + h][0][0]) + ora_mvbits(b->mv[0][1] - pmv[4 * v + h][0][1]);
      mvd_bits += ora_mvbits(b->mv[1][0] - pmv[4 * v + h][1][0]) + ora_mvbits(b->mv[1][1] - pmv[4 * v + h][1][1]);
    }
  mcost = (lambda_factor * mvd_bits) >> 16;                                                                 /* WEIGHTED_COST, defines.h:125 */
  for (byy = 0, v = by; v < by + step_v0; byy += 4, v++)
    for (bxx = 0, h = bx; h < bx + step_h0; bxx += 4, h++) {
      diff4x4(cur, cur_stride, l0, l1, &blocks[4 * v + h], d);
      for (j = 0; j < 4; j++) for (i = 0; i < 4; i++) curr_blk[byy + j][bxx + i] = d[4 * j + i];
      if (!transform8x8 || blocktype > 4) mcost += ora_distortion4x4(metric, d);
    }
  if (transform8x8 && blocktype <= 4)
    for (byy = 0; byy < blc_size[parttype][1]; byy += bsy)
      for (bxx = 0; bxx < blc_size[parttype][0]; bxx += bsx) {
        for (k = 0, j = byy; j < byy + 8; j++, k += 8) memcpy(&diff64[k], &curr_blk[j][bxx], 8 * sizeof(int));
        mcost += ora_distortion8x8(metric, diff64);
      }
  return mcost;
}

/* ------------------------------------------------------------------------------------------------ deblocking */

static const uint8_t DB_ALPHA[52] = { 0,0,0,0,0,0,0,0,0,0,0,0, 0,0,0,0,4,4,5,6, 7,8,9,10,12,13,15,17, 20,22,25,28,32,36,40,45, 50,56,63,71,80,90,101,113,
                                      127,144,162,182,203,226,255,255 };                                   /* loopFilter.c:42 */
static const uint8_t DB_BETA[52] = { 0,0,0,0,0,0,0,0,0,0,0,0, 0,0,0,0,2,2,2,3, 3,3,3,4,4,4,6,6, 7,7,8,8,9,9,10,10, 11,11,12,12,13,13,14,14,
                                     15,15,16,16,17,17,18,18 };                                            /* loopFilter.c:43 */
static const uint8_t DB_CLIP[52][5] = {                                                                     /* loopFilter.c:44-53 */
  {0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},
  {0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},{0,0,0,0,0},
  {0,0,0,0,0},{0,0,0,1,1},{0,0,0,1,1},{0,0,0,1,1},{0,0,0,1,1},{0,0,1,1,1},{0,0,1,1,1},{0,1,1,1,1},
  {0,1,1,1,1},{0,1,1,1,1},{0,1,1,1,1},{0,1,1,2,2},{0,1,1,2,2},{0,1,1,2,2},{0,1,1,2,2},{0,1,2,3,3},
  {0,1,2,3,3},{0,2,2,3,3},{0,2,2,4,4},{0,2,3,4,4},{0,2,3,4,4},{0,3,3,5,5},{0,3,4,6,6},{0,3,4,6,6},
  {0,4,5,7,7},{0,4,5,8,8},{0,4,6,9,9},{0,5,7,10,10},{0,6,8,11,11},{0,6,8,13,13},{0,7,10,14,14},{0,8,11,16,16},
  {0,9,12,18,18},{0,10,13,20,20},{0,11,15,23,23},{0,13,17,25,25} };

/* EdgeLoopLumaNormal (loopFilter.c:529-672) for one edge of 16 lines: q0 = first sample on the Q side of line 0, `step` from line to
 * line, `inc` across the edge (1 for a vertical edge, the pitch for a horizontal one) */
static void edge_loop_luma(uint8_t *q0, ptrdiff_t step, ptrdiff_t inc, const uint8_t *strength, int qp_p, int qp_q, int alpha_off, int beta_off)
{
  const int QP = (qp_p + qp_q + 1) >> 1;
  const int indexA = clampi(QP + alpha_off, 0, 51), indexB = clampi(QP + beta_off, 0, 51);
  const int Alpha = DB_ALPHA[indexA], Beta = DB_BETA[indexB];
  int pel;
  for (pel = 0; pel < 16; pel++) {
    uint8_t *Q = q0 + pel * step, *P = Q - inc;
    const int Strng = strength[pel];
    int L3, L2, L1, L0, R0, R1, R2, R3, Delta, AbsDelta, C0;
    if (!Strng) continue;
    L3 = P[-inc * 3]; L2 = P[-inc * 2]; L1 = P[-inc]; L0 = P[0]; R0 = Q[0]; R1 = Q[inc]; R2 = Q[inc * 2]; R3 = Q[inc * 3];
    Delta = R0 - L0; AbsDelta = iabs_(Delta);
    if (AbsDelta >= Alpha) continue;
    C0 = DB_CLIP[indexA][Strng];
    if (!(((iabs_(R0 - R1) - Beta) & (iabs_(L0 - L1) - Beta)) < 0)) continue;
    if (Strng == 4) {
      const int RL0 = L0 + R0, small_gap = AbsDelta < ((Alpha >> 2) + 2);
      const int aq = (iabs_(R0 - R2) < Beta) & small_gap, ap = (iabs_(L0 - L2) < Beta) & small_gap;
      if (ap) { P[-inc * 2] = (uint8_t)((((L3 + L2) << 1) + L2 + L1 + RL0 + 4) >> 3); P[-inc] = (uint8_t)((L2 + L1 + L0 + R0 + 2) >> 2);
                P[0] = (uint8_t)((R1 + ((L1 + RL0) << 1) + L2 + 4) >> 3); }
      else P[0] = (uint8_t)(((L1 << 1) + L0 + R1 + 2) >> 2);
      if (aq) { Q[0] = (uint8_t)((L1 + ((R1 + RL0) << 1) + R2 + 4) >> 3); Q[inc] = (uint8_t)((R2 + R0 + R1 + L0 + 2) >> 2);
                Q[inc * 2] = (uint8_t)((((R3 + R2) << 1) + R2 + R1 + RL0 + 4) >> 3); }
      else Q[0] = (uint8_t)(((R1 << 1) + R0 + L1 + 2) >> 2);
    } else {
      const int RL0 = (L0 + R0 + 1) >> 1, aq = iabs_(R0 - R2) < Beta, ap = iabs_(L0 - L2) < Beta;
      const int c0 = C0 + ap + aq, dif = clampi(((Delta << 2) + (L1 - R1) + 4) >> 3, -c0, c0);
      if (ap) P[-inc] = (uint8_t)(L1 + clampi((L2 + RL0 - (L1 << 1)) >> 1, -C0, C0));
      P[0] = (uint8_t)clampi(L0 + dif, 0, 255);
      Q[0] = (uint8_t)clampi(R0 - dif, 0, 255);
      if (aq) Q[inc] = (uint8_t)(R1 + clampi((R2 + RL0 - (R1 << 1)) >> 1, -C0, C0));
    }
  }
}

/* EdgeLoopChromaNormal (loopFilter.c:814-921), 4:2:0: eight lines, strength index ((pel >> 1) << 2) + (pel & 1) */
static void edge_loop_chroma(uint8_t *q0, ptrdiff_t step, ptrdiff_t inc, const uint8_t *strength, int qpc_p, int qpc_q, int alpha_off, int beta_off)
{
  const int QP = (qpc_p + qpc_q + 1) >> 1;
  const int indexA = clampi(QP + alpha_off, 0, 51), indexB = clampi(QP + beta_off, 0, 51);
  const int Alpha = DB_ALPHA[indexA], Beta = DB_BETA[indexB];
  int pel;
  for (pel = 0; pel < 8; pel++) {
    uint8_t *Q = q0 + pel * step, *P = Q - inc;
    const int Strng = strength[((pel >> 1) << 2) + (pel & 1)];
    int L1, L0, R0, R1, Delta;
    if (!Strng) continue;
    L1 = P[-inc]; L0 = P[0]; R0 = Q[0]; R1 = Q[inc];
    Delta = R0 - L0;
    if (iabs_(Delta) >= Alpha) continue;
    if (!(((iabs_(R0 - R1) - Beta) & (iabs_(L0 - L1) - Beta)) < 0)) continue;
    if (Strng == 4) { Q[0] = (uint8_t)(((R1 << 1) + R0 + L1 + 2) >> 2); P[0] = (uint8_t)(((L1 << 1) + L0 + R1 + 2) >> 2); }
    else {
      const int c0 = DB_CLIP[indexA][Strng] + 1, dif = clampi(((Delta << 2) + (L1 - R1) + 4) >> 3, -c0, c0);
      P[0] = (uint8_t)clampi(L0 + dif, 0, 255);
      Q[0] = (uint8_t)clampi(R0 - dif, 0, 255);
    }
  }
}

/* DeblockFrame / DeblockMb (loopFilter.c:87-253) with the boundary strengths given (me_deblock_mb): macroblocks in raster order,
 * vertical edges then horizontal ones, luma then the two chroma components per edge.  u, v NULL: 4:0:0 */
void ora_deblock_frame(uint8_t *y, int ystride, uint8_t *u, uint8_t *v, int cstride, int w, int h, const me_deblock_mb *mbs)
{
  const int mbw = w / 16, mbh = h / 16;
  int my, mx, dir, e, k;
  for (my = 0; my < mbh; my++)
    for (mx = 0; mx < mbw; mx++) {
      const me_deblock_mb *Q = &mbs[my * mbw + mx];
      for (dir = 0; dir < 2; dir++)
        for (e = 0; e < 4; e++) {
          const uint8_t *s = Q->strength[dir][e];
          const me_deblock_mb *P = Q;
          int sum = 0;
          for (k = 0; k < 16; k++) sum += s[k];
          if (!sum) continue;                                                       /* loopFilter.c:187-193 */
          if (e == 0 && (dir ? my > 0 : mx > 0)) P = dir ? Q - mbw : Q - 1;           /* MbP: the neighbour across a macroblock edge */
          if (!((Q->luma_skip >> (4 * dir + e)) & 1)) {
            if (dir) edge_loop_luma(y + (size_t)(16 * my + 4 * e) * ystride + 16 * mx, 1, ystride, s, P->qp, Q->qp, Q->alpha_off, Q->beta_off);
            else edge_loop_luma(y + (size_t)(16 * my) * ystride + 16 * mx + 4 * e, ystride, 1, s, P->qp, Q->qp, Q->alpha_off, Q->beta_off);
          }
          if (u && v && !(e & 1))                                                   /* chroma_edge[dir][edge][YUV420] = 0, -4, 4, -4 */
            for (k = 0; k < 2; k++) {
              uint8_t *c = k ? v : u;
              if (dir) edge_loop_chroma(c + (size_t)(8 * my + 2 * e) * cstride + 8 * mx, 1, cstride, s, P->qpc[k], Q->qpc[k], Q->alpha_off, Q->beta_off);
              else edge_loop_chroma(c + (size_t)(8 * my) * cstride + 8 * mx + 2 * e, cstride, 1, s, P->qpc[k], Q->qpc[k], Q->alpha_off, Q->beta_off);
            }
        }
    }
}

#endif /* !ORA_PEL16 */

/* ------------------------------------------------------------------------------------------------ sub-pel stages of the pattern searches
 * (SURVEY.md 8f row 4).  Every candidate is scored with the FULL distortion: the thresholds the reference hands to computeUniPred /
 * computeBiPred on these paths only ever cut a candidate that the comparison after the call rejects anyway (a partial sum above the
 * threshold comes back, the full sum is at least as large), and position 0 with check_position0 is scored against INT_MAX
 * (mv-search.c:785-788).  FAST_ACCESS / UMV_ACCESS: a centre inside the margin the reference tests keeps every candidate of the
 * pattern inside the padded plane, where the clamp of UMVLine4X is the identity -- so the clamped access serves both. */

typedef struct pat_ctx {
  const ora_cfg *c; const ora_pel *src; int stride, bsx, bsy, use8x8, ppx, ppy, p4x, p4y;
  const ora_ref *r1, *r2;      /* r2 != NULL: bi-predictive, r1 read at the fixed vector (smx, smy), r2 at the candidate */
  int smx, smy, pmvx, pmvy, fixed_rate_lambda_level, s_mv_x, s_mv_y, pmv2x, pmv2y;
} pat_ctx;

/* motion-vector rate + distortion of the candidate (cx, cy) (quarter-pel vector) at sub-pel level `level` with factor `lam` */
static int pat_cost(const pat_ctx *k, int level, int lam, int cx, int cy)
{
  int m = ora_mv_cost(lam, cx, cy, k->pmvx, k->pmvy);
  if (k->r2) {
    m += ora_mv_cost(lam, k->s_mv_x, k->s_mv_y, k->pmv2x, k->pmv2y);            /* me_epzs.c:2867, :3021: the fixed vector's rate too */
    return m + bipred_dist(k->c, level, k->src, k->stride, k->bsx, k->bsy, k->r1, k->smx, k->smy, k->r2, cx + k->p4x, cy + k->p4y,
                           k->use8x8, k->ppx, k->ppy);
  }
  return m + unipred(k->c, level, k->src, k->stride, k->bsx, k->bsy, k->r1, cx + k->p4x, cy + k->p4y, k->use8x8, k->ppx, k->ppy);
}

/* search_point_qp (me_epzs.c:44); search_point_hp (:43) is twice that.  Entry 9 repeats entry 5. */
static const int8_t epzs_pt[10][2] = { {0, 0}, {-1, 0}, {0, 1}, {1, 0}, {0, -1}, {-1, 1}, {1, 1}, {1, -1}, {-1, -1}, {-1, 1} };

/* One level of EPZSSubPelBlockMotionSearch / EPZSSubPelBlockSearchBiPred (me_epzs.c:2470-2580 = :2603-2712, :2858-2960 = :3014-3118):
 * the four diamond points (and the centre from `first`), then the corner(s) the best two of them point at.
 * hp: the half-pel level (its `case 0` of the second switch evaluates nothing; the quarter-pel level has that case commented out).
 * early_thr >= 0: the uni-predictive early return (:2496, :2629); returns 1 when it was taken. */
static int epzs_level(const pat_ctx *k, int level, int lam, int step, int hp, int first, int end_default, int early_thr,
                      short *mv_x, short *mv_y, int *min_mcost)
{
  int pos, best = 0, second = 0, second_cost = INT_MAX, mcost, start = 5, end = end_default;
  for (pos = first; pos < 5; pos++) {
    mcost = pat_cost(k, level, lam, *mv_x + step * epzs_pt[pos][0], *mv_y + step * epzs_pt[pos][1]);
    if (mcost < *min_mcost) { second_cost = *min_mcost; second = best; *min_mcost = mcost; best = pos; }
    else if (mcost < second_cost) { second_cost = mcost; second = pos; }
  }
  if (early_thr >= 0 && best == 0 && k->pmvx == *mv_x && k->pmvy - *mv_y == 0 && *min_mcost < early_thr) return 1;
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
  if (best != 0 || (iabs_(k->pmvx - *mv_x) + iabs_(k->pmvy - *mv_y)))
    for (pos = start; pos < end; pos++) {
      mcost = pat_cost(k, level, lam, *mv_x + step * epzs_pt[pos][0], *mv_y + step * epzs_pt[pos][1]);
      if (mcost < *min_mcost) { *min_mcost = mcost; best = pos; }
    }
  if (best) { *mv_x += step * epzs_pt[best][0]; *mv_y += step * epzs_pt[best][1]; }
  return 0;
}

static void pat_init(pat_ctx *k, const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r1, const ora_ref *r2,
                     int pic_pix_x, int pic_pix_y, int blocktype, int pmv_x, int pmv_y)
{
  memset(k, 0, sizeof *k);
  k->c = c; k->stride = cur_stride; k->src = cur + (size_t)pic_pix_y * cur_stride + pic_pix_x;
  ora_block_size(blocktype, &k->bsx, &k->bsy);
  k->use8x8 = c->transform8x8 && blocktype <= 4;                 /* test8x8transform, mv-search.c:640 */
  k->ppx = pic_pix_x; k->ppy = pic_pix_y; k->p4x = (pic_pix_x + PAD) << 2; k->p4y = (pic_pix_y + PAD) << 2;
  k->r1 = r1; k->r2 = r2; k->pmvx = pmv_x; k->pmvy = pmv_y;
}

/* EPZSSubPelBlockMotionSearch (me_epzs.c:2390-2720).  subthres = subthres[blocktype] (:354) */
int ora_epzs_subpel(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r,
                    int pic_pix_x, int pic_pix_y, int blocktype, int pmv_x, int pmv_y, short *mv_x, short *mv_y,
                    int search_pos2, int search_pos4, int min_mcost, const int *lambda, int subthres)
{
  pat_ctx k;
  int shp = start_hp(c), sqp = start_qp(c);
  int max_pos2 = ((!shp || !sqp) && search_pos2 < 1) ? 1 : search_pos2;          /* :2411 */
  pat_init(&k, c, cur, cur_stride, r, NULL, pic_pix_x, pic_pix_y, blocktype, pmv_x, pmv_y);
  if (epzs_level(&k, 1, lambda[1], 2, 1, shp, max_pos2, subthres, mv_x, mv_y, &min_mcost)) return min_mcost;
  if (!sqp) min_mcost = INT_MAX;                                                  /* :2581 */
  epzs_level(&k, 2, lambda[2], 1, 0, sqp, search_pos4, subthres, mv_x, mv_y, &min_mcost);
  return min_mcost;
}

/* EPZSSubPelBlockSearchBiPred (me_epzs.c:2728-3125): r1 = listX[list][ref] read at the fixed vector s_mv, r2 = listX[list ^ 1][0] read
 * at the candidate; both vectors carry a rate term; no early return (:2894 is commented out in the reference) */
int ora_epzs_subpel_bipred(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r1, const ora_ref *r2,
                           int pic_pix_x, int pic_pix_y, int blocktype, int pmv1_x, int pmv1_y, int pmv2_x, int pmv2_y,
                           short *mv_x, short *mv_y, const short *s_mv_x, const short *s_mv_y,
                           int search_pos2, int search_pos4, int min_mcost, const int *lambda)
{
  pat_ctx k;
  int shp = (min_mcost == INT_MAX) ? 0 : start_hp(c), sqp = start_qp(c);         /* :2760 */
  int max_pos2 = ((!start_hp(c) || !sqp) && search_pos2 < 1) ? 1 : search_pos2;   /* :2761 */
  pat_init(&k, c, cur, cur_stride, r1, r2, pic_pix_x, pic_pix_y, blocktype, pmv1_x, pmv1_y);
  k.smx = *s_mv_x + k.p4x; k.smy = *s_mv_y + k.p4y; k.s_mv_x = *s_mv_x; k.s_mv_y = *s_mv_y; k.pmv2x = pmv2_x; k.pmv2y = pmv2_y;
  epzs_level(&k, 1, lambda[1], 2, 1, shp, max_pos2, -1, mv_x, mv_y, &min_mcost);
  if (!sqp) min_mcost = INT_MAX;                                                  /* :3008 */
  epzs_level(&k, 2, lambda[2], 1, 0, sqp, search_pos4, -1, mv_x, mv_y, &min_mcost);
  return min_mcost;
}

/* The quarter-pel diamond walks of UMHEXSubPelBlockMotionSearch (me_umhex.c:562-730) and smpUMHEXSubPelBlockMotionSearch
 * (me_umhexsmp.c:616-812): inside +-3 quarter pel of the start vector, every position scored at most once (SearchState), metric
 * and factor of the quarter-pel level throughout.
 * smp == 0: diamond order (-1,0) (0,1) (1,0) (0,-1); no thresholds.
 * smp == 1: diamond order (-1,0) (1,0) (0,-1) (0,1); uplayer_x / y = smpUMHEX_pred_MV_uplayer_X / _Y (:1203-1214), thr1 / thr3 =
 * SubPelThreshold1 / 3 >> block_type_shift_factor[blocktype] (:106-107, :46). */
static int umhex_walk(const pat_ctx *k, int smp, int lam, int uplayer_x, int uplayer_y, int thr1, int thr3,
                      short *mv_x, short *mv_y, int min_mcost)
{
  static const int8_t dia[2][4][2] = { { {-1, 0}, {0, 1}, {1, 0}, {0, -1} }, { {-1, 0}, {1, 0}, {0, -1}, {0, 1} } };
  unsigned char seen[7][7];
  const int R = 3, mx0 = *mv_x, my0 = *mv_y;
  int cx, cy, mcost, curx = 0, cury = 0, i, m, abort_search;
  int fx = (k->pmvx - mx0) % 4, fy = (k->pmvy - my0) % 4;                         /* C remainder: sign of the dividend */
  int ux = (uplayer_x - mx0) % 4, uy = (uplayer_y - my0) % 4;
  memset(seen, 0, sizeof seen);
  seen[R][R] = 1;
  if (!start_hp(k->c)) {
    mcost = pat_cost(k, 2, lam, mx0, my0);
    if (mcost < min_mcost) { min_mcost = mcost; curx = mx0; cury = my0; }         /* otherwise the walk starts from vector (0, 0), as in the reference */
  } else { curx = mx0; cury = my0; }
  if (smp && mx0 == 0 && my0 == 0 && fx == 0 && ux == 0 && fy == 0 && uy == 0 && min_mcost < thr1) {
    *mv_x = (short)curx; *mv_y = (short)cury;
    return min_mcost;
  }
  if (fx || fy) {
    cx = mx0 + fx; cy = my0 + fy;
    mcost = pat_cost(k, 2, lam, cx, cy);
    seen[cy - my0 + R][cx - mx0 + R] = 1;
    if (mcost < min_mcost) { min_mcost = mcost; curx = cx; cury = cy; }
  }
  for (i = 0; i < R; i++) {
    int nx = curx, ny = cury;
    abort_search = 1;
    for (m = 0; m < 4; m++) {
      cx = nx + dia[smp][m][0]; cy = ny + dia[smp][m][1];
      if (iabs_(cx - mx0) <= R && iabs_(cy - my0) <= R && !seen[cy - my0 + R][cx - mx0 + R]) {
        mcost = pat_cost(k, 2, lam, cx, cy);
        seen[cy - my0 + R][cx - mx0 + R] = 1;
        if (mcost < min_mcost) { min_mcost = mcost; curx = cx; cury = cy; abort_search = 0; }
        if (smp && min_mcost < thr3) { *mv_x = (short)curx; *mv_y = (short)cury; return min_mcost; }
      }
    }
    if (abort_search) break;
  }
  *mv_x = (short)curx; *mv_y = (short)cury;
  return min_mcost;
}

int ora_umhex_subpel(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r,
                     int pic_pix_x, int pic_pix_y, int blocktype, int pmv_x, int pmv_y, short *mv_x, short *mv_y,
                     int min_mcost, int lambda_factor)
{
  pat_ctx k;
  pat_init(&k, c, cur, cur_stride, r, NULL, pic_pix_x, pic_pix_y, blocktype, pmv_x, pmv_y);
  return umhex_walk(&k, 0, lambda_factor, 0, 0, 0, 0, mv_x, mv_y, min_mcost);
}

static const int8_t smp_shift[8] = { 0, 0, 1, 1, 2, 3, 3, 1 };                    /* block_type_shift_factor, me_umhexsmp.c:46 */

int ora_smpumhex_subpel(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r,
                        int pic_pix_x, int pic_pix_y, int blocktype, int pmv_x, int pmv_y, short *mv_x, short *mv_y,
                        int min_mcost, int lambda_factor, int uplayer_x, int uplayer_y, int threshold1, int threshold3)
{
  pat_ctx k;
  pat_init(&k, c, cur, cur_stride, r, NULL, pic_pix_x, pic_pix_y, blocktype, pmv_x, pmv_y);
  return umhex_walk(&k, 1, lambda_factor, uplayer_x, uplayer_y, threshold1 >> smp_shift[blocktype], threshold3 >> smp_shift[blocktype],
                    mv_x, mv_y, min_mcost);
}

/* smpUMHEXFullSubPelBlockMotionSearch (me_umhexsmp.c:422-608): the nine half-pel and nine quarter-pel positions of the plain
 * refinement, both scored with the quarter-pel metric and factor, each loop left as soon as the minimum is below SubPelThreshold3 */
int ora_smpumhex_full_subpel(const ora_cfg *c, const ora_pel *cur, int cur_stride, const ora_ref *r,
                             int pic_pix_x, int pic_pix_y, int blocktype, int ref, int slice_type, int pmv_x, int pmv_y,
                             short *mv_x, short *mv_y, int search_pos2, int search_pos4, int min_mcost, int lambda_factor,
                             int threshold1, int threshold3)
{
  pat_ctx k;
  int pos, best, mcost, cx, cy;
  int check0 = (!c->rdopt && slice_type != ME_B_SLICE && ref == 0 && blocktype == 1 && *mv_x == 0 && *mv_y == 0);
  int shp = start_hp(c), sqp = start_qp(c);
  int max_pos2 = (!shp && search_pos2 < 1) ? 1 : search_pos2;
  int thr1 = threshold1 >> smp_shift[blocktype], thr3 = threshold3 >> smp_shift[blocktype];
  int16_t sx[9], sy[9];
  pat_init(&k, c, cur, cur_stride, r, NULL, pic_pix_x, pic_pix_y, blocktype, pmv_x, pmv_y);
  first_spiral(9, sx, sy);
  for (best = 0, pos = shp; pos < max_pos2; pos++) {
    cx = *mv_x + 2 * sx[pos]; cy = *mv_y + 2 * sy[pos];
    if (ora_mv_cost(lambda_factor, cx, cy, pmv_x, pmv_y) >= min_mcost) continue;  /* skips the threshold test below as well (:519) */
    mcost = pat_cost(&k, 2, lambda_factor, cx, cy);
    if (pos == 0 && check0) mcost -= (lambda_factor * 16) >> 16;
    if (mcost < min_mcost) { min_mcost = mcost; best = pos; }
    if (min_mcost < thr3) break;
  }
  if (best) { *mv_x += 2 * sx[best]; *mv_y += 2 * sy[best]; }
  if (*mv_x == 0 && *mv_y == 0 && pmv_x == 0 && pmv_y == 0 && min_mcost < thr1) return min_mcost;
  if (!sqp) min_mcost = INT_MAX;
  for (best = 0, pos = sqp; pos < search_pos4; pos++) {
    cx = *mv_x + sx[pos]; cy = *mv_y + sy[pos];
    if (ora_mv_cost(lambda_factor, cx, cy, pmv_x, pmv_y) >= min_mcost) continue;
    mcost = pat_cost(&k, 2, lambda_factor, cx, cy);
    if (mcost < min_mcost) { min_mcost = mcost; best = pos; }
    if (min_mcost < thr3) break;
  }
  if (best) { *mv_x += sx[best]; *mv_y += sy[best]; }
  return min_mcost;
}
