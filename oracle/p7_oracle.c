/* p7_oracle.c — CPU oracle for the HMMER3 p7_Pipeline() hot path.  TEST INFRASTRUCTURE ONLY
 * (see p7_oracle.h for the rules and the parity status).
 *
 * Every function cites the reference code it restates (paths under reiverjohn/biobench2 HMMER/src).
 * The restatement is in "k-space": model node k = 1..M at array index k-1, no SSE striping.
 * Build: gcc -O2 -std=c99 -fno-fast-math -ffp-contract=off (no FMA contraction: the reference's SSE2
 * build has none either).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "p7_oracle.h"
#include "../include/p7gpu.h"

#define IMAX(a,b) ((a) > (b) ? (a) : (b))
#define IMIN(a,b) ((a) < (b) ? (a) : (b))

static int nq_of(int M, int w) { int q = (M - 1) / w + 1; return q < 2 ? 2 : q; }   /* impl_sse.h:25-27 */

/* De-striping: node k sits in vector q=(k-1)%Q, lane z=(k-1)/Q (p7_oprofile.c:789-792, 845-848, 938-941);
 * transitions are 7 interleaved vectors per q (BM,MM,IM,DM,MD,MI,II) then Q DD vectors (:852-879). */
P7O_PROFILE *p7o_profile_from_view(const void *view)
{
  const p7gpu_oprofile_view *v = (const p7gpu_oprofile_view *) view;
  P7O_PROFILE *p = (P7O_PROFILE *) calloc(1, sizeof(P7O_PROFILE));
  int M = v->M, Kp = v->Kp, x, k, t;
  int Q16 = nq_of(M, 16), Q8 = nq_of(M, 8), Q4 = nq_of(M, 4);
  p->M = M; p->Kp = Kp;
  p->rb = (uint8_t *) malloc((size_t) Kp * M);
  p->sb = (int8_t  *) malloc((size_t) Kp * M);
  p->rw = (int16_t *) malloc(sizeof(int16_t) * Kp * M);
  p->tw = (int16_t *) malloc(sizeof(int16_t) * 8 * M);
  p->rf = (float   *) malloc(sizeof(float) * Kp * M);
  p->tf = (float   *) malloc(sizeof(float) * 8 * M);
  for (x = 0; x < Kp; x++)
    for (k = 1; k <= M; k++) {
      int q16 = (k-1) % Q16, z16 = (k-1) / Q16;
      int q8  = (k-1) % Q8,  z8  = (k-1) / Q8;
      int q4  = (k-1) % Q4,  z4  = (k-1) / Q4;
      p->rb[x*M + k-1] = v->rbv[((size_t) x * Q16 + q16) * 16 + z16];
      if (v->sbv) p->sb[x*M + k-1] = (int8_t) v->sbv[((size_t) x * (Q16 + 17) + q16) * 16 + z16];
      else {      /* sf_conversion, p7_oprofile.c:711-752: ((127+bias) -sat_u8 rbv) ^ 127 */
        int a = (127 + v->bias_b) - v->rbv[((size_t) x * Q16 + q16) * 16 + z16]; if (a < 0) a = 0;
        p->sb[x*M + k-1] = (int8_t) ((uint8_t) a ^ 127);
      }
      p->rw[x*M + k-1] = v->rwv[((size_t) x * Q8 + q8) * 8 + z8];
      p->rf[x*M + k-1] = v->rfv[((size_t) x * Q4 + q4) * 4 + z4];
    }
  for (k = 1; k <= M; k++) {
    int q8 = (k-1) % Q8, z8 = (k-1) / Q8, q4 = (k-1) % Q4, z4 = (k-1) / Q4;
    for (t = 0; t < 7; t++) {
      p->tw[t*M + k-1] = v->twv[((size_t) 7*q8 + t) * 8 + z8];
      p->tf[t*M + k-1] = v->tfv[((size_t) 7*q4 + t) * 4 + z4];
    }
    p->tw[P7O_DD*M + k-1] = v->twv[((size_t) 7*Q8 + q8) * 8 + z8];
    p->tf[P7O_DD*M + k-1] = v->tfv[((size_t) 7*Q4 + q4) * 4 + z4];
  }
  p->tbm_b = v->tbm_b; p->tec_b = v->tec_b; p->base_b = v->base_b; p->bias_b = v->bias_b; p->scale_b = v->scale_b;
  p->xw_E_move = v->xw[P7O_E][P7O_MOVE]; p->xw_E_loop = v->xw[P7O_E][P7O_LOOP];
  p->base_w = v->base_w; p->ddbound_w = v->ddbound_w; p->scale_w = v->scale_w;
  p->xf_E_move = v->xf[P7O_E][P7O_MOVE]; p->xf_E_loop = v->xf[P7O_E][P7O_LOOP];
  p->nj = v->nj;
  memcpy(p->evparam, v->evparam, sizeof(p->evparam));
  return p;
}

void p7o_profile_free(P7O_PROFILE *p)
{
  if (!p) return;
  free(p->rb); free(p->sb); free(p->rw); free(p->tw); free(p->rf); free(p->tf); free(p);
}

/* ---- per-target length configuration ------------------------------------------------------- */
/* unbiased_byteify(), p7_oprofile.c:676-684 */
static uint8_t byteify_cost(float scale_b, float sc)
{
  sc = -1.0f * roundf(scale_b * sc);
  return (sc > 255.) ? 255 : (uint8_t) sc;
}
/* wordify(), p7_oprofile.c:692-699 */
static int16_t wordify(float scale_w, float sc)
{
  sc = roundf(scale_w * sc);
  if      (sc >=  32767.0) return  32767;
  else if (sc <= -32768.0) return -32768;
  else return (int16_t) sc;
}
/* p7_oprofile_ReconfigMSVLength, p7_oprofile.c:1085-1090 */
uint8_t p7o_tjb_b(const P7O_PROFILE *p, int L) { return byteify_cost(p->scale_b, logf(3.0f / (float) (L+3))); }
/* p7_oprofile_ReconfigRestLength, p7_oprofile.c:1108-1127 */
void p7o_xf_ncj(const P7O_PROFILE *p, int L, float *pmove, float *ploop)
{
  *pmove = (2.0f + p->nj) / ((float) L + 2.0f + p->nj);
  *ploop = 1.0f - *pmove;
}
int16_t p7o_xw_move(const P7O_PROFILE *p, int L)
{
  float pmove, ploop; p7o_xf_ncj(p, L, &pmove, &ploop);
  return wordify(p->scale_w, logf(pmove));
}
/* p7_bg_SetLength + p7_bg_NullOne, p7_bg.c:190-199, 358-363 (double log of a float p1) */
float p7o_nullsc(int L)
{
  float p1 = (float) L / (float) (L+1);
  return (float) L * log(p1) + log(1. - p1);
}

/* ---- SSV: ssvfilter.c:833-926 -------------------------------------------------------------- */
int p7o_ssvfilter(const P7O_PROFILE *p, const uint8_t *dsq, int L, float *ret_sc)
{
  int M = p->M, i, k;
  uint8_t tjb_b = p7o_tjb_b(p, L);
  uint16_t xE, xJ;
  int *u, maxu = 0;

  if (tjb_b + p->tbm_b + p->tec_b + p->bias_b >= 127) return P7O_ENORESULT;      /* :884 */

  /* get_xE(): running diagonals; u = sv + 128, floor by signed saturation at -128 (:670-711).
   * The upper end is left unclamped: once any cell reaches 255-bias the outcome below is fixed. */
  u = (int *) calloc(M + 1, sizeof(int));
  for (i = 1; i <= L; i++) {
    const int8_t *sb = p->sb + (size_t) dsq[i] * M;
    for (k = M; k >= 1; k--) {
      int v = u[k-1] - sb[k-1];
      if (v < 0) v = 0;
      if (v > 255) v = 255;
      u[k] = v;
      if (v > maxu) maxu = v;
    }
    u[0] = 0;
  }
  free(u);
  xE = (uint16_t) IMIN(255, 128 + maxu);     /* unsigned max of bytes that start at 128 (:851-852) */

  if (xE >= 255 - p->bias_b) {               /* :892-904 */
    *ret_sc = INFINITY;
    if (p->base_b - tjb_b - p->tbm_b < 128) return P7O_ENORESULT;
    return P7O_ERANGE;
  }
  xE += p->base_b - tjb_b - p->tbm_b;        /* :906-907, uint16 arithmetic */
  xE -= 128;
  if (xE >= 255 - p->bias_b) { *ret_sc = INFINITY; return P7O_ERANGE; }   /* :909-914 */
  xJ = xE - p->tec_b;
  if (xJ > p->base_b) return P7O_ENORESULT;  /* :918 */
  *ret_sc = ((float) (xJ - tjb_b) - (float) p->base_b);
  *ret_sc /= p->scale_b;
  *ret_sc -= 3.0;
  return P7O_OK;
}

/* ---- MSV: msvfilter.c:75-210 --------------------------------------------------------------- */
static uint8_t adds_u8(int a, int b) { int s = a + b; return (uint8_t) (s > 255 ? 255 : s); }
static uint8_t subs_u8(int a, int b) { int s = a - b; return (uint8_t) (s < 0 ? 0 : s); }

int p7o_msvfilter(const P7O_PROFILE *p, const uint8_t *dsq, int L, float *ret_sc)
{
  int M = p->M, i, k, status;
  uint8_t tjb_b = p7o_tjb_b(p, L);
  uint8_t tjbm  = (uint8_t) ((int8_t) tjb_b + (int8_t) p->tbm_b);     /* :118 */
  uint8_t xJ, xB, xE, *mx;

  status = p7o_ssvfilter(p, dsq, L, ret_sc);                          /* :105-106 */
  if (status != P7O_ENORESULT) return status;

  mx = (uint8_t *) calloc(M + 1, 1);
  xJ = 0;
  xB = subs_u8(p->base_b, tjbm);
  for (i = 1; i <= L; i++) {
    const uint8_t *rb = p->rb + (size_t) dsq[i] * M;
    xE = 0;
    for (k = M; k >= 1; k--) {                                        /* :141-154 */
      uint8_t sv = IMAX(mx[k-1], xB);
      sv = adds_u8(sv, p->bias_b);
      sv = subs_u8(sv, rb[k-1]);
      if (sv > xE) xE = sv;
      mx[k] = sv;
    }
    mx[0] = 0;
    if (adds_u8(xE, p->bias_b) == 255) { free(mx); *ret_sc = INFINITY; return P7O_ERANGE; }   /* :157-182 */
    xE = subs_u8(xE, p->tec_b);                                       /* :184-188 */
    xJ = IMAX(xJ, xE);
    xB = IMAX(p->base_b, xJ);
    xB = subs_u8(xB, tjbm);
  }
  free(mx);
  *ret_sc = ((float) (xJ - tjb_b) - (float) p->base_b);               /* :205-207 */
  *ret_sc /= p->scale_b;
  *ret_sc -= 3.0;
  return P7O_OK;
}

/* ---- ViterbiFilter: vitfilter.c:83-249 ----------------------------------------------------- */
long long p7o_dbg_vit_rows = 0, p7o_dbg_vit_ddrows = 0;   /* diagnostics: how often lazy-F evaluates D->D */
static int16_t adds16(int a, int b) { int s = a + b; return (int16_t) (s > 32767 ? 32767 : (s < -32768 ? -32768 : s)); }

static int vit_engine(const P7O_PROFILE *p, const uint8_t *dsq, int L, int always_dd, float *ret_sc)
{
  int M = p->M, i, k;
  const int16_t *tBM = p->tw + P7O_BM*M, *tMM = p->tw + P7O_MM*M, *tIM = p->tw + P7O_IM*M, *tDM = p->tw + P7O_DM*M;
  const int16_t *tMD = p->tw + P7O_MD*M, *tMI = p->tw + P7O_MI*M, *tII = p->tw + P7O_II*M, *tDD = p->tw + P7O_DD*M;
  int16_t xw_move = p7o_xw_move(p, L);          /* xw[N|C|J][MOVE]; the LOOPs are 0 (p7_oprofile.c:891-895) */
  int16_t xE, xB, xC, xJ, xN, Dmax;
  int16_t *mmx = (int16_t *) malloc(sizeof(int16_t) * 3 * (M + 2)), *imx = mmx + (M + 2), *dmx = imx + (M + 2);
  int16_t *nm  = (int16_t *) malloc(sizeof(int16_t) * 3 * (M + 2)), *ni  = nm + (M + 2),  *nd  = ni + (M + 2);

  for (k = 0; k <= M + 1; k++) mmx[k] = imx[k] = dmx[k] = -32768;     /* :113-114 */
  xN = p->base_w;
  xB = xN + xw_move;
  xJ = -32768; xC = -32768; xE = -32768;

  for (i = 1; i <= L; i++) {
    const int16_t *rw = p->rw + (size_t) dsq[i] * M;
    int16_t dcv = -32768;
    xE = -32768; Dmax = -32768;
    nm[0] = ni[0] = nd[0] = -32768;
    for (k = 1; k <= M; k++) {                                        /* :143-173 */
      int16_t sv =   adds16(xB,       tBM[k-1]);
      sv = IMAX(sv, adds16(mmx[k-1], tMM[k-1]));
      sv = IMAX(sv, adds16(imx[k-1], tIM[k-1]));
      sv = IMAX(sv, adds16(dmx[k-1], tDM[k-1]));
      sv = adds16(sv, rw[k-1]);
      xE = IMAX(xE, sv);
      nm[k] = sv;
      nd[k] = dcv;                        /* D_k = M_{k-1} + tMD_{k-1}, so far */
      dcv   = adds16(sv, tMD[k-1]);
      Dmax  = IMAX(Dmax, dcv);
      ni[k] = IMAX(adds16(mmx[k], tMI[k-1]), adds16(imx[k], tII[k-1]));
    }
    if (xE >= 32767) { free(mmx); free(nm); *ret_sc = INFINITY; return P7O_ERANGE; }       /* :177 */
    /* specials: int arithmetic, truncated on assignment to int16_t (:178-181) */
    xN = (int16_t) (xN + 0);
    xC = (int16_t) IMAX(xC + 0, xE + p->xw_E_move);
    xJ = (int16_t) IMAX(xJ + 0, xE + p->xw_E_loop);
    xB = (int16_t) IMAX(xJ + xw_move, xN + xw_move);
    p7o_dbg_vit_rows++;
    if (always_dd || Dmax + p->ddbound_w > xB) {                      /* lazy F, :198-232 */
      p7o_dbg_vit_ddrows++;
      for (k = 2; k <= M; k++) nd[k] = IMAX(nd[k], adds16(nd[k-1], tDD[k-2]));
    }
    { int16_t *t; t = mmx; mmx = nm; nm = t; t = imx; imx = ni; ni = t; t = dmx; dmx = nd; nd = t; }
  }
  if (mmx < nm) { free(mmx); free(nm); } else { free(nm); free(mmx); }
  if (xC > -32768) {                                                  /* :240-247 */
    *ret_sc  = (float) xC + (float) xw_move - (float) p->base_w;
    *ret_sc /= p->scale_w;
    *ret_sc -= 3.0;
  } else *ret_sc = -INFINITY;
  return P7O_OK;
}
int p7o_vitfilter(const P7O_PROFILE *p, const uint8_t *dsq, int L, float *ret_sc) { return vit_engine(p, dsq, L, 0, ret_sc); }
int p7o_vitfilter_alwaysdd(const P7O_PROFILE *p, const uint8_t *dsq, int L, float *ret_sc) { return vit_engine(p, dsq, L, 1, ret_sc); }

/* ---- ForwardParser: fwdback.c:257-465 (do_full = 0) ---------------------------------------- */
int p7o_forward(const P7O_PROFILE *p, const uint8_t *dsq, int L, float *xmx, float *ret_totscale, float *ret_sc)
{
  int M = p->M, i, k;
  const float *tBM = p->tf + P7O_BM*M, *tMM = p->tf + P7O_MM*M, *tIM = p->tf + P7O_IM*M, *tDM = p->tf + P7O_DM*M;
  const float *tMD = p->tf + P7O_MD*M, *tMI = p->tf + P7O_MI*M, *tII = p->tf + P7O_II*M, *tDD = p->tf + P7O_DD*M;
  float pmove, ploop, xN, xE, xB, xC, xJ, totscale = 0.0f;
  float *mmx = (float *) calloc(6 * (M + 2), sizeof(float)), *imx = mmx + (M+2), *dmx = imx + (M+2);
  float *nm = dmx + (M+2), *ni = nm + (M+2), *nd = ni + (M+2);
  p7o_xf_ncj(p, L, &pmove, &ploop);
  xE = 0.f; xN = 1.f; xJ = 0.f; xB = pmove; xC = 0.f;                 /* :287-294 */
  if (xmx) { xmx[0] = xE; xmx[1] = xN; xmx[2] = xJ; xmx[3] = xB; xmx[4] = xC; xmx[5] = 1.0f; }

  for (i = 1; i <= L; i++) {
    const float *rf = p->rf + (size_t) dsq[i] * M;
    float dcv = 0.f, sumM = 0.f, sumD = 0.f, scale;
    nm[0] = ni[0] = nd[0] = 0.f;
    for (k = 1; k <= M; k++) {                                        /* :318-345 */
      float sv = xB * tBM[k-1];
      sv = sv + mmx[k-1] * tMM[k-1];
      sv = sv + imx[k-1] * tIM[k-1];
      sv = sv + dmx[k-1] * tDM[k-1];
      sv = sv * rf[k-1];
      sumM += sv;
      nm[k] = sv;
      nd[k] = dcv;                        /* M_{k-1} tMD_{k-1} */
      dcv   = sv * tMD[k-1];
      ni[k] = mmx[k] * tMI[k-1] + imx[k] * tII[k-1];
    }
    for (k = 2; k <= M; k++) nd[k] = nd[k] + nd[k-1] * tDD[k-2];      /* DD paths, :351-399, serialised */
    for (k = 1; k <= M; k++) sumD += nd[k];                           /* :402 */
    xE = sumM + sumD;
    xN = xN * ploop;                                                  /* :413-416 */
    xC = (xC * ploop) + (xE * p->xf_E_move);
    xJ = (xJ * ploop) + (xE * p->xf_E_loop);
    xB = (xJ * pmove) + (xN * pmove);
    if (xE > 1.0e4) {                                                 /* :420-437 */
      float inv = 1.0 / xE;
      xN = xN / xE; xC = xC / xE; xJ = xJ / xE; xB = xB / xE;
      for (k = 1; k <= M; k++) { nm[k] *= inv; nd[k] *= inv; ni[k] *= inv; }
      scale = xE;
      totscale += log(xE);
      xE = 1.0f;
    } else scale = 1.0f;
    if (xmx) { float *r = xmx + (size_t) i * 6; r[0] = xE; r[1] = xN; r[2] = xJ; r[3] = xB; r[4] = xC; r[5] = scale; }
    { float *t; t = mmx; mmx = nm; nm = t; t = imx; imx = ni; ni = t; t = dmx; dmx = nd; nd = t; }
  }
  free(mmx < nm ? mmx : nm);
  if (ret_totscale) *ret_totscale = totscale;
  if (isnan(xC) || (L > 0 && xC == 0.0) || isinf(xC)) { *ret_sc = xC; return P7O_ERANGE; }   /* :459-461 */
  *ret_sc = totscale + log(xC * pmove);                               /* :463 */
  return P7O_OK;
}

/* ---- BackwardParser: fwdback.c:469-735 (do_full = 0) --------------------------------------- */
int p7o_backward(const P7O_PROFILE *p, const uint8_t *dsq, int L, const float *fwd_xmx, float *bck_xmx,
                 float *ret_totscale, int *ret_own_scales, float *ret_sc)
{
  int M = p->M, i, k, own_scales = 0;
  const float *tBM = p->tf + P7O_BM*M, *tMM = p->tf + P7O_MM*M, *tIM = p->tf + P7O_IM*M, *tDM = p->tf + P7O_DM*M;
  const float *tMD = p->tf + P7O_MD*M, *tMI = p->tf + P7O_MI*M, *tII = p->tf + P7O_II*M, *tDD = p->tf + P7O_DD*M;
  float pmove, ploop, xN, xE, xB, xC, xJ, totscale, scale;
  float *mmx = (float *) calloc(6 * (M + 3), sizeof(float)), *imx = mmx + (M+3), *dmx = imx + (M+3);
  float *nm = dmx + (M+3), *ni = nm + (M+3), *nd = ni + (M+3);
  p7o_xf_ncj(p, L, &pmove, &ploop);

  /* row L, :492-552 */
  xJ = 0.f; xB = 0.f; xN = 0.f;
  xC = pmove;
  xE = xC * p->xf_E_move;
  mmx[M+1] = imx[M+1] = dmx[M+1] = 0.f;
  for (k = M; k >= 1; k--) {
    dmx[k] = xE + dmx[k+1] * tDD[k-1];
    mmx[k] = xE + dmx[k+1] * tMD[k-1];
    imx[k] = 0.f;
  }
  scale = fwd_xmx[(size_t) L * 6 + 5];
  if (scale > 1.0f) {
    float inv = 1.0 / scale;
    xE = xE / scale; xN = xN / scale; xC = xC / scale; xJ = xJ / scale; xB = xB / scale;
    for (k = 1; k <= M; k++) { mmx[k] *= inv; dmx[k] *= inv; imx[k] *= inv; }
  }
  totscale = log(scale);
  if (bck_xmx) { float *r = bck_xmx + (size_t) L * 6; r[0] = xE; r[1] = xN; r[2] = xJ; r[3] = xB; r[4] = xC; r[5] = scale; }

  for (i = L-1; i >= 1; i--) {                                        /* :566-692 */
    const float *rf = p->rf + (size_t) dsq[i+1] * M;
    float sumB = 0.f;
    nm[M+1] = ni[M+1] = nd[M+1] = 0.f;
    for (k = M; k >= 1; k--) {
      float mpv = (k < M) ? mmx[k+1] * rf[k] : 0.f;                   /* M(i+1,k+1) e(k+1) */
      float tmm = (k < M) ? tMM[k] : 0.f, tim = (k < M) ? tIM[k] : 0.f, tdm = (k < M) ? tDM[k] : 0.f;
      ni[k] = imx[k] * tII[k-1] + mpv * tim;
      nd[k] = mpv * tdm;
      nm[k] = imx[k] * tMI[k-1] + mpv * tmm;
      sumB += (mmx[k] * rf[k-1]) * tBM[k-1];
    }
    xB = sumB;
    xC = xC * ploop;
    xJ = (xB * pmove) + (xJ * ploop);
    xN = (xB * pmove) + (xN * ploop);
    xE = (xC * p->xf_E_move) + (xJ * p->xf_E_loop);
    for (k = M; k >= 1; k--) {                                        /* phases 3-5 */
      float dnext = nd[k+1];
      nd[k] = nd[k] + (dnext * tDD[k-1] + xE);
      nm[k] = (nm[k] + xE) + dnext * tMD[k-1];
    }
    if (xB > 1.0e16) own_scales = 1;                                  /* :661 */
    if (own_scales) scale = (xB > 1.0e4) ? xB : 1.0f;
    else            scale = fwd_xmx[(size_t) i * 6 + 5];
    if (scale > 1.0f) {
      float inv = 1.0 / scale;
      xE /= scale; xN /= scale; xJ /= scale; xB /= scale; xC /= scale;
      for (k = 1; k <= M; k++) { nm[k] *= inv; nd[k] *= inv; ni[k] *= inv; }
      totscale += log(scale);
    }
    if (bck_xmx) { float *r = bck_xmx + (size_t) i * 6; r[0] = xE; r[1] = xN; r[2] = xJ; r[3] = xB; r[4] = xC; r[5] = scale; }
    { float *t; t = mmx; mmx = nm; nm = t; t = imx; imx = ni; ni = t; t = dmx; dmx = nd; nd = t; }
  }
  /* termination, :697-733 */
  {
    const float *rf = p->rf + (size_t) dsq[1] * M;
    float sumB = 0.f;
    if (L >= 1) for (k = 1; k <= M; k++) sumB += (mmx[k] * rf[k-1]) * tBM[k-1];
    xB = sumB;
    xN = (xB * pmove) + (xN * ploop);
    if (bck_xmx) { bck_xmx[0] = 0.f; bck_xmx[1] = xN; bck_xmx[2] = 0.f; bck_xmx[3] = xB; bck_xmx[4] = 0.f; bck_xmx[5] = 1.0f; }
  }
  free(mmx < nm ? mmx : nm);
  if (ret_totscale)   *ret_totscale = totscale;
  if (ret_own_scales) *ret_own_scales = own_scales;
  if (isnan(xN) || (L > 0 && xN == 0.0) || isinf(xN)) { if (ret_sc) *ret_sc = xN; return P7O_ERANGE; }
  if (ret_sc) *ret_sc = totscale + log(xN);
  return P7O_OK;
}


/* p7_DomainDecoding (impl_sse/decoding.c:166-200) restated on plain arrays; the N/J/C loop probabilities are the
 * length-L values p7_oprofile_ReconfigRestLength sets (p7_oprofile.c:1112-1116), all three equal to ploop.
 * The trigger is the rt1 test of p7_domaindef_ByPosteriorHeuristics (p7_domaindef.c:109,419). */
int p7o_domain_decoding(const P7O_PROFILE *p, int L, const float *fx, const float *bx, int own,
                        float *btot, float *etot, float *mocc, float *mocc_max, int *triggered)
{
  float pmove, ploop, scaleproduct = 1.0 / bx[1], njcp, b = 0.0f, e = 0.0f, m, mx = -INFINITY;
  int i, trig = 0;
  p7o_xf_ncj(p, L, &pmove, &ploop);
  if (btot) btot[0] = 0.0f;
  if (etot) etot[0] = 0.0f;
  if (mocc) mocc[0] = 0.0f;
  for (i = 1; i <= L; i++) {
    b = b + (fx[(i-1)*6+3] * bx[(i-1)*6+3] * fx[(i-1)*6+5] * scaleproduct);
    if (own) scaleproduct *= fx[(i-1)*6+5] / bx[(i-1)*6+5];
    e = e + (fx[i*6+0] * bx[i*6+0] * fx[i*6+5] * scaleproduct);
    njcp  = fx[(i-1)*6+1] * bx[i*6+1] * ploop * scaleproduct;
    njcp += fx[(i-1)*6+2] * bx[i*6+2] * ploop * scaleproduct;
    njcp += fx[(i-1)*6+4] * bx[i*6+4] * ploop * scaleproduct;
    m = 1. - njcp;
    if (btot) btot[i] = b;
    if (etot) etot[i] = e;
    if (mocc) mocc[i] = m;
    if (m > mx) mx = m;
    if (m >= 0.25f) trig = 1;
  }
  if (mocc_max) *mocc_max = mx;
  if (triggered) *triggered = trig;
  return isinf(scaleproduct) ? 16 : 0;
}
/* ---- bias filter: esl_hmm_Forward with M=2 (easel/esl_hmm.c:351-409) + p7_bg_FilterScore (p7_bg.c:472-484) ---- */
float p7o_filtersc(const float *t6, const float *pi3, const float *eo, const uint8_t *dsq, int L)
{
  float p1 = (float) L / (float) (L+1);
  float t[2][3], dp0, dp1, max, logsc = 0.f, sc, last;
  int i;
  memcpy(t, t6, sizeof(t));
  t[0][0] = p1; t[0][1] = 1.0f - p1;                 /* p7_bg_SetLength, p7_bg.c:193-196 */
  if (L == 0) return (float) log(pi3[2]) + (float) L * logf(p1) + logf(1. - p1);
  dp0 = eo[dsq[1]*2+0] * pi3[0];
  dp1 = eo[dsq[1]*2+1] * pi3[1];
  max = 0.0f; if (dp0 > max) max = dp0; if (dp1 > max) max = dp1;
  dp0 /= max; dp1 /= max;
  sc = log(max); logsc += sc;
  for (i = 2; i <= L; i++) {
    float n0 = 0.0f, n1 = 0.0f;
    n0 += dp0 * t[0][0]; n0 += dp1 * t[1][0]; n0 *= eo[dsq[i]*2+0];
    n1 += dp0 * t[0][1]; n1 += dp1 * t[1][1]; n1 *= eo[dsq[i]*2+1];
    max = 0.0f; if (n0 > max) max = n0; if (n1 > max) max = n1;
    dp0 = n0 / max; dp1 = n1 / max;
    sc = log(max); logsc += sc;
  }
  last = 0.0f; last += dp0 * t[0][2]; last += dp1 * t[1][2];
  last = log(last); logsc += last;
  return logsc + (float) L * logf(p1) + logf(1. - p1);
}

double p7o_gumbel_surv(double x, double mu, double lambda)
{
  double y = lambda * (x - mu), ey = -exp(-y);
  if (fabs(ey) < 5e-9) return -ey;
  return 1 - exp(ey);
}
double p7o_exp_surv(double x, double mu, double lambda)
{
  if (x < mu) return 1.0;
  return exp(-lambda * (x - mu));
}

int p7o_stage_batch(const P7O_PROFILE *p, int which, const uint8_t *residues, const int64_t *offsets,
                    const int32_t *lengths, const int64_t *idx, int64_t n, float *sc, int32_t *status)
{
  int64_t j;
  for (j = 0; j < n; j++) {
    int64_t s = idx ? idx[j] : j;
    const uint8_t *dsq = residues + offsets[s] - 1;
    int L = lengths[s];
    switch (which) {
    case 0: sc[j] = 0.f; status[j] = p7o_ssvfilter(p, dsq, L, &sc[j]); break;
    case 1: status[j] = p7o_msvfilter(p, dsq, L, &sc[j]); break;
    case 2: status[j] = p7o_vitfilter(p, dsq, L, &sc[j]); break;
    case 3: status[j] = p7o_forward(p, dsq, L, NULL, NULL, &sc[j]); break;
    default: return -1;
    }
  }
  return 0;
}

/* ---- the filter chain of p7_Pipeline, p7_pipeline.c:682-731 -------------------------------- */
#define LN2 0.69314718055994529
int p7o_pipeline(const P7O_PROFILE *p, const float *t6, const float *pi3, const float *eo,
                 const uint8_t *residues, const int64_t *offsets, const int32_t *lengths, int64_t n,
                 double F1, double F2, double F3, int do_biasfilter,
                 uint64_t *counters, uint8_t *stage_reached, float *usc_o, float *vfsc_o, float *fwdsc_o, float *filtersc_o)
{
  int64_t s;
  memset(counters, 0, 6 * sizeof(uint64_t));
  for (s = 0; s < n; s++) {
    const uint8_t *dsq = residues + offsets[s] - 1;
    int L = lengths[s];
    float usc, vfsc = NAN, fwdsc = NAN, filtersc = NAN, nullsc, seq_score;
    double P;
    stage_reached[s] = 0;
    if (usc_o) usc_o[s] = NAN;
    if (vfsc_o) vfsc_o[s] = NAN;
    if (fwdsc_o) fwdsc_o[s] = NAN;
    if (filtersc_o) filtersc_o[s] = NAN;
    counters[0]++; counters[1] += L;                                   /* p7_pli_NewSeq, :572-579 */
    if (L == 0) continue;                                              /* :682 */
    nullsc = p7o_nullsc(L);
    p7o_msvfilter(p, dsq, L, &usc);                                    /* status ignored, :690 */
    if (usc_o) usc_o[s] = usc;
    seq_score = (usc - nullsc) / LN2;
    P = p7o_gumbel_surv(seq_score, p->evparam[0], p->evparam[1]);
    if (P > F1) continue;
    counters[2]++; stage_reached[s] = 1;
    if (do_biasfilter) {
      filtersc = p7o_filtersc(t6, pi3, eo, dsq, L);
      seq_score = (usc - filtersc) / LN2;
      P = p7o_gumbel_surv(seq_score, p->evparam[0], p->evparam[1]);
      if (filtersc_o) filtersc_o[s] = filtersc;
      if (P > F1) continue;
    } else filtersc = nullsc;
    if (filtersc_o) filtersc_o[s] = filtersc;
    counters[3]++; stage_reached[s] = 2;
    if (P > F2) {
      p7o_vitfilter(p, dsq, L, &vfsc);
      if (vfsc_o) vfsc_o[s] = vfsc;
      seq_score = (vfsc - filtersc) / LN2;
      P = p7o_gumbel_surv(seq_score, p->evparam[2], p->evparam[3]);
      if (P > F2) continue;
    }
    counters[4]++; stage_reached[s] = 3;
    p7o_forward(p, dsq, L, NULL, NULL, &fwdsc);
    if (fwdsc_o) fwdsc_o[s] = fwdsc;
    seq_score = (fwdsc - filtersc) / LN2;
    P = p7o_exp_surv(seq_score, p->evparam[4], p->evparam[5]);
    if (P > F3) continue;
    counters[5]++; stage_reached[s] = 4;
  }
  return 0;
}
