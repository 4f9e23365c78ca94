/* p7_fbx_oracle.c — CPU oracle, part 2: Forward/Backward parsers in the reference's OWN ORDER of float operations.
 * TEST INFRASTRUCTURE ONLY (same rules as p7_oracle.h).
 *
 * p7_oracle.c restates the two parsers in k-space (node order), which agrees with the reference to ~1e-6 relative but
 * not bit for bit, because float addition is not associative.  This file restates them in the striped order of
 * reiverjohn/biobench2 HMMER/src/impl_sse/fwdback.c:257-465 (forward_engine) and :469-735 (backward_engine): a
 * "vector" is four floats z = 0..3, vector q of a stripe holds nodes k = z*Q + q + 1, Q = p7O_NQF(M) (impl_sse.h:27),
 * every add and multiply happens in the sequence the SSE code issues it, with flush-to-zero and denormals-are-zero on
 * (impl_Init, impl_sse.h:356-376).  It reads the reference's striped arrays (om->rfv, om->tfv) through the
 * p7gpu_oprofile_view untouched.  The CUDA kernels of biobench2_b200/csrc/fbx.cu follow the same sequence with one GPU
 * lane per vector element.
 *
 * Parity status: PINNED bit-exact — tests/test_fbx_oracle.py compares every special of every row (xmx), totscale, the
 * score and has_own_scales with the unmodified reference (oracle/_ref) on all golden profiles and on fresh seeded inputs.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <xmmintrin.h>
#include <pmmintrin.h>
#include "p7_oracle.h"
#include "../include/p7gpu.h"

typedef struct { float v[4]; } V4;

static int nqf(int M) { int q = (M - 1) / 4 + 1; return q < 2 ? 2 : q; }             /* impl_sse.h:27 */

static V4 v4_set1(float x)            { V4 r; int z; for (z = 0; z < 4; z++) r.v[z] = x; return r; }
static V4 v4_load(const float *p)     { V4 r; memcpy(r.v, p, 16); return r; }
static V4 v4_add(V4 a, V4 b)          { V4 r; int z; for (z = 0; z < 4; z++) r.v[z] = a.v[z] + b.v[z]; return r; }
static V4 v4_mul(V4 a, V4 b)          { V4 r; int z; for (z = 0; z < 4; z++) r.v[z] = a.v[z] * b.v[z]; return r; }
static V4 v4_rshift(V4 a)             { V4 r; r.v[0] = 0.0f; r.v[1] = a.v[0]; r.v[2] = a.v[1]; r.v[3] = a.v[2]; return r; }   /* esl_sse_rightshift_ps(a, zerov) */
static V4 v4_lshift(V4 a)             { V4 r; r.v[0] = a.v[1]; r.v[1] = a.v[2]; r.v[2] = a.v[3]; r.v[3] = 0.0f; return r; }   /* move_ss + shuffle(0,3,2,1) */
static float v4_hsum(V4 x)            { return (x.v[0] + x.v[1]) + (x.v[2] + x.v[3]); }   /* element 0 of the two-shuffle sum, fwdback.c:393-395 */
static int v4_any_gt(V4 a, V4 b)      { int z, r = 0; for (z = 0; z < 4; z++) r |= (a.v[z] > b.v[z]); return r; }

typedef struct { unsigned int saved; } FTZ;
static FTZ  ftz_on(void)   { FTZ f; f.saved = _mm_getcsr(); _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON); _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON); return f; }
static void ftz_off(FTZ f) { _mm_setcsr(f.saved); }

#define TV(q, t)  v4_load(tfv + ((size_t) 7 * (q) + (t)) * 4)          /* BM,MM,IM,DM,MD,MI,II of vector q (p7_oprofile.c:852-879) */
#define TDD(q)    v4_load(tfv + ((size_t) 7 * Q + (q)) * 4)
#define RV(x, q)  v4_load(rfv + ((size_t) (x) * Q + (q)) * 4)

/* p7_ForwardParser (fwdback.c:133-144 -> forward_engine(FALSE, ...), :257-465) */
int p7o_forward_striped(const void *view, const uint8_t *dsq, int L, float *xmx, float *ret_totscale, float *ret_sc)
{
  const p7gpu_oprofile_view *om = (const p7gpu_oprofile_view *) view;
  const float *rfv = om->rfv, *tfv = om->tfv;
  const int Q = nqf(om->M);
  const float pmove = (2.0f + om->nj) / ((float) L + 2.0f + om->nj), ploop = 1.0f - pmove;     /* p7_oprofile.c:1099-1103 */
  const float xfE_move = om->xf[P7O_E][P7O_MOVE], xfE_loop = om->xf[P7O_E][P7O_LOOP];
  V4 *Mx = (V4 *) calloc((size_t) Q, sizeof(V4)), *Dx = (V4 *) calloc((size_t) Q, sizeof(V4)), *Ix = (V4 *) calloc((size_t) Q, sizeof(V4));
  float xN, xE, xB, xC, xJ, totscale = 0.0f;
  int i, q, j, status = P7O_OK;
  FTZ fz = ftz_on();

  xE = 0.f; xN = 1.f; xJ = 0.f; xB = pmove; xC = 0.f;
  xmx[0] = xE; xmx[1] = xN; xmx[2] = xJ; xmx[3] = xB; xmx[4] = xC; xmx[5] = 1.0f;
  for (i = 1; i <= L; i++) {
    V4 dcv = v4_set1(0.f), xEv = v4_set1(0.f), xBv = v4_set1(xB), sv;
    V4 mpv = v4_rshift(Mx[Q-1]), dpv = v4_rshift(Dx[Q-1]), ipv = v4_rshift(Ix[Q-1]);
    const int x = dsq[i];
    for (q = 0; q < Q; q++) {
      sv  =            v4_mul(xBv, TV(q, 0));
      sv  = v4_add(sv, v4_mul(mpv, TV(q, 1)));
      sv  = v4_add(sv, v4_mul(ipv, TV(q, 2)));
      sv  = v4_add(sv, v4_mul(dpv, TV(q, 3)));
      sv  = v4_mul(sv, RV(x, q));
      xEv = v4_add(xEv, sv);
      mpv = Mx[q]; dpv = Dx[q]; ipv = Ix[q];
      Mx[q] = sv;  Dx[q] = dcv;
      dcv = v4_mul(sv, TV(q, 4));
      sv  = v4_mul(mpv, TV(q, 5));
      Ix[q] = v4_add(sv, v4_mul(ipv, TV(q, 6)));
    }
    dcv = v4_rshift(dcv);
    Dx[0] = v4_set1(0.f);
    for (q = 0; q < Q; q++) { Dx[q] = v4_add(dcv, Dx[q]); dcv = v4_mul(Dx[q], TDD(q)); }
    if (om->M < 100) {
      for (j = 1; j < 4; j++) {
        dcv = v4_rshift(dcv);
        for (q = 0; q < Q; q++) { Dx[q] = v4_add(dcv, Dx[q]); dcv = v4_mul(dcv, TDD(q)); }
      }
    } else {
      for (j = 1; j < 4; j++) {
        int cv = 0;
        dcv = v4_rshift(dcv);
        for (q = 0; q < Q; q++) { sv = v4_add(dcv, Dx[q]); cv |= v4_any_gt(sv, Dx[q]); Dx[q] = sv; dcv = v4_mul(dcv, TDD(q)); }
        if (!cv) break;
      }
    }
    for (q = 0; q < Q; q++) xEv = v4_add(Dx[q], xEv);
    xE = v4_hsum(xEv);
    xN =  xN * ploop;
    xC = (xC * ploop) + (xE * xfE_move);
    xJ = (xJ * ploop) + (xE * xfE_loop);
    xB = (xJ * pmove) + (xN * pmove);
    if (xE > 1.0e4) {
      V4 sc;
      xN = xN / xE; xC = xC / xE; xJ = xJ / xE; xB = xB / xE;
      sc = v4_set1((float) (1.0 / xE));
      for (q = 0; q < Q; q++) { Mx[q] = v4_mul(Mx[q], sc); Dx[q] = v4_mul(Dx[q], sc); Ix[q] = v4_mul(Ix[q], sc); }
      xmx[i*6 + 5] = xE;
      totscale += log(xE);
      xE = 1.0f;
    } else xmx[i*6 + 5] = 1.0f;
    xmx[i*6 + 0] = xE; xmx[i*6 + 1] = xN; xmx[i*6 + 2] = xJ; xmx[i*6 + 3] = xB; xmx[i*6 + 4] = xC;
  }
  if (isnan(xC) || (L > 0 && xC == 0.0f) || isinf(xC)) status = P7O_ERANGE;                 /* fwdback.c:456-458 */
  if (ret_totscale) *ret_totscale = totscale;
  if (ret_sc) *ret_sc = totscale + log(xC * pmove);
  ftz_off(fz);
  free(Mx); free(Dx); free(Ix);
  return status;
}

/* p7_BackwardParser (fwdback.c:237-249 -> backward_engine(FALSE, ...), :469-735) */
int p7o_backward_striped(const void *view, const uint8_t *dsq, int L, const float *fwd_xmx, float *bck_xmx,
                         float *ret_totscale, int *ret_own_scales, float *ret_sc)
{
  const p7gpu_oprofile_view *om = (const p7gpu_oprofile_view *) view;
  const float *rfv = om->rfv, *tfv = om->tfv;
  const int Q = nqf(om->M);
  const float pmove = (2.0f + om->nj) / ((float) L + 2.0f + om->nj), ploop = 1.0f - pmove;
  const float xfE_move = om->xf[P7O_E][P7O_MOVE], xfE_loop = om->xf[P7O_E][P7O_LOOP];
  V4 *Mx = (V4 *) calloc((size_t) Q, sizeof(V4)), *Dx = (V4 *) calloc((size_t) Q, sizeof(V4)), *Ix = (V4 *) calloc((size_t) Q, sizeof(V4));
  V4 mpv, ipv, dpv, mcv, dcv, tmmv, timv, tdmv, xBv, xEv;
  float xN, xE, xB, xC, xJ, totscale, scale;
  int i, q, j, own = 0, status = P7O_OK;
  FTZ fz = ftz_on();

  xJ = 0.f; xB = 0.f; xN = 0.f;
  xC = pmove;
  xE = xC * xfE_move;
  xEv = v4_set1(xE);
  dcv = v4_set1(0.f);
  for (q = 0; q < Q; q++) { Mx[q] = Dx[q] = xEv; Ix[q] = v4_set1(0.f); }
  dpv = v4_lshift(Dx[Q-1]);
  for (q = Q-1; q >= 0; q--) { dcv = v4_mul(dpv, TDD(q)); Dx[q] = v4_add(Dx[q], dcv); dpv = Dx[q]; }
  for (j = 1; j < 4; j++) {
    dcv = v4_lshift(dcv);
    for (q = Q-1; q >= 0; q--) { dcv = v4_mul(dcv, TDD(q)); Dx[q] = v4_add(Dx[q], dcv); }
  }
  dcv = v4_lshift(Dx[0]);
  for (q = Q-1; q >= 0; q--) { Mx[q] = v4_add(Mx[q], v4_mul(dcv, TV(q, 4))); dcv = Dx[q]; }
  scale = fwd_xmx[L*6 + 5];
  if (scale > 1.0f) {
    V4 sc;
    xE = xE / scale; xN = xN / scale; xC = xC / scale; xJ = xJ / scale; xB = xB / scale;
    sc = v4_set1((float) (1.0 / scale));
    for (q = 0; q < Q; q++) { Mx[q] = v4_mul(Mx[q], sc); Dx[q] = v4_mul(Dx[q], sc); Ix[q] = v4_mul(Ix[q], sc); }
  }
  bck_xmx[L*6 + 5] = scale;
  totscale = log(scale);
  bck_xmx[L*6 + 0] = xE; bck_xmx[L*6 + 1] = xN; bck_xmx[L*6 + 2] = xJ; bck_xmx[L*6 + 3] = xB; bck_xmx[L*6 + 4] = xC;

  for (i = L-1; i >= 1; i--) {
    const int x = dsq[i+1];
    tmmv = v4_lshift(TV(0, 1)); timv = v4_lshift(TV(0, 2)); tdmv = v4_lshift(TV(0, 3));
    mpv  = v4_lshift(v4_mul(Mx[0], RV(x, 0)));
    xBv  = v4_set1(0.f);
    for (q = Q-1; q >= 0; q--) {
      ipv   = Ix[q];
      Ix[q] = v4_add(v4_mul(ipv, TV(q, 6)), v4_mul(mpv, timv));
      Dx[q] =                               v4_mul(mpv, tdmv);
      mcv   = v4_add(v4_mul(ipv, TV(q, 5)), v4_mul(mpv, tmmv));
      mpv   = v4_mul(Mx[q], RV(x, q));
      Mx[q] = mcv;
      tdmv = TV(q, 3); timv = TV(q, 2); tmmv = TV(q, 1);
      xBv = v4_add(xBv, v4_mul(mpv, TV(q, 0)));
    }
    xB = v4_hsum(xBv);
    xC =  xC * ploop;
    xJ = (xB * pmove) + (xJ * ploop);
    xN = (xB * pmove) + (xN * ploop);
    xE = (xC * xfE_move) + (xJ * xfE_loop);
    xEv = v4_set1(xE);
    dpv = v4_lshift(v4_add(Dx[0], xEv));
    for (q = Q-1; q >= 0; q--) {
      dcv   = v4_mul(dpv, TDD(q));
      Dx[q] = v4_add(Dx[q], v4_add(dcv, xEv));
      dpv   = Dx[q];
      Mx[q] = v4_add(Mx[q], xEv);
    }
    for (j = 1; j < 4; j++) {
      dcv = v4_lshift(dcv);
      for (q = Q-1; q >= 0; q--) { dcv = v4_mul(dcv, TDD(q)); Dx[q] = v4_add(Dx[q], dcv); }
    }
    dcv = v4_lshift(Dx[0]);
    for (q = Q-1; q >= 0; q--) { Mx[q] = v4_add(Mx[q], v4_mul(dcv, TV(q, 4))); dcv = Dx[q]; }
    if (xB > 1.0e16) own = 1;                                          /* double literal, as fwdback.c:661 */
    if (own) scale = (xB > 1.0e4) ? xB : 1.0f;
    else     scale = fwd_xmx[i*6 + 5];
    bck_xmx[i*6 + 5] = scale;
    if (scale > 1.0f) {
      V4 sc;
      xE /= scale; xN /= scale; xJ /= scale; xB /= scale; xC /= scale;
      sc = v4_set1((float) (1.0 / scale));
      for (q = 0; q < Q; q++) { Mx[q] = v4_mul(Mx[q], sc); Dx[q] = v4_mul(Dx[q], sc); Ix[q] = v4_mul(Ix[q], sc); }
      totscale += log(scale);
    }
    bck_xmx[i*6 + 0] = xE; bck_xmx[i*6 + 1] = xN; bck_xmx[i*6 + 2] = xJ; bck_xmx[i*6 + 3] = xB; bck_xmx[i*6 + 4] = xC;
  }
  {
    const int x = dsq[1];
    xBv = v4_set1(0.f);
    for (q = 0; q < Q; q++) { mpv = v4_mul(Mx[q], RV(x, q)); mpv = v4_mul(mpv, TV(q, 0)); xBv = v4_add(xBv, mpv); }
    xB = v4_hsum(xBv);
    xN = (xB * pmove) + (xN * ploop);
  }
  bck_xmx[3] = xB; bck_xmx[4] = 0.f; bck_xmx[2] = 0.f; bck_xmx[1] = xN; bck_xmx[0] = 0.f; bck_xmx[5] = 1.0f;
  if (isnan(xN) || (L > 0 && xN == 0.0f) || isinf(xN)) status = P7O_ERANGE;                 /* fwdback.c:731-733 */
  if (ret_totscale) *ret_totscale = totscale;
  if (ret_own_scales) *ret_own_scales = own;
  if (ret_sc) *ret_sc = totscale + log(xN);
  ftz_off(fz);
  free(Mx); free(Dx); free(Ix);
  return status;
}
