// fbx.cu — p7_ForwardParser / p7_BackwardParser in the reference's OWN ORDER of float operations ("exact-order" kernels).
//
// The fast parsers (fwd_mw.cu, bwd_mw.cu) associate the sums of a DP row differently from the reference's striped SSE code
// and agree with it to ~4e-6 relative.  That is inside north_star's 1e-4, but what the host builds from the specials — the
// posterior region boundaries of p7_domaindef_ByPosteriorHeuristics, the bit score printed with one decimal, the F3 test
// next to its threshold — is decided by comparisons, so a last-bit difference can change a reported line.  These kernels
// remove the difference instead of bounding it: they execute forward_engine / backward_engine
// (HMMER/src/impl_sse/fwdback.c:257-465, :469-735, do_full = 0) operation by operation.
//
//   * An SSE vector is four floats; vector q of a stripe holds nodes k = z*Q + q + 1, z = 0..3, Q = p7O_NQF(M)
//     (impl_sse.h:27).  Here FOUR GPU LANES play the four vector elements of one target, so a warp works on eight targets;
//     _mm_add_ps/_mm_mul_ps become __fadd_rn/__fmul_rn per lane (never contracted to FMA; the library is built with
//     -ftz=true: the reference runs FTZ + DAZ, impl_sse.h:356-376), the element shifts (esl_sse_rightshift_ps, the
//     move_ss + shuffle left shift) are width-4 warp shuffles, the horizontal sums are two xor-shuffles in the reference's
//     (x0+x1)+(x2+x3) association, the "did any DD path change a D cell" test of :375-389 is a ballot over the four lanes.
//   * The one DP row the parsers keep (M, D, I per vector) lives in shared memory, [q][3][32 lanes] (conflict-free), or for
//     long models in an L2-resident global scratch of the same shape.
//   * The profile is read in the reference's own striped values (om->rfv, om->tfv through the view), re-laid out only for
//     load width: tfx[q][z][8] = {BM,MM,IM,DM,MD,MI,II,DD} (two 16-byte loads per vector), rfx[x][z][Qp] (one 16-byte load
//     per four vectors).
//   * Eight targets of different lengths share a warp: the warp loops over ROWS, every group with a target advances its own
//     target by one row per iteration and draws a new target when it finishes one.
//
// The double-precision log() of the scale factors (totscale, fwdback.c:434,:689) is the only libm call; the device value
// is provisional — whenever specials are returned to the host, the host recomputes totscale and the score from them with
// its own libm in the reference's expressions (p7gpu_host.cu: fbx_host_scores), so nothing reported depends on device libm.
#include "dp_common.cuh"
#include <math_constants.h>

namespace p7g {

struct FbxParams {
  DpCommon c;
  const float *tfx;       // [Q][4][8]
  const float *rfx;       // [NTABROWS][4][Qp]
  int Q, Qp, M;
  FwdScalars s;
  float *xmx;             // specials of this direction, [(L+1)*6] per target at xmx_off[out] (backward: may be null)
  const float *fwd_xmx;   // backward only
  const int64_t *xmx_off;
  float *totscale; int32_t *own;
  float *gstate;          // global-state variant: [gridDim.x][Q][3][32]
};

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float x_rshift(float v, int z) { const float r = __shfl_up_sync(FULL, v, 1, 4);   return z == 0 ? 0.0f : r; }
__device__ __forceinline__ float x_lshift(float v, int z) { const float r = __shfl_down_sync(FULL, v, 1, 4); return z == 3 ? 0.0f : r; }
__device__ __forceinline__ float x_hsum(float x)
{
  const float a = __fadd_rn(x, __shfl_xor_sync(FULL, x, 1));          // lanes 0,1: x0+x1   lanes 2,3: x2+x3
  return __fadd_rn(a, __shfl_xor_sync(FULL, a, 2));                   // (x0+x1)+(x2+x3) on every lane of the group
}
__device__ __forceinline__ void x_store_row(float *row, int z, float xE, float xN, float xJ, float xB, float xC, float scale)
{
  float2 *r2 = reinterpret_cast<float2 *>(row);
  if (z == 0) r2[0] = make_float2(xE, xN); else if (z == 1) r2[1] = make_float2(xJ, xB); else if (z == 2) r2[2] = make_float2(xC, scale);
}

#define ST(q, s)  st[((q) * 3 + (s)) * 32 + lane]
#define TDD(q)    __ldg(p.tfx + ((size_t) (q) * 4 + z) * 8 + 7)
#define TMD(q)    __ldg(p.tfx + ((size_t) (q) * 4 + z) * 8 + 4)

// draw the next target for the groups that need one; returns the work index (valid on every lane of the group)
__device__ __forceinline__ bool fbx_draw(const DpCommon &c, bool need, int z, int gl, long long *s, long long *out)
{
  unsigned long long w = 0;
  if (need && z == 0) w = atomicAdd(c.work_counter, 1ULL);
  w = __shfl_sync(FULL, w, gl);
  if (!need || w >= (unsigned long long) c.n) return false;
  *s   = c.idx ? c.idx[w] : (c.order ? (long long) c.order[w] : c.first + (long long) w);
  *out = c.idx ? (long long) w : *s;
  return true;
}

template <bool GS>
__global__ void __launch_bounds__(32) fbx_fwd_kernel(const FbxParams p)
{
  extern __shared__ float s_state[];
  float *st = GS ? p.gstate + (size_t) blockIdx.x * p.Q * 96 : s_state;
  const int lane = threadIdx.x, z = lane & 3, gl = lane & ~3;
  const int Q = p.Q;
  const float4 *tf4 = reinterpret_cast<const float4 *>(p.tfx) + z * 2;
  const bool early = p.M >= 100;                                       // fwdback.c:352: full serialisation below M = 100

  bool have = false, exhausted = false;
  long long out = 0;
  const uint8_t *seq = nullptr; float *xrow = nullptr;
  int L = 0, i = 0, xnext = 0;
  float xE = 0.f, xN = 0.f, xJ = 0.f, xB = 0.f, xC = 0.f, pmove = 0.f, ploop = 0.f, totscale = 0.f;

  for (;;) {
    {
      const bool need = !have && !exhausted;
      long long s;
      const bool got = fbx_draw(p.c, need, z, gl, &s, &out);
      if (need && !got) exhausted = true;
      if (got) {
        L = p.c.len[s]; seq = p.c.res + p.c.off[s];
        pmove = p.c.lc[L].pmove; ploop = __fsub_rn(1.0f, pmove);
        xrow = p.xmx + p.xmx_off[out];
        for (int q = 0; q < Q; q++) { ST(q, 0) = 0.f; ST(q, 1) = 0.f; ST(q, 2) = 0.f; }
        xE = 0.f; xN = 1.f; xJ = 0.f; xB = pmove; xC = 0.f; totscale = 0.f;      // fwdback.c:287-294
        x_store_row(xrow, z, xE, xN, xJ, xB, xC, 1.0f);
        i = 1; have = true;
        xnext = L > 0 ? (int) seq[0] : 0;
      }
    }
    if (!__any_sync(FULL, have)) { if (__all_sync(FULL, exhausted)) break; continue; }

    if (__any_sync(FULL, have && i <= L)) {
      // ---------------- one DP row (fwdback.c:303-452) for every group that has a target ----------------
      const bool act = have && i <= L;
      const int x = act ? xnext : 0;
      xnext = (act && i < L) ? (int) seq[i] : 0;
      const float *rf = p.rfx + ((size_t) x * 4 + z) * p.Qp;
      float mpv = x_rshift(ST(Q - 1, 0), z), dpv = x_rshift(ST(Q - 1, 1), z), ipv = x_rshift(ST(Q - 1, 2), z);
      float dcv = 0.f, xEv = 0.f;
      const float xBv = xB;
      for (int q0 = 0; q0 < Q; q0 += 4) {
        const float4 e4 = __ldg(reinterpret_cast<const float4 *>(rf + q0));
        const float ee[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int q = q0 + u;
          if (q < Q) {
            const float4 ta = __ldg(tf4 + q * 8), tb = __ldg(tf4 + q * 8 + 1);
            float sv = __fmul_rn(xBv, ta.x);                            // :320-326
            sv = __fadd_rn(sv, __fmul_rn(mpv, ta.y));
            sv = __fadd_rn(sv, __fmul_rn(ipv, ta.z));
            sv = __fadd_rn(sv, __fmul_rn(dpv, ta.w));
            sv = __fmul_rn(sv, ee[u]);
            xEv = __fadd_rn(xEv, sv);
            mpv = ST(q, 0); dpv = ST(q, 1); ipv = ST(q, 2);             // :331-333
            ST(q, 0) = sv; ST(q, 1) = dcv;                              // :336-337
            dcv = __fmul_rn(sv, tb.x);                                  // :342
            ST(q, 2) = __fadd_rn(__fmul_rn(mpv, tb.y), __fmul_rn(ipv, tb.z));   // :345-346
          }
        }
      }
      // DD paths: one full pass including the M->D parts (:358-366)
      dcv = x_rshift(dcv, z);
      for (int q = 0; q < Q; q++) {
        float d = (q == 0) ? 0.0f : ST(q, 1);
        d = __fadd_rn(dcv, d);
        ST(q, 1) = d;
        dcv = __fmul_rn(d, TDD(q));
      }
      // up to three more passes that extend only the DD component (:352-391)
      bool more = act;
      for (int j = 1; j < 4; j++) {
        if (!__any_sync(FULL, more)) break;
        dcv = x_rshift(dcv, z);
        bool changed = false;
        for (int q = 0; q < Q; q++) {
          const float d = ST(q, 1), sv = __fadd_rn(dcv, d);
          changed |= (sv > d);
          if (more) ST(q, 1) = sv;
          dcv = __fmul_rn(dcv, TDD(q));
        }
        if (early) {
          const unsigned b = __ballot_sync(FULL, changed);
          if (((b >> gl) & 0xfu) == 0) more = false;                   // :389 "DD's didn't change any DMO(q)? Then done"
        }
      }
      for (int q = 0; q < Q; q++) xEv = __fadd_rn(ST(q, 1), xEv);       // :394
      xE = x_hsum(xEv);                                                 // :401-403
      xN = __fmul_rn(xN, ploop);                                        // :405-408
      xC = __fadd_rn(__fmul_rn(xC, ploop), __fmul_rn(xE, p.s.xf_E_move));
      xJ = __fadd_rn(__fmul_rn(xJ, ploop), __fmul_rn(xE, p.s.xf_E_loop));
      xB = __fadd_rn(__fmul_rn(xJ, pmove), __fmul_rn(xN, pmove));
      float scale = 1.0f;
      if (xE > 1.0e4f) {                                                // :412-428
        xN = __fdiv_rn(xN, xE); xC = __fdiv_rn(xC, xE); xJ = __fdiv_rn(xJ, xE); xB = __fdiv_rn(xB, xE);
        const float inv = (float) (1.0 / (double) xE);
        for (int q = 0; q < Q; q++) { ST(q, 0) = __fmul_rn(ST(q, 0), inv); ST(q, 1) = __fmul_rn(ST(q, 1), inv); ST(q, 2) = __fmul_rn(ST(q, 2), inv); }
        scale = xE;
        totscale = (float) ((double) totscale + log((double) xE));
        xE = 1.0f;
      }
      if (act) x_store_row(xrow + (size_t) i * 6, z, xE, xN, xJ, xB, xC, scale);
      if (act) i++;
    }
    if (have && i > L) {                                                // :456-461
      if (z == 0) {
        int status = P7GPU_OK; float sc;
        if (isnan(xC) || (L > 0 && xC == 0.0f) || isinf(xC)) { status = P7GPU_ERANGE; sc = xC; }
        else sc = (float) ((double) totscale + log((double) __fmul_rn(xC, pmove)));
        p.c.sc[out] = sc; p.c.status[out] = status;
        if (p.totscale) p.totscale[out] = totscale;
      }
      have = false;
    }
  }
}

// one pass of "extend the DD component" of the Backward row (fwdback.c:519-527, :634-643): dcv <- dcv * tDD(q); D(q) += dcv
#define BWD_DD_PASSES(cond)                                                           \
  for (int j = 1; j < 4; j++) {                                                       \
    dcv = x_lshift(dcv, z);                                                           \
    for (int q = Q - 1; q >= 0; q--) { dcv = __fmul_rn(dcv, TDD(q)); const float d_ = __fadd_rn(ST(q, 1), dcv); if (cond) ST(q, 1) = d_; } \
  }
// M(q) += D(next node) * tMD(q)   (:529-535, :646-652)
#define BWD_MD_ADD(cond)                                                              \
  {                                                                                   \
    float dn = x_lshift(ST(0, 1), z);                                                 \
    for (int q = Q - 1; q >= 0; q--) { const float m_ = __fadd_rn(ST(q, 0), __fmul_rn(dn, TMD(q))); if (cond) ST(q, 0) = m_; dn = ST(q, 1); } \
  }

template <bool GS>
__global__ void __launch_bounds__(32) fbx_bwd_kernel(const FbxParams p)
{
  extern __shared__ float s_state[];
  float *st = GS ? p.gstate + (size_t) blockIdx.x * p.Q * 96 : s_state;
  const int lane = threadIdx.x, z = lane & 3, gl = lane & ~3;
  const int Q = p.Q;
  const float4 *tf4 = reinterpret_cast<const float4 *>(p.tfx) + z * 2;

  bool have = false, exhausted = false, own = false;
  long long out = 0;
  const uint8_t *seq = nullptr; float *xrow = nullptr; const float *frow = nullptr;
  int L = 0, i = 0;
  float xE = 0.f, xN = 0.f, xJ = 0.f, xB = 0.f, xC = 0.f, pmove = 0.f, ploop = 0.f, totscale = 0.f;

  for (;;) {
    {
      const bool need = !have && !exhausted;
      long long s;
      const bool got = fbx_draw(p.c, need, z, gl, &s, &out);
      if (need && !got) exhausted = true;
      // ---------------- row L (fwdback.c:493-563), executed by the groups that just drew a target ----------------
      if (__any_sync(FULL, got)) {
        if (got) {
          L = p.c.len[s]; seq = p.c.res + p.c.off[s];
          pmove = p.c.lc[L].pmove; ploop = __fsub_rn(1.0f, pmove);
          xrow = p.xmx ? p.xmx + p.xmx_off[out] : nullptr;
          frow = p.fwd_xmx + p.xmx_off[out];
          own = false; have = true;
          xJ = 0.f; xB = 0.f; xN = 0.f;
          xC = pmove;
          xE = __fmul_rn(xC, p.s.xf_E_move);
          for (int q = 0; q < Q; q++) { ST(q, 0) = xE; ST(q, 1) = xE; ST(q, 2) = 0.f; }
        }
        float dcv = 0.f;
        float dpv = x_lshift(ST(Q - 1, 1), z);
        for (int q = Q - 1; q >= 0; q--) { dcv = __fmul_rn(dpv, TDD(q)); dpv = __fadd_rn(ST(q, 1), dcv); if (got) ST(q, 1) = dpv; }
        BWD_DD_PASSES(got)                                               // every lane runs the shuffles; only the new targets store
        BWD_MD_ADD(got)
        if (got) {
          const float scale = frow[(size_t) L * 6 + 5];
          if (scale > 1.0f) {                                            // :538-551
            xE = __fdiv_rn(xE, scale); xN = __fdiv_rn(xN, scale); xC = __fdiv_rn(xC, scale); xJ = __fdiv_rn(xJ, scale); xB = __fdiv_rn(xB, scale);
            const float inv = (float) (1.0 / (double) scale);
            for (int q = 0; q < Q; q++) { ST(q, 0) = __fmul_rn(ST(q, 0), inv); ST(q, 1) = __fmul_rn(ST(q, 1), inv); ST(q, 2) = __fmul_rn(ST(q, 2), inv); }
          }
          totscale = (float) log((double) scale);                        // :553
          if (xrow) x_store_row(xrow + (size_t) L * 6, z, xE, xN, xJ, xB, xC, scale);
          i = L - 1;
        }
      }
    }
    if (__any_sync(FULL, have && i < 1)) {
      // ---------------- termination at row 0 (:714-733) ----------------
      const bool fin = have && i < 1;
      const int x = (fin && L > 0) ? (int) seq[0] : 0;
      const float *rf = p.rfx + ((size_t) x * 4 + z) * p.Qp;
      float xBv = 0.f;
      for (int q = 0; q < Q; q++) {
        float mpv = __fmul_rn(ST(q, 0), __ldg(rf + q));
        mpv = __fmul_rn(mpv, __ldg(p.tfx + ((size_t) q * 4 + z) * 8));
        xBv = __fadd_rn(xBv, mpv);
      }
      const float xB0 = x_hsum(xBv);
      if (fin) {
        xB = xB0;
        xN = __fadd_rn(__fmul_rn(xB, pmove), __fmul_rn(xN, ploop));
        if (xrow) x_store_row(xrow, z, 0.0f, xN, 0.0f, xB, 0.0f, 1.0f);
        if (z == 0) {
          int status = P7GPU_OK; float sc;
          if (isnan(xN) || (L > 0 && xN == 0.0f) || isinf(xN)) { status = P7GPU_ERANGE; sc = xN; }
          else sc = (float) ((double) totscale + log((double) xN));
          p.c.sc[out] = sc; p.c.status[out] = status;
          if (p.totscale) p.totscale[out] = totscale;
          if (p.own) p.own[out] = own ? 1 : 0;
        }
        have = false;
      }
    }
    if (!__any_sync(FULL, have)) { if (__all_sync(FULL, exhausted)) break; continue; }

    if (__any_sync(FULL, have && i >= 1)) {
      // ---------------- one DP row i (fwdback.c:566-711) ----------------
      const bool act = have && i >= 1;
      const int x = act ? (int) seq[i] : 0;                              // dsq[i+1]
      const float fscale = act ? frow[(size_t) i * 6 + 5] : 1.0f;
      const float *rf = p.rfx + ((size_t) x * 4 + z) * p.Qp;
      // phase 1 (:575-606)
      float tmmv, timv, tdmv;
      {
        const float4 t0 = __ldg(tf4);                                    // vector 0: BM MM IM DM
        tmmv = x_lshift(t0.y, z); timv = x_lshift(t0.z, z); tdmv = x_lshift(t0.w, z);
      }
      float mpv = x_lshift(__fmul_rn(ST(0, 0), __ldg(rf)), z);
      float xBv = 0.f;
      for (int q = Q - 1; q >= 0; q--) {
        const float4 ta = __ldg(tf4 + q * 8), tb = __ldg(tf4 + q * 8 + 1);
        const float ipv = ST(q, 2);
        ST(q, 2) = __fadd_rn(__fmul_rn(ipv, tb.z), __fmul_rn(mpv, timv));          // :589
        ST(q, 1) = __fmul_rn(mpv, tdmv);                                            // :590
        const float mcv = __fadd_rn(__fmul_rn(ipv, tb.y), __fmul_rn(mpv, tmmv));   // :591
        mpv = __fmul_rn(ST(q, 0), __ldg(rf + q));                                   // :593
        ST(q, 0) = mcv;
        tdmv = ta.w; timv = ta.z; tmmv = ta.y;                                      // :596-598
        xBv = __fadd_rn(xBv, __fmul_rn(mpv, ta.x));                                 // :600
      }
      // phase 2 (:609-619)
      xB = x_hsum(xBv);
      xC = __fmul_rn(xC, ploop);
      xJ = __fadd_rn(__fmul_rn(xB, pmove), __fmul_rn(xJ, ploop));
      xN = __fadd_rn(__fmul_rn(xB, pmove), __fmul_rn(xN, ploop));
      xE = __fadd_rn(__fmul_rn(xC, p.s.xf_E_move), __fmul_rn(xJ, p.s.xf_E_loop));
      // phase 3 (:622-632)
      float dcv = 0.f;
      {
        float dpv = x_lshift(__fadd_rn(ST(0, 1), xE), z);
        for (int q = Q - 1; q >= 0; q--) {
          dcv = __fmul_rn(dpv, TDD(q));
          dpv = __fadd_rn(ST(q, 1), __fadd_rn(dcv, xE));
          ST(q, 1) = dpv;
          ST(q, 0) = __fadd_rn(ST(q, 0), xE);
        }
      }
      BWD_DD_PASSES(true)                                                // phase 4 (:634-643)
      BWD_MD_ADD(true)                                                   // phase 5 (:646-652)
      // sparse rescaling (:661-680)
      if ((double) xB > 1.0e16) own = true;
      float scale;
      if (own) scale = (xB > 1.0e4f) ? xB : 1.0f;
      else     scale = fscale;
      if (scale > 1.0f) {
        xE = __fdiv_rn(xE, scale); xN = __fdiv_rn(xN, scale); xJ = __fdiv_rn(xJ, scale); xB = __fdiv_rn(xB, scale); xC = __fdiv_rn(xC, scale);
        const float inv = (float) (1.0 / (double) scale);
        for (int q = 0; q < Q; q++) { ST(q, 0) = __fmul_rn(ST(q, 0), inv); ST(q, 1) = __fmul_rn(ST(q, 1), inv); ST(q, 2) = __fmul_rn(ST(q, 2), inv); }
        if (act) totscale = (float) ((double) totscale + log((double) scale));
      }
      if (act && xrow) x_store_row(xrow + (size_t) i * 6, z, xE, xN, xJ, xB, xC, scale);
      if (act) i--;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
static int fbx_fill(p7gpu_ctx *ctx, const Lane &ln, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
                    float *d_sc, int32_t *d_status, int counter_slot, FbxParams *p)
{
  if (!prof->d_tfx || !prof->d_rfx) { set_error("exact-order Forward/Backward: the profile was uploaded without its Forward part"); return P7GPU_EINVAL; }
  P7G_CHECK(ensure_len_table(prof, db.max_length));
  p->c.res = db.res; p->c.off = db.off; p->c.len = db.len; p->c.order = d_idx ? nullptr : db.order; p->c.idx = d_idx; p->c.first = db.first;
  p->c.n = n; p->c.lc = prof->d_len; p->c.sc = d_sc; p->c.status = d_status;
  p->c.work_counter = (unsigned long long *) (ln.counters + counter_slot);
  p->tfx = prof->d_tfx; p->rfx = prof->d_rfx; p->Q = prof->fbxQ; p->Qp = prof->fbxQp; p->M = prof->M; p->s = prof->fwd;
  p->gstate = nullptr;
  P7G_CUDA(cudaMemsetAsync(ln.counters + counter_slot, 0, 2 * sizeof(unsigned int), ln.stream));
  return P7GPU_OK;
}

// grid + where the DP row lives: shared memory while at least two warps fit on an SM, else an L2-resident global scratch
template <class K>
static int fbx_launch(p7gpu_ctx *ctx, cudaStream_t stream, int kind, K kern_smem, K kern_glob, FbxParams &p)
{
  const size_t row_bytes = (size_t) p.Q * 96 * sizeof(float);
  const long long groups = (p.c.n + 7) / 8;
  if (!ctx->fbx_global && row_bytes * 2 <= (size_t) ctx->smem_optin - 2048) {
    P7G_CUDA(cudaFuncSetAttribute(kern_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) row_bytes));
    int occ = 0;
    P7G_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern_smem, 32, row_bytes));
    if (occ < 1) { set_error("fbx: zero occupancy (Q=%d)", p.Q); return P7GPU_ESYS; }
    long long blocks = std::min<long long>((long long) ctx->sm_count * occ, std::max<long long>(groups, 1));
    { LaunchTimer t(ctx, kind, stream); kern_smem<<<(unsigned) blocks, 32, row_bytes, stream>>>(p); }
  } else {
    // keep the rows of all resident warps inside ~96 MB of the 126 MB L2
    long long per_sm = (long long) (((size_t) 96 << 20) / (row_bytes * (size_t) ctx->sm_count));
    per_sm = std::max<long long>(1, std::min<long long>(per_sm, 16));
    if (ctx->fbx_warps > 0) per_sm = ctx->fbx_warps;
    long long blocks = std::min<long long>((long long) ctx->sm_count * per_sm, std::max<long long>(groups, 1));
    P7G_CHECK(ensure_fbx_state(ctx, (size_t) blocks * row_bytes));
    p.gstate = (float *) ctx->d_fbx_state;
    { LaunchTimer t(ctx, kind, stream); kern_glob<<<(unsigned) blocks, 32, 0, stream>>>(p); }
  }
  P7G_CUDA(cudaGetLastError());
  ctx->launches++;
  return P7GPU_OK;
}

int launch_fbx_fwd(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
                   float *d_sc, int32_t *d_status, float *d_xmx, const int64_t *d_xmx_off, float *d_totscale)
{
  if (n <= 0) return P7GPU_OK;
  if (!d_xmx || !d_xmx_off) { set_error("exact-order Forward needs the specials buffer"); return P7GPU_EINVAL; }
  const Lane &ln = ctx->lanes[lane];
  FbxParams p;
  P7G_CHECK(fbx_fill(ctx, ln, prof, db, d_idx, n, d_sc, d_status, CTR_FWD_WORK, &p));
  p.xmx = d_xmx; p.fwd_xmx = nullptr; p.xmx_off = d_xmx_off; p.totscale = d_totscale; p.own = nullptr;
  return fbx_launch(ctx, ln.stream, P7GPU_KERNEL_FWDX, fbx_fwd_kernel<false>, fbx_fwd_kernel<true>, p);
}

int launch_fbx_bwd(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
                   const float *d_fwd_xmx, const int64_t *d_xmx_off, float *d_sc, int32_t *d_status,
                   float *d_bck_xmx, float *d_totscale, int32_t *d_own)
{
  if (n <= 0) return P7GPU_OK;
  if (!d_fwd_xmx || !d_xmx_off) { set_error("exact-order Backward needs the Forward specials"); return P7GPU_EINVAL; }
  const Lane &ln = ctx->lanes[lane];
  FbxParams p;
  P7G_CHECK(fbx_fill(ctx, ln, prof, db, d_idx, n, d_sc, d_status, CTR_BWD_WORK, &p));
  p.xmx = d_bck_xmx; p.fwd_xmx = d_fwd_xmx; p.xmx_off = d_xmx_off; p.totscale = d_totscale; p.own = d_own;
  return fbx_launch(ctx, ln.stream, P7GPU_KERNEL_BWDX, fbx_bwd_kernel<false>, fbx_bwd_kernel<true>, p);
}

}  // namespace p7g
