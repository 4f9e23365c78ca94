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
//     (impl_sse.h:27).  Here FOUR GPU LANES play the four vector elements of one target ("group"); a warp carries G groups
//     (G = 8, 4, 2 or 1, see below).  _mm_add_ps/_mm_mul_ps become __fadd_rn/__fmul_rn per lane (never contracted to FMA; the
//     library is built -ftz=true: the reference runs FTZ + DAZ, impl_sse.h:356-376), the element shifts
//     (esl_sse_rightshift_ps, the move_ss + shuffle left shift) are width-4 warp shuffles, the horizontal sums are two
//     xor-shuffles in the reference's (x0+x1)+(x2+x3) association, the "did any DD path change a D cell" test of :375-389 is a
//     ballot over the four lanes.
//   * The one DP row the parsers keep (M, D, I per vector) lives in shared memory, [q][3][4G] (conflict-free).  Per target that
//     is 48*Q bytes, which bounds the targets in flight per SM (47 at M=400, 9 at M=2000).  The work of one target-row is a
//     fixed sequence of instructions whatever the number of lanes that execute it, so what matters is (a) targets in flight
//     and (b) lanes served per instruction: G is the largest of 8, 4, 2, 1 whose DP rows fit (fbx_groups below has the measurements).
//   * The profile is read in the reference's own striped values (om->rfv, om->tfv through the view), re-laid out only for
//     load width: tfx[q][z][8] = {BM,MM,IM,DM,MD,MI,II,DD} (two 16-byte loads per vector), rfx[x][z][Qp], tddx/tmdx[z][Qp]
//     (one 16-byte load per four vectors).
//   * Loops over q run four vectors per trip with every load of the trip issued before its first use; per element the
//     operations and their order are the reference's.  Fused without changing any element's operation sequence: the sum of
//     the D cells into xE rides along the DD passes (Forward); "M += xE" and "M += D(next)*tMD" ride along the last DD pass
//     (Backward).
//   * Targets of different lengths share a warp: the warp loops over ROWS, every group with a target advances its own
//     target by one row per iteration and draws a new target when it finishes one.
//
// The double-precision log() of the scale factors (totscale, fwdback.c:434,:689) is the only libm call; the device value
// is provisional — whenever specials are returned to the host, the host recomputes totscale and the score from them with
// its own libm in the reference's expressions (p7gpu_host.cu: fbx_host_*_score), so nothing reported depends on device libm.
#include "dp_common.cuh"
#include <math_constants.h>

namespace p7g {

struct FbxParams {
  DpCommon c;
  const float *tfx;       // [Q][4][8]
  const float *rfx;       // [NTABROWS][4][Qp]
  const float *tddx;      // [4][Qp]  tDD of vector q, element z
  const float *tmdx;      // [4][Qp]  tMD
  int Q, Qp, M;
  FwdScalars s;
  float *xmx;             // specials of this direction, [(L+1)*6] per target at xmx_off[out] (backward: may be null)
  const float *fwd_xmx;   // backward only
  const int64_t *xmx_off;
  float *totscale; int32_t *own;
};

template <int G> struct FbxGeom {
  static constexpr int SW = 4 * G;                                       // active lanes = state columns
  static constexpr unsigned MASK = (G == 8) ? 0xffffffffu : ((1u << (4 * G)) - 1u);
};

#define X_RSHIFT(v)  ({ const float r_ = __shfl_up_sync(MASK, (v), 1, 4);   (z == 0) ? 0.0f : r_; })   /* esl_sse_rightshift_ps(v, zerov) */
#define X_LSHIFT(v)  ({ const float r_ = __shfl_down_sync(MASK, (v), 1, 4); (z == 3) ? 0.0f : r_; })   /* move_ss + shuffle(0,3,2,1) */
#define X_HSUM(x)    ({ const float a_ = __fadd_rn((x), __shfl_xor_sync(MASK, (x), 1)); __fadd_rn(a_, __shfl_xor_sync(MASK, a_, 2)); })

__device__ __forceinline__ void x_store_row(float *row, int z, float xE, float xN, float xJ, float xB, float xC, float scale)
{
  float2 *r2 = reinterpret_cast<float2 *>(row);
  if (z == 0) r2[0] = make_float2(xE, xN); else if (z == 1) r2[1] = make_float2(xJ, xB); else if (z == 2) r2[2] = make_float2(xC, scale);
}

// state cell (vector q, state s = 0 M / 1 D / 2 I) of this lane; SQ(u, s): the same relative to the trip pointer sp
#define ST(q, s)   st[((q) * 3 + (s)) * SW]
#define SQ(u, s)   sp[((u) * 3 + (s)) * SW]
#define TDD(q)     __ldg(p.tfx + ((size_t) (q) * 4 + z) * 8 + 7)
#define TMD(q)     __ldg(p.tfx + ((size_t) (q) * 4 + z) * 8 + 4)

// draw the next target for the groups that need one (valid on every lane of the group)
template <unsigned MASK>
__device__ __forceinline__ bool fbx_draw(const DpCommon &c, bool need, int z, int gl, long long *s, long long *out)
{
  unsigned long long w = 0;
  if (need && z == 0) w = atomicAdd(c.work_counter, 1ULL);
  w = __shfl_sync(MASK, w, gl);
  if (!need || w >= (unsigned long long) c.n) return false;
  *s   = c.idx ? c.idx[w] : (c.order ? (long long) c.order[w] : c.first + (long long) w);
  *out = c.idx ? (long long) w : *s;
  return true;
}

// ---- Forward: one trip of four vectors of the main loop (fwdback.c:318-347); K = number of valid vectors in the trip ----
#define FWD_MAIN_TRIP(K)                                                                                 \
  {                                                                                                      \
    const float4 e4 = __ldg(reinterpret_cast<const float4 *>(rf + q0));                                  \
    const float ee[4] = {e4.x, e4.y, e4.z, e4.w};                                                        \
    float4 ta[4], tb[4]; float pm[4], pd[4], pi[4];                                                      \
    _Pragma("unroll") for (int u = 0; u < (K); u++) {                                                    \
      ta[u] = __ldg(tq + u * 8); tb[u] = __ldg(tq + u * 8 + 1);                                          \
      pm[u] = SQ(u, 0); pd[u] = SQ(u, 1); pi[u] = SQ(u, 2);                                              \
    }                                                                                                    \
    _Pragma("unroll") for (int u = 0; u < (K); u++) {                                                    \
      float sv = __fmul_rn(xBv, ta[u].x);                                                                \
      sv = __fadd_rn(sv, __fmul_rn(mpv, ta[u].y));                                                       \
      sv = __fadd_rn(sv, __fmul_rn(ipv, ta[u].z));                                                       \
      sv = __fadd_rn(sv, __fmul_rn(dpv, ta[u].w));                                                       \
      sv = __fmul_rn(sv, ee[u]);                                                                         \
      xEv = __fadd_rn(xEv, sv);                                                                          \
      mpv = pm[u]; dpv = pd[u]; ipv = pi[u];                                                             \
      SQ(u, 0) = sv; SQ(u, 1) = dcv;                                                                     \
      dcv = __fmul_rn(sv, tb[u].x);                                                                      \
      SQ(u, 2) = __fadd_rn(__fmul_rn(mpv, tb[u].y), __fmul_rn(ipv, tb[u].z));                            \
    }                                                                                                    \
  }
// first full DD pass (:358-366): D(q) = dcv + D(q); dcv = D(q) * tDD(q)
#define FWD_PASS0_TRIP(K)                                                                                \
  {                                                                                                      \
    const float4 t4 = __ldg(reinterpret_cast<const float4 *>(tdd + q0));                                 \
    const float tt[4] = {t4.x, t4.y, t4.z, t4.w};                                                        \
    float d[4];                                                                                          \
    _Pragma("unroll") for (int u = 0; u < (K); u++) d[u] = SQ(u, 1);                                     \
    if (q0 == 0) d[0] = 0.0f;                                                                            \
    _Pragma("unroll") for (int u = 0; u < (K); u++) { const float dd = __fadd_rn(dcv, d[u]); SQ(u, 1) = dd; dcv = __fmul_rn(dd, tt[u]); } \
  }
// later DD passes (:368-391): sv = dcv + D(q); changed |= sv > D(q); D(q) = sv; dcv *= tDD(q); the row sum xs rides along
#define FWD_PASSJ_TRIP(K)                                                                                \
  {                                                                                                      \
    const float4 t4 = __ldg(reinterpret_cast<const float4 *>(tdd + q0));                                 \
    const float tt[4] = {t4.x, t4.y, t4.z, t4.w};                                                        \
    float d[4];                                                                                          \
    _Pragma("unroll") for (int u = 0; u < (K); u++) d[u] = SQ(u, 1);                                     \
    _Pragma("unroll") for (int u = 0; u < (K); u++) {                                                    \
      const float sv = __fadd_rn(dcv, d[u]);                                                             \
      if (sv > d[u]) changed = 1;                                                                        \
      if (more) SQ(u, 1) = sv;                                                                           \
      xs = __fadd_rn(sv, xs);                                                                            \
      dcv = __fmul_rn(dcv, tt[u]);                                                                       \
    }                                                                                                    \
  }

template <int G>
__global__ void __launch_bounds__(32) fbx_fwd_kernel(const FbxParams p)
{
  constexpr int SW = FbxGeom<G>::SW; constexpr unsigned MASK = FbxGeom<G>::MASK;
  extern __shared__ float s_state[];
  const int lane = threadIdx.x;
  if (lane >= SW) return;                                                // every sync primitive below names MASK only
  float *st = s_state + lane;
  const int z = lane & 3, gl = lane & ~3;
  const int Q = p.Q, Q4 = p.Q & ~3, QT = p.Q & 3;
  const float4 *tf4 = reinterpret_cast<const float4 *>(p.tfx) + z * 2;
  const float *tdd = p.tddx + (size_t) z * p.Qp;
  const bool early = p.M >= 100;                                         // fwdback.c:352: full serialisation below M = 100

  bool have = false, exhausted = false;
  long long out = 0;
  const uint8_t *seq = nullptr; float *xrow = nullptr;
  int L = 0, i = 0, xnext = 0;
  float xE = 0.f, xN = 0.f, xJ = 0.f, xB = 0.f, xC = 0.f, pmove = 0.f, ploop = 0.f, totscale = 0.f;

  for (;;) {
    {
      const bool need = !have && !exhausted;
      long long s;
      const bool got = fbx_draw<MASK>(p.c, need, z, gl, &s, &out);
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
    if (!__any_sync(MASK, have)) { if (__all_sync(MASK, exhausted)) break; continue; }

    if (__any_sync(MASK, have && i <= L)) {
      // ---------------- one DP row (fwdback.c:303-452) for every group that has a target ----------------
      const bool act = have && i <= L;
      const int x = act ? xnext : 0;
      xnext = (act && i < L) ? (int) seq[i] : 0;
      const float *rf = p.rfx + ((size_t) x * 4 + z) * p.Qp;
      float mpv = X_RSHIFT(ST(Q - 1, 0)), dpv = X_RSHIFT(ST(Q - 1, 1)), ipv = X_RSHIFT(ST(Q - 1, 2));
      float dcv = 0.f, xEv = 0.f;
      const float xBv = xB;
      {
        float *sp = st; const float4 *tq = tf4; int q0 = 0;
        for (; q0 < Q4; q0 += 4, sp += 12 * SW, tq += 32) FWD_MAIN_TRIP(4)
        if (QT == 1) FWD_MAIN_TRIP(1) else if (QT == 2) FWD_MAIN_TRIP(2) else if (QT == 3) FWD_MAIN_TRIP(3)
      }
      dcv = X_RSHIFT(dcv);
      {
        float *sp = st; int q0 = 0;
        for (; q0 < Q4; q0 += 4, sp += 12 * SW) FWD_PASS0_TRIP(4)
        if (QT == 1) FWD_PASS0_TRIP(1) else if (QT == 2) FWD_PASS0_TRIP(2) else if (QT == 3) FWD_PASS0_TRIP(3)
      }
      // up to three more passes that extend only the DD component (:352-391).  The sum of the D cells into xE (:394) rides
      // along: it is only valid for the pass that turns out to be the group's last one (no cell changed, or the third).
      bool more = act;
      float xEfin = xEv;
      for (int j = 1; j < 4; j++) {
        if (!__any_sync(MASK, more)) break;
        dcv = X_RSHIFT(dcv);
        int changed = 0;
        float xs = xEv;
        {
          float *sp = st; int q0 = 0;
          for (; q0 < Q4; q0 += 4, sp += 12 * SW) FWD_PASSJ_TRIP(4)
          if (QT == 1) FWD_PASSJ_TRIP(1) else if (QT == 2) FWD_PASSJ_TRIP(2) else if (QT == 3) FWD_PASSJ_TRIP(3)
        }
        bool grp_changed = true;
        if (early) grp_changed = ((__ballot_sync(MASK, changed != 0) >> gl) & 0xfu) != 0;   // :389 "DD's didn't change any DMO(q)? Then done"
        if (more && (j == 3 || !grp_changed)) xEfin = xs;
        if (!grp_changed) more = false;
      }
      xEv = xEfin;
      xE = X_HSUM(xEv);                                                 // :401-403
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

// ---- Backward, row L only (once per target; plain loops): the DD passes (:519-527) and the M<-D step (:529-535) ----
#define BWD_INIT_DD_PASSES(cond)                                                      \
  for (int j = 1; j < 4; j++) {                                                       \
    dcv = X_LSHIFT(dcv);                                                              \
    for (int q = Q - 1; q >= 0; q--) { dcv = __fmul_rn(dcv, TDD(q)); const float d_ = __fadd_rn(ST(q, 1), dcv); if (cond) ST(q, 1) = d_; } \
  }
#define BWD_INIT_MD_ADD(cond)                                                         \
  {                                                                                   \
    float dn = X_LSHIFT(ST(0, 1));                                                    \
    for (int q = Q - 1; q >= 0; q--) { const float m_ = __fadd_rn(ST(q, 0), __fmul_rn(dn, TMD(q))); if (cond) ST(q, 0) = m_; dn = ST(q, 1); } \
  }
// ---- Backward main rows: trips of four vectors, q descending; K valid vectors (the top trip may be partial) ----
#define BWD_MAIN_TRIP(K)                                                                                 \
  {                                                                                                      \
    const float4 e4 = __ldg(reinterpret_cast<const float4 *>(rf + q0));                                  \
    const float ee[4] = {e4.x, e4.y, e4.z, e4.w};                                                        \
    float4 ta[4], tb[4]; float pm[4], pi[4];                                                             \
    _Pragma("unroll") for (int u = 0; u < (K); u++) {                                                    \
      ta[u] = __ldg(tq + u * 8); tb[u] = __ldg(tq + u * 8 + 1);                                          \
      pm[u] = SQ(u, 0); pi[u] = SQ(u, 2);                                                                \
    }                                                                                                    \
    _Pragma("unroll") for (int u = (K) - 1; u >= 0; u--) {                                               \
      const float ipv = pi[u];                                                                           \
      SQ(u, 2) = __fadd_rn(__fmul_rn(ipv, tb[u].z), __fmul_rn(mpv, timv));          /* :589 */           \
      SQ(u, 1) = __fmul_rn(mpv, tdmv);                                               /* :590 */           \
      const float mcv = __fadd_rn(__fmul_rn(ipv, tb[u].y), __fmul_rn(mpv, tmmv));   /* :591 */           \
      mpv = __fmul_rn(pm[u], ee[u]);                                                 /* :593 */           \
      SQ(u, 0) = mcv;                                                                                    \
      tdmv = ta[u].w; timv = ta[u].z; tmmv = ta[u].y;                                /* :596-598 */       \
      xBv = __fadd_rn(xBv, __fmul_rn(mpv, ta[u].x));                                 /* :600 */           \
    }                                                                                                    \
  }
#define BWD_PHASE3_TRIP(K)                                                                               \
  {                                                                                                      \
    const float4 t4 = __ldg(reinterpret_cast<const float4 *>(tdd + q0));                                 \
    const float tt[4] = {t4.x, t4.y, t4.z, t4.w};                                                        \
    float d[4];                                                                                          \
    _Pragma("unroll") for (int u = 0; u < (K); u++) d[u] = SQ(u, 1);                                     \
    _Pragma("unroll") for (int u = (K) - 1; u >= 0; u--) { dcv = __fmul_rn(dpv, tt[u]); dpv = __fadd_rn(d[u], __fadd_rn(dcv, xE)); SQ(u, 1) = dpv; } \
  }
#define BWD_PASSJ_TRIP(K)                                                                                \
  {                                                                                                      \
    const float4 t4 = __ldg(reinterpret_cast<const float4 *>(tdd + q0));                                 \
    const float tt[4] = {t4.x, t4.y, t4.z, t4.w};                                                        \
    float d[4];                                                                                          \
    _Pragma("unroll") for (int u = 0; u < (K); u++) d[u] = SQ(u, 1);                                     \
    _Pragma("unroll") for (int u = (K) - 1; u >= 0; u--) { dcv = __fmul_rn(dcv, tt[u]); SQ(u, 1) = __fadd_rn(d[u], dcv); } \
  }
// last DD pass with "M += xE" (phase 3) and "M += D(next)*tMD" (phase 5) fused in; vector Q-1 gets its phase-5 term after the loop
#define BWD_LAST_TRIP(K)                                                                                 \
  {                                                                                                      \
    const float4 t4 = __ldg(reinterpret_cast<const float4 *>(tdd + q0));                                 \
    const float4 m4 = __ldg(reinterpret_cast<const float4 *>(tmd + q0));                                 \
    const float tt[4] = {t4.x, t4.y, t4.z, t4.w}, tm[4] = {m4.x, m4.y, m4.z, m4.w};                      \
    float d[4], mm[4];                                                                                   \
    _Pragma("unroll") for (int u = 0; u < (K); u++) { d[u] = SQ(u, 1); mm[u] = SQ(u, 0); }               \
    _Pragma("unroll") for (int u = (K) - 1; u >= 0; u--) {                                               \
      dcv = __fmul_rn(dcv, tt[u]);                                                                       \
      const float df = __fadd_rn(d[u], dcv);                                                             \
      SQ(u, 1) = df;                                                                                     \
      float mn = __fadd_rn(mm[u], xE);                                                                   \
      if (q0 + u < Q - 1) mn = __fadd_rn(mn, __fmul_rn(dn, tm[u]));                                      \
      SQ(u, 0) = mn;                                                                                     \
      dn = df;                                                                                           \
    }                                                                                                    \
  }
// run TRIP over all vectors, q descending: the partial top trip first, then full trips
#define BWD_SWEEP(TRIP)                                                                                  \
  {                                                                                                      \
    int q0 = Q4; float *sp = st + (size_t) Q4 * 3 * SW; const float4 *tq = tf4 + (size_t) Q4 * 8; (void) tq; \
    if (QT == 1) TRIP(1) else if (QT == 2) TRIP(2) else if (QT == 3) TRIP(3)                             \
    for (q0 = Q4 - 4, sp -= 12 * SW, tq -= 32; q0 >= 0; q0 -= 4, sp -= 12 * SW, tq -= 32) TRIP(4)        \
  }

template <int G>
__global__ void __launch_bounds__(32) fbx_bwd_kernel(const FbxParams p)
{
  constexpr int SW = FbxGeom<G>::SW; constexpr unsigned MASK = FbxGeom<G>::MASK;
  extern __shared__ float s_state[];
  const int lane = threadIdx.x;
  if (lane >= SW) return;
  float *st = s_state + lane;
  const int z = lane & 3, gl = lane & ~3;
  const int Q = p.Q, Q4 = p.Q & ~3, QT = p.Q & 3;
  const float4 *tf4 = reinterpret_cast<const float4 *>(p.tfx) + z * 2;
  const float *tdd = p.tddx + (size_t) z * p.Qp, *tmd = p.tmdx + (size_t) z * p.Qp;

  bool have = false, exhausted = false, own = false;
  long long out = 0;
  const uint8_t *seq = nullptr; float *xrow = nullptr; const float *frow = nullptr;
  int L = 0, i = 0;
  float xE = 0.f, xN = 0.f, xJ = 0.f, xB = 0.f, xC = 0.f, pmove = 0.f, ploop = 0.f, totscale = 0.f;

  for (;;) {
    {
      const bool need = !have && !exhausted;
      long long s;
      const bool got = fbx_draw<MASK>(p.c, need, z, gl, &s, &out);
      if (need && !got) exhausted = true;
      // ---------------- row L (fwdback.c:493-563), executed by the groups that just drew a target ----------------
      if (__any_sync(MASK, got)) {
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
        float dpv = X_LSHIFT(ST(Q - 1, 1));
        for (int q = Q - 1; q >= 0; q--) { dcv = __fmul_rn(dpv, TDD(q)); dpv = __fadd_rn(ST(q, 1), dcv); if (got) ST(q, 1) = dpv; }
        BWD_INIT_DD_PASSES(got)                                          // every lane runs the shuffles; only the new targets store
        BWD_INIT_MD_ADD(got)
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
    if (__any_sync(MASK, have && i < 1)) {
      // ---------------- termination at row 0 (:714-733): before the row block, which overwrites the state of every lane ----------------
      const bool fin = have && i < 1;
      const int x = (fin && L > 0) ? (int) seq[0] : 0;
      const float *rf = p.rfx + ((size_t) x * 4 + z) * p.Qp;
      float xBv = 0.f;
      for (int q = 0; q < Q; q++) {
        float mpv = __fmul_rn(ST(q, 0), __ldg(rf + q));
        mpv = __fmul_rn(mpv, __ldg(p.tfx + ((size_t) q * 4 + z) * 8));
        xBv = __fadd_rn(xBv, mpv);
      }
      const float xB0 = X_HSUM(xBv);
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
    if (!__any_sync(MASK, have)) { if (__all_sync(MASK, exhausted)) break; continue; }

    if (__any_sync(MASK, have && i >= 1)) {
      // ---------------- one DP row i (fwdback.c:566-711) ----------------
      const bool act = have && i >= 1;
      const int x = act ? (int) seq[i] : 0;                              // dsq[i+1]
      const float fscale = act ? frow[(size_t) i * 6 + 5] : 1.0f;
      const float *rf = p.rfx + ((size_t) x * 4 + z) * p.Qp;
      // phase 1 (:575-606)
      float tmmv, timv, tdmv;
      {
        const float4 t0 = __ldg(tf4);                                    // vector 0: BM MM IM DM
        tmmv = X_LSHIFT(t0.y); timv = X_LSHIFT(t0.z); tdmv = X_LSHIFT(t0.w);
      }
      float mpv = X_LSHIFT(__fmul_rn(ST(0, 0), __ldg(rf)));
      float xBv = 0.f;
      BWD_SWEEP(BWD_MAIN_TRIP)
      // phase 2 (:609-619)
      xB = X_HSUM(xBv);
      xC = __fmul_rn(xC, ploop);
      xJ = __fadd_rn(__fmul_rn(xB, pmove), __fmul_rn(xJ, ploop));
      xN = __fadd_rn(__fmul_rn(xB, pmove), __fmul_rn(xN, ploop));
      xE = __fadd_rn(__fmul_rn(xC, p.s.xf_E_move), __fmul_rn(xJ, p.s.xf_E_loop));
      // phase 3 (:622-632): D(q) += D(next)*tDD(q) + xE.  Its "M(q) += xE" moves into the last pass below (same operand order).
      float dcv = 0.f;
      {
        float dpv = X_LSHIFT(__fadd_rn(ST(0, 1), xE));
        BWD_SWEEP(BWD_PHASE3_TRIP)
      }
      // phase 4 (:634-643), passes 1 and 2: dcv <- dcv * tDD(q); D(q) += dcv
      for (int j = 1; j < 3; j++) {
        dcv = X_LSHIFT(dcv);
        BWD_SWEEP(BWD_PASSJ_TRIP)
      }
      // pass 3 of phase 4 with phase 3's M(q) += xE and phase 5's M(q) += D(next)*tMD(q) (:646-652) fused in: D(q) is final as soon
      // as this pass has touched it, and q runs downwards, so D(next) is the value of the previous step; only vector Q-1 needs
      // the left-shifted final D(0) and is completed after the loop.  Per element: (M + xE) + D(next)*tMD, the reference's order.
      {
        dcv = X_LSHIFT(dcv);
        float dn = 0.f;
        BWD_SWEEP(BWD_LAST_TRIP)
        const float dl = X_LSHIFT(dn);                                   // final D(0), left-shifted
        ST(Q - 1, 0) = __fadd_rn(ST(Q - 1, 0), __fmul_rn(dl, TMD(Q - 1)));
      }
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
  p->tfx = prof->d_tfx; p->rfx = prof->d_rfx; p->tddx = prof->d_tddx; p->tmdx = prof->d_tmdx;
  p->Q = prof->fbxQ; p->Qp = prof->fbxQp; p->M = prof->M; p->s = prof->fwd;
  P7G_CUDA(cudaMemsetAsync(ln.counters + counter_slot, 0, 2 * sizeof(unsigned int), ln.stream));
  return P7GPU_OK;
}

// Groups per warp: as many as fit.  Measured (scripts/fbx_rate.py, 40,000 iid targets, Forward GCUPS at G = 1/2/4/8): M=100
// 44/69/89/101, M=400 52/65/83/124, M=800 41/56/72/68, M=2000 31/32/33/35.  With few groups per warp the kernels are bound by
// the rate at which an SM accepts shared-memory instructions (one per cycle whatever the number of lanes it serves); with
// eight they are bound by the latency of one or a few resident warps, which is the better of the two everywhere but M=800.
static int fbx_groups(const p7gpu_ctx *ctx, int Q)
{
  if (ctx->fbx_groups == 1 || ctx->fbx_groups == 2 || ctx->fbx_groups == 4 || ctx->fbx_groups == 8) return ctx->fbx_groups;
  const size_t per_target = (size_t) Q * 48;
  const long long T = (long long) (((size_t) ctx->smem_optin - 1024) / per_target);     // targets in flight per SM
  return T >= 8 ? 8 : T >= 4 ? 4 : T >= 2 ? 2 : 1;
}

template <class K>
static int fbx_launch_g(p7gpu_ctx *ctx, cudaStream_t stream, int kind, K kern, int G, const FbxParams &p)
{
  const size_t row_bytes = (size_t) p.Q * 48 * G;
  if (row_bytes > (size_t) ctx->smem_optin - 1024) { set_error("exact-order Forward/Backward: model too long (Q=%d) for the shared-memory DP row", p.Q); return P7GPU_EINVAL; }
  P7G_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) row_bytes));
  int occ = 0;
  P7G_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32, row_bytes));
  if (occ < 1) { set_error("fbx: zero occupancy (Q=%d, G=%d)", p.Q, G); return P7GPU_ESYS; }
  if (ctx->fbx_warps > 0) occ = std::min(occ, ctx->fbx_warps);
  const long long warps = (p.c.n + G - 1) / G;
  const long long blocks = std::min<long long>((long long) ctx->sm_count * occ, std::max<long long>(warps, 1));
  { LaunchTimer t(ctx, kind, stream); kern<<<(unsigned) blocks, 32, row_bytes, stream>>>(p); }
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
  switch (fbx_groups(ctx, p.Q)) {
  case 8:  return fbx_launch_g(ctx, ln.stream, P7GPU_KERNEL_FWDX, fbx_fwd_kernel<8>, 8, p);
  case 4:  return fbx_launch_g(ctx, ln.stream, P7GPU_KERNEL_FWDX, fbx_fwd_kernel<4>, 4, p);
  case 2:  return fbx_launch_g(ctx, ln.stream, P7GPU_KERNEL_FWDX, fbx_fwd_kernel<2>, 2, p);
  default: return fbx_launch_g(ctx, ln.stream, P7GPU_KERNEL_FWDX, fbx_fwd_kernel<1>, 1, p);
  }
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
  switch (fbx_groups(ctx, p.Q)) {
  case 8:  return fbx_launch_g(ctx, ln.stream, P7GPU_KERNEL_BWDX, fbx_bwd_kernel<8>, 8, p);
  case 4:  return fbx_launch_g(ctx, ln.stream, P7GPU_KERNEL_BWDX, fbx_bwd_kernel<4>, 4, p);
  case 2:  return fbx_launch_g(ctx, ln.stream, P7GPU_KERNEL_BWDX, fbx_bwd_kernel<2>, 2, p);
  default: return fbx_launch_g(ctx, ln.stream, P7GPU_KERNEL_BWDX, fbx_bwd_kernel<1>, 1, p);
  }
}

}  // namespace p7g
