// fwdbwd.cu — sm_100a kernels for p7_ForwardParser / p7_BackwardParser
// (HMMER/src/impl_sse/fwdback.c:133-144,237-249; engines :257-465 and :469-735, do_full = 0).
//
// Same blocked layout as the Viterbi kernel (dp_common.cuh): one warp per target, lane l owns nodes
// l*C+1..l*C+C, float32 in odds-ratio space, the reference's sparse rescaling rule (xE > 1e4) and its
// xmx[(L+1)*6] = {E,N,J,B,C,SCALE} output.  The serial D->D recurrence (a first-order linear
// recurrence) is evaluated exactly-in-structure by an affine prefix scan over the warp.  Compiled with
// -fmad=false -ftz=true: the reference's SSE2 build has no FMA and runs with FTZ/DAZ (impl_sse.h:356-376).
#include "dp_common.cuh"
#include "fwdbwd_mw.cuh"
#include <math_constants.h>

namespace p7g {

struct FbParams {
  DpCommon c;
  const float *rf;      // [NTABROWS][C][32]
  const float *tf;      // [8][C][32]
  FwdScalars s;
  float *xmx;           // forward: output specials (may be null); backward: output specials (may be null)
  const float *fwd_xmx; // backward only: forward specials (scale factors)
  const int64_t *xmx_off;
  float *totscale;      // per-target total log scale (may be null)
  int32_t *own;         // backward only: has_own_scales per target (may be null)
};

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

__device__ __forceinline__ void store_specials(float *row, float xE, float xN, float xJ, float xB, float xC, float scale)
{
  float2 *r2 = reinterpret_cast<float2 *>(row);
  r2[0] = make_float2(xE, xN); r2[1] = make_float2(xJ, xB); r2[2] = make_float2(xC, scale);
}

template <int C, bool SMEM>
__device__ __forceinline__ void stage_tables(const FbParams &p, float *s_fb, const float *&tf, const float *&rf)
{
  tf = p.tf; rf = p.rf;
  if (SMEM) {
    const int ntf = 8 * C * 32, nrf = NTABROWS * C * 32;
    for (int i = threadIdx.x; i < ntf; i += blockDim.x) s_fb[i] = p.tf[i];
    for (int i = threadIdx.x; i < nrf; i += blockDim.x) s_fb[ntf + i] = p.rf[i];
    __syncthreads();
    tf = s_fb; rf = s_fb + ntf;
  }
}

// ------------------------------------------------------------------------------------------------
// TREG: the per-node transition vectors live in registers (C <= 16) so that the row loop is bound by the FP32 pipe
// instead of by ~11 LDS per cell (see vit.cu).
template <int C, bool SMEM, bool TREG>
__global__ void __launch_bounds__(DP_THREADS) fwd_kernel(const FbParams p)
{
  extern __shared__ __align__(16) float s_fb[];
  const float *tf, *rf;
  stage_tables<C, SMEM>(p, s_fb, tf, rf);
  const int lane = threadIdx.x & 31;
  const float *tBM = tf + 0 * C * 32 + lane, *tMM = tf + 1 * C * 32 + lane, *tIM = tf + 2 * C * 32 + lane,
              *tDM = tf + 3 * C * 32 + lane, *tMD = tf + 4 * C * 32 + lane, *tMI = tf + 5 * C * 32 + lane,
              *tII = tf + 6 * C * 32 + lane, *tDD = tf + 7 * C * 32 + lane;
  float rBM[TREG ? C : 1], rMM[TREG ? C : 1], rIM[TREG ? C : 1], rDM[TREG ? C : 1], rMD[TREG ? C : 1], rMI[TREG ? C : 1],
        rII[TREG ? C : 1], rDD[TREG ? C : 1];
  if (TREG) {
#pragma unroll
    for (int j = 0; j < C; j++) {
      rBM[j] = tBM[j * 32]; rMM[j] = tMM[j * 32]; rIM[j] = tIM[j * 32]; rDM[j] = tDM[j * 32];
      rMD[j] = tMD[j * 32]; rMI[j] = tMI[j * 32]; rII[j] = tII[j * 32]; rDD[j] = tDD[j * 32];
    }
  }
#define T_(name, j) (TREG ? r##name[j] : t##name[(j) * 32])
  long long s, out;
  while (next_target(p.c, lane, &s, &out)) {
    const int L = p.c.len[s];
    const uint8_t *seq = p.c.res + p.c.off[s];
    const float pmove = p.c.lc[L].pmove, ploop = 1.0f - pmove;
    float *xrow = p.xmx ? p.xmx + p.xmx_off[out] : nullptr;
    float mp[C], ip[C], dp[C];
#pragma unroll
    for (int j = 0; j < C; j++) mp[j] = ip[j] = dp[j] = 0.0f;
    float xE = 0.f, xN = 1.f, xJ = 0.f, xB = pmove, xC = 0.f, totscale = 0.f;     // fwdback.c:287-294
    if (xrow && lane == 0) store_specials(xrow, xE, xN, xJ, xB, xC, 1.0f);

    for (int i = 1; i <= L; i++) {
      const int x = __ldg(seq + i - 1);
      const float *rfx = rf + (size_t) x * C * 32 + lane;
      float pm = __shfl_up_sync(0xffffffffu, mp[C - 1], 1);
      float pi = __shfl_up_sync(0xffffffffu, ip[C - 1], 1);
      float pd = __shfl_up_sync(0xffffffffu, dp[C - 1], 1);
      if (lane == 0) pm = pi = pd = 0.0f;
      float sumM = 0.f, dcv = 0.f;
#pragma unroll
      for (int j = 0; j < C; j++) {                                   // :318-345
        float sv = xB * T_(BM, j);
        sv = sv + pm * T_(MM, j);
        sv = sv + pi * T_(IM, j);
        sv = sv + pd * T_(DM, j);
        sv = sv * rfx[j * 32];
        sumM += sv;
        pm = mp[j]; pi = ip[j]; pd = dp[j];
        mp[j] = sv;
        dp[j] = dcv;                                                  // M(i,k-1) * tMD(k-1)
        dcv = sv * T_(MD, j);
        ip[j] = pm * T_(MI, j) + pi * T_(II, j);
      }
      {
        float d0 = __shfl_up_sync(0xffffffffu, dcv, 1);
        dp[0] = (lane == 0) ? 0.0f : d0;
      }
      // D->D paths (:351-399): D_k = P_k + D_{k-1} * tDD_{k-1}.  Lane map: out = A + B * carry
      float A = 0.f, B = 1.f;
#pragma unroll
      for (int j = 0; j < C; j++) { const float t = T_(DD, j); A = (dp[j] + A) * t; B = B * t; }
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        float A2 = __shfl_up_sync(0xffffffffu, A, d), B2 = __shfl_up_sync(0xffffffffu, B, d);
        if (lane >= d) { A = A + B * A2; B = B * B2; }
      }
      float carry = __shfl_up_sync(0xffffffffu, A, 1);
      if (lane == 0) carry = 0.f;
      float sumD = 0.f;
#pragma unroll
      for (int j = 0; j < C; j++) {
        dp[j] = dp[j] + carry;
        carry = dp[j] * T_(DD, j);
        sumD += dp[j];
      }
      xE = warp_sum(sumM + sumD);                                     // :402-411
      xN = xN * ploop;                                                // :413-416
      xC = (xC * ploop) + (xE * p.s.xf_E_move);
      xJ = (xJ * ploop) + (xE * p.s.xf_E_loop);
      xB = (xJ * pmove) + (xN * pmove);
      float scale = 1.0f;
      if (xE > 1.0e4f) {                                              // :420-437
        const float inv = (float) (1.0 / (double) xE);
        xN = xN / xE; xC = xC / xE; xJ = xJ / xE; xB = xB / xE;
#pragma unroll
        for (int j = 0; j < C; j++) { mp[j] *= inv; dp[j] *= inv; ip[j] *= inv; }
        scale = xE;
        totscale = (float) ((double) totscale + log((double) xE));
        xE = 1.0f;
      }
      if (xrow && lane == 0) store_specials(xrow + (size_t) i * 6, xE, xN, xJ, xB, xC, scale);
    }
    if (lane == 0) {
      int status = P7GPU_OK; float sc;
      if (isnan(xC) || (L > 0 && xC == 0.0f) || isinf(xC)) { status = P7GPU_ERANGE; sc = xC; }   // :459-461
      else sc = (float) ((double) totscale + log((double) (xC * pmove)));                         // :463
      p.c.sc[out] = sc;
      p.c.status[out] = status;
      if (p.totscale) p.totscale[out] = totscale;
    }
  }
#undef T_
}

// ------------------------------------------------------------------------------------------------
template <int C, bool SMEM, bool TREG>
__global__ void __launch_bounds__(DP_THREADS) bwd_kernel(const FbParams p)
{
  extern __shared__ __align__(16) float s_fb[];
  const float *tf, *rf;
  stage_tables<C, SMEM>(p, s_fb, tf, rf);
  const int lane = threadIdx.x & 31;
  const float *tBM = tf + 0 * C * 32 + lane, *tMM = tf + 1 * C * 32 + lane, *tIM = tf + 2 * C * 32 + lane,
              *tDM = tf + 3 * C * 32 + lane, *tMD = tf + 4 * C * 32 + lane, *tMI = tf + 5 * C * 32 + lane,
              *tII = tf + 6 * C * 32 + lane, *tDD = tf + 7 * C * 32 + lane;
  float rBM[TREG ? C : 1], rMD[TREG ? C : 1], rMI[TREG ? C : 1], rII[TREG ? C : 1], rDD[TREG ? C : 1];
  if (TREG) {
#pragma unroll
    for (int j = 0; j < C; j++) { rBM[j] = tBM[j * 32]; rMD[j] = tMD[j * 32]; rMI[j] = tMI[j * 32]; rII[j] = tII[j * 32]; rDD[j] = tDD[j * 32]; }
  }
#define T_(name, j) (TREG ? r##name[j] : t##name[(j) * 32])
  long long s, out;
  while (next_target(p.c, lane, &s, &out)) {
    const int L = p.c.len[s];
    const uint8_t *seq = p.c.res + p.c.off[s];            // seq[i-1] = x_i
    const float pmove = p.c.lc[L].pmove, ploop = 1.0f - pmove;
    const float *frow = p.fwd_xmx + p.xmx_off[out];
    float *brow = p.xmx ? p.xmx + p.xmx_off[out] : nullptr;
    float mp[C], ip[C], dp[C];
    float xJ = 0.f, xB = 0.f, xN = 0.f, xC = pmove, xE = xC * p.s.xf_E_move, totscale;   // row L, :492-500
    bool own_scales = false;

    // the transitions INTO M(k+1) seen from cell k (rotated tables: entry of node k+1), :586-588
    float nMM[C], nIM[C], nDM[C];
#pragma unroll
    for (int j = 0; j < C; j++) {
      float a = (j + 1 < C) ? tMM[(j + 1) * 32] : __shfl_down_sync(0xffffffffu, tMM[0], 1);
      float b = (j + 1 < C) ? tIM[(j + 1) * 32] : __shfl_down_sync(0xffffffffu, tIM[0], 1);
      float c = (j + 1 < C) ? tDM[(j + 1) * 32] : __shfl_down_sync(0xffffffffu, tDM[0], 1);
      if (j + 1 == C && lane == 31) { a = b = c = 0.f; }
      nMM[j] = a; nIM[j] = b; nDM[j] = c;
    }

    // ---- row L: D_k = xE + tDD_k D_{k+1};  M_k = xE + tMD_k D_{k+1};  I_k = 0  (:501-533) ----
    {
      // suffix scan: lane map from carry (D of the next lane's first cell) to this lane's first D: out = A + B*carry
      float A = 0.f, B = 1.f;
#pragma unroll
      for (int j = C - 1; j >= 0; j--) { const float t = T_(DD, j); A = xE + t * A; B = t * B; }
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        float A2 = __shfl_down_sync(0xffffffffu, A, d), B2 = __shfl_down_sync(0xffffffffu, B, d);
        if (lane + d < 32) { A = A + B * A2; B = B * B2; }
      }
      float carry = __shfl_down_sync(0xffffffffu, A, 1);
      if (lane == 31) carry = 0.f;
#pragma unroll
      for (int j = C - 1; j >= 0; j--) {
        mp[j] = xE + carry * T_(MD, j);
        dp[j] = xE + carry * T_(DD, j);
        ip[j] = 0.f;
        carry = dp[j];
      }
    }
    {
      const float scale = frow[(size_t) L * 6 + 5];
      if (scale > 1.0f) {                                             // :536-549
        const float inv = (float) (1.0 / (double) scale);
        xE = xE / scale; xN = xN / scale; xC = xC / scale; xJ = xJ / scale; xB = xB / scale;
#pragma unroll
        for (int j = 0; j < C; j++) { mp[j] *= inv; dp[j] *= inv; ip[j] *= inv; }
      }
      totscale = (float) log((double) scale);
      if (brow && lane == 0) store_specials(brow + (size_t) L * 6, xE, xN, xJ, xB, xC, scale);
    }

    // ---- rows L-1 .. 1 (:566-692) ----
    for (int i = L - 1; i >= 1; i--) {
      const int x = __ldg(seq + i);                                   // x_{i+1}
      const float *rfx = rf + (size_t) x * C * 32 + lane;
      // me[j] = M(i+1,k) * e_k(x_{i+1})
      float me[C];
#pragma unroll
      for (int j = 0; j < C; j++) me[j] = mp[j] * rfx[j * 32];
      float me_next = __shfl_down_sync(0xffffffffu, me[0], 1);        // cell k+1 of the last cell of this lane
      if (lane == 31) me_next = 0.f;
      float sumB = 0.f;
#pragma unroll
      for (int j = 0; j < C; j++) {
        const float mpv = (j + 1 < C) ? me[j + 1] : me_next;
        sumB += me[j] * T_(BM, j);
        const float ipv = ip[j];
        ip[j] = ipv * T_(II, j) + mpv * nIM[j];
        dp[j] = mpv * nDM[j];
        mp[j] = ipv * T_(MI, j) + mpv * nMM[j];
      }
      xB = warp_sum(sumB);
      xC = xC * ploop;
      xJ = (xB * pmove) + (xJ * ploop);
      xN = (xB * pmove) + (xN * ploop);
      xE = (xC * p.s.xf_E_move) + (xJ * p.s.xf_E_loop);
      // D_k = P_k + xE + tDD_k D_{k+1}: suffix scan
      {
        float A = 0.f, B = 1.f;
#pragma unroll
        for (int j = C - 1; j >= 0; j--) { const float t = T_(DD, j); A = dp[j] + (t * A + xE); B = t * B; }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          float A2 = __shfl_down_sync(0xffffffffu, A, d), B2 = __shfl_down_sync(0xffffffffu, B, d);
          if (lane + d < 32) { A = A + B * A2; B = B * B2; }
        }
        float carry = __shfl_down_sync(0xffffffffu, A, 1);
        if (lane == 31) carry = 0.f;
#pragma unroll
        for (int j = C - 1; j >= 0; j--) {
          mp[j] = (mp[j] + xE) + carry * T_(MD, j);
          dp[j] = dp[j] + (carry * T_(DD, j) + xE);
          carry = dp[j];
        }
      }
      if (xB > 1.0e16f) own_scales = true;                            // :661
      float scale;
      if (own_scales) scale = (xB > 1.0e4f) ? xB : 1.0f;
      else            scale = frow[(size_t) i * 6 + 5];
      if (scale > 1.0f) {
        const float inv = (float) (1.0 / (double) scale);
        xE /= scale; xN /= scale; xJ /= scale; xB /= scale; xC /= scale;
#pragma unroll
        for (int j = 0; j < C; j++) { mp[j] *= inv; dp[j] *= inv; ip[j] *= inv; }
        totscale = (float) ((double) totscale + log((double) scale));
      }
      if (brow && lane == 0) store_specials(brow + (size_t) i * 6, xE, xN, xJ, xB, xC, scale);
    }

    // ---- termination at i = 0 (:697-733) ----
    {
      float sumB = 0.f;
      if (L >= 1) {
        const int x = __ldg(seq);
        const float *rfx = rf + (size_t) x * C * 32 + lane;
#pragma unroll
        for (int j = 0; j < C; j++) sumB += (mp[j] * rfx[j * 32]) * T_(BM, j);
      }
      xB = warp_sum(sumB);
      xN = (xB * pmove) + (xN * ploop);
      if (brow && lane == 0) store_specials(brow, 0.f, xN, 0.f, xB, 0.f, 1.0f);
    }
    if (lane == 0) {
      int status = P7GPU_OK; float sc;
      if (isnan(xN) || (L > 0 && xN == 0.0f) || isinf(xN)) { status = P7GPU_ERANGE; sc = xN; }
      else sc = (float) ((double) totscale + log((double) xN));
      p.c.sc[out] = sc;
      p.c.status[out] = status;
      if (p.totscale) p.totscale[out] = totscale;
      if (p.own) p.own[out] = own_scales ? 1 : 0;
    }
  }
#undef T_
}

// ------------------------------------------------------------------------------------------------
template <int C, bool BWD>
static int launch_fb_c(p7gpu_ctx *ctx, cudaStream_t stream, const FbParams &p)
{
  const size_t smem = (size_t) (8 + NTABROWS) * C * 32 * sizeof(float);
  const bool use_smem = smem + 1024 <= (size_t) ctx->smem_optin;
  void (*kern)(const FbParams);
  if (BWD) kern = use_smem ? bwd_kernel<C, true, false> : bwd_kernel<C, false, false>;   // TREG costs Backward more in occupancy (190 regs) than it saves in LDS: measured 703 -> 602 GCUPS
  else     kern = use_smem ? fwd_kernel<C, true, (C <= 16)> : fwd_kernel<C, false, false>;
  const size_t dyn = use_smem ? smem : 0;
  if (use_smem) P7G_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dyn));
  int occ = 0;
  P7G_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, DP_THREADS, dyn));
  if (occ < 1) { set_error("fwd/bwd kernel<%d>: zero occupancy", C); return P7GPU_ESYS; }
  long long blocks = (long long) ctx->sm_count * occ;
  long long need = (p.c.n + (DP_THREADS / 32) - 1) / (DP_THREADS / 32);
  if (need < blocks) blocks = need;
  if (blocks < 1) blocks = 1;
  { LaunchTimer t(ctx, BWD ? P7GPU_KERNEL_BWD : P7GPU_KERNEL_FWD, stream); kern<<<(unsigned) blocks, DP_THREADS, dyn, stream>>>(p); }
  P7G_CUDA(cudaGetLastError());
  ctx->launches++;
  return P7GPU_OK;
}

static int fill_common(const Lane &ln, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
                       float *d_sc, int32_t *d_status, int counter_slot, FbParams *p)
{
  P7G_CHECK(ensure_len_table(prof, db.max_length));
  p->c.res = db.res; p->c.off = db.off; p->c.len = db.len; p->c.order = d_idx ? nullptr : db.order; p->c.idx = d_idx; p->c.first = db.first;
  p->c.n = n; p->c.lc = prof->d_len; p->c.sc = d_sc; p->c.status = d_status;
  p->c.work_counter = (unsigned long long *) (ln.counters + counter_slot);
  p->rf = prof->d_rf; p->tf = prof->d_tf; p->s = prof->fwd;
  P7G_CUDA(cudaMemsetAsync(ln.counters + counter_slot, 0, 2 * sizeof(unsigned int), ln.stream));
  return P7GPU_OK;
}

int launch_fwd(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
               float *d_sc, int32_t *d_status, float *d_xmx, const int64_t *d_xmx_off, float *d_totscale)
{
  if (n <= 0) return P7GPU_OK;
  const Lane &ln = ctx->lanes[lane];
  FbParams p;
  P7G_CHECK(fill_common(ln, prof, db, d_idx, n, d_sc, d_status, CTR_FWD_WORK, &p));
  p.xmx = d_xmx; p.fwd_xmx = nullptr; p.xmx_off = d_xmx_off; p.totscale = d_totscale; p.own = nullptr;
  if (prof->mwW > 0 && !ctx->fb_one_warp) {                           // long model: one CTA per target
    FbMwParams q; q.c = p.c; q.rf = prof->d_rf_mw; q.tf = prof->d_tf_mw; q.s = p.s;
    q.xmx = p.xmx; q.fwd_xmx = nullptr; q.xmx_off = p.xmx_off; q.totscale = p.totscale; q.own = nullptr;
    return launch_fwd_mw(ctx, ln.stream, prof, q);
  }
  switch (prof->dpC) {
#define X(CV) case CV: return launch_fb_c<CV, false>(ctx, ln.stream, p);
    P7G_FB_ONEWARP_C_CASES(X)
#undef X
  default: set_error("fwd: the one-warp cross-check kernels cover M <= 1024 (cells per lane %d)", prof->dpC); return P7GPU_EINVAL;
  }
}

int launch_bwd(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
               const float *d_fwd_xmx, const int64_t *d_xmx_off, float *d_sc, int32_t *d_status,
               float *d_bck_xmx, float *d_totscale, int32_t *d_own)
{
  if (n <= 0) return P7GPU_OK;
  const Lane &ln = ctx->lanes[lane];
  FbParams p;
  P7G_CHECK(fill_common(ln, prof, db, d_idx, n, d_sc, d_status, CTR_BWD_WORK, &p));
  p.xmx = d_bck_xmx; p.fwd_xmx = d_fwd_xmx; p.xmx_off = d_xmx_off; p.totscale = d_totscale; p.own = d_own;
  if (prof->mwW > 0 && !ctx->fb_one_warp) {
    FbMwParams q; q.c = p.c; q.rf = prof->d_rf_mw; q.tf = prof->d_tf_mw; q.s = p.s;
    q.xmx = p.xmx; q.fwd_xmx = p.fwd_xmx; q.xmx_off = p.xmx_off; q.totscale = p.totscale; q.own = d_own;
    return launch_bwd_mw(ctx, ln.stream, prof, q);
  }
  switch (prof->dpC) {
#define X(CV) case CV: return launch_fb_c<CV, true>(ctx, ln.stream, p);
    P7G_FB_ONEWARP_C_CASES(X)
#undef X
  default: set_error("bwd: the one-warp cross-check kernels cover M <= 1024 (cells per lane %d)", prof->dpC); return P7GPU_EINVAL;
  }
}

}  // namespace p7g
