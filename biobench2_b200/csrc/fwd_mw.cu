// fwd_mw.cu — p7_ForwardParser for long models, one CTA of W warps per target (see fwdbwd_mw.cuh).
//
// One CTA barrier per DP row.  The serial part of a Forward row is the D->D chain D_k = P_k + tDD_{k-1} D_{k-1}
// (fwdback.c:351-399) followed by the row sum xE = sum_k M_k + D_k (:402-411) that feeds xB of the next row.  Both are
// affine in what enters a lane (Z) and all coefficients are products of tDD, i.e. ROW-INVARIANT, so:
//   lane      D[j] = D0[j] + Z * PP[j]            PP[j] = prod_{j'<j} tDD[j'],  D0 = the chain with nothing entering
//             leaves the lane   P + Bl * Z        P = M_last tMD_last + D0_last tDD_last,  Bl = prod_j tDD[j]
//             row-sum part      U + SPP * Z       U = sum_j M[j] + D0[j],  SPP = sum_j PP[j]
//   warp      ONE Kogge-Stone pass over (P, U) composes the lane maps (constants kB/kV per step, built once per CTA)
//   CTA       lane 31 publishes the warp's (P, U) and its last node's M, I, D-part; after the barrier every thread
//             runs the <= W-step recurrence X_{v+1} = P_v + B_v X_v and xE = sum_v U_v + V_v X_v from shared memory.
// Arithmetic inside a cell is that of fwdbwd.cu (FTZ; these production kernels allow FMA contraction, csrc/Makefile); only the association order of the D->D sums differs.
#include "fwdbwd_mw.cuh"

namespace p7g {

template <int C, int W>
__global__ void __launch_bounds__(32 * W) fwd_mw_kernel(const FbMwParams p)
{
  constexpr int G = 32 * W;
  __shared__ long long s_work[2];
  __shared__ float s_pp[C][G];          // PP[j] per thread
  __shared__ float s_kB[5][G], s_kV[5][G];
  __shared__ float s_wB[W], s_wV[W], s_wK[W];      // per warp: B, V of the whole warp; K = Bexc(lane 31) * PP[C-1]
  __shared__ float s_pub[2][W][6];      // per row parity: P, U, M_last, I_last, D_last part (X = 0), pad
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = threadIdx.x;
  float rBM[C], rMM[C], rIM[C], rDM[C], rMD[C], rMI[C], rII[C], rDD[C];
#pragma unroll
  for (int j = 0; j < C; j++) {
    rBM[j] = __ldg(p.tf + (0 * C + j) * G + g); rMM[j] = __ldg(p.tf + (1 * C + j) * G + g);
    rIM[j] = __ldg(p.tf + (2 * C + j) * G + g); rDM[j] = __ldg(p.tf + (3 * C + j) * G + g);
    rMD[j] = __ldg(p.tf + (4 * C + j) * G + g); rMI[j] = __ldg(p.tf + (5 * C + j) * G + g);
    rII[j] = __ldg(p.tf + (6 * C + j) * G + g); rDD[j] = __ldg(p.tf + (7 * C + j) * G + g);
  }
  // ---- row-invariant coefficients of the D->D maps ----
  float SPP, Bexc;
  {
    float pp = 1.0f, spp = 0.f;
#pragma unroll
    for (int j = 0; j < C; j++) { s_pp[j][g] = pp; spp += pp; pp = pp * rDD[j]; }
    float B = pp, V = spp;
    SPP = spp;
#pragma unroll
    for (int d = 0; d < 5; d++) {
      s_kB[d][g] = B; s_kV[d][g] = V;                                 // this lane's segment before step d
      const float Bl = __shfl_up_sync(0xffffffffu, B, 1 << d), Vl = __shfl_up_sync(0xffffffffu, V, 1 << d);
      if (lane >= (1 << d)) { V = Vl + V * Bl; B = B * Bl; }
    }
    Bexc = __shfl_up_sync(0xffffffffu, B, 1);
    if (lane == 0) Bexc = 1.0f;
    if (lane == 31) { s_wB[w] = B; s_wV[w] = V; s_wK[w] = Bexc * s_pp[C - 1][g]; }
  }
  const float *rfg = p.rf + g;
  long long s, out;
  while (mw_next_target(p.c, s_work, &s, &out)) {         // (its barriers also order the coefficient stores above)
    const int L = p.c.len[s];
    const uint8_t *seq = p.c.res + p.c.off[s];
    const float pmove = p.c.lc[L].pmove, ploop = 1.0f - pmove;
    float *xrow = p.xmx ? p.xmx + p.xmx_off[out] : nullptr;
    float mp[C], ip[C], dp[C], e[C];
#pragma unroll
    for (int j = 0; j < C; j++) mp[j] = ip[j] = dp[j] = e[j] = 0.0f;
    float xE = 0.f, xN = 1.f, xJ = 0.f, xB = pmove, xC = 0.f, totscale = 0.f;     // fwdback.c:287-294
    float bm = 0.f, bi = 0.f, bd = 0.f;                   // M, I, D of node k-1 of this lane's first node (row i-1)
    if (xrow && g == 0) mw_store_specials(xrow, xE, xN, xJ, xB, xC, 1.0f);
    if (L > 0) {
      const float *rfx = rfg + (size_t) __ldg(seq) * (C * G);
#pragma unroll
      for (int j = 0; j < C; j++) e[j] = __ldg(rfx + j * G);
    }

#pragma unroll (C >= 5 ? 2 : 1)
    for (int i = 1; i <= L; i++) {
      float pm = __shfl_up_sync(0xffffffffu, mp[C - 1], 1);
      float pi = __shfl_up_sync(0xffffffffu, ip[C - 1], 1);
      float pd = __shfl_up_sync(0xffffffffu, dp[C - 1], 1);
      if (lane == 0) { pm = bm; pi = bi; pd = bd; }
      float sumM = 0.f, dcv = 0.f;
#pragma unroll
      for (int j = 0; j < C; j++) {                                   // :318-345
        float sv = xB * rBM[j];
        sv = sv + pm * rMM[j];
        sv = sv + pi * rIM[j];
        sv = sv + pd * rDM[j];
        sv = sv * e[j];
        sumM += sv;
        pm = mp[j]; pi = ip[j]; pd = dp[j];
        mp[j] = sv;
        dp[j] = dcv;                                                  // M(i,k-1) * tMD(k-1); 0 for the lane's first node
        dcv = sv * rMD[j];
        ip[j] = pm * rMI[j] + pi * rII[j];
      }
      if (i < L) {                                                    // next row's emissions: in flight during the scan
        const float *rfx = rfg + (size_t) __ldg(seq + i) * (C * G);
#pragma unroll
        for (int j = 0; j < C; j++) e[j] = __ldg(rfx + j * G);
      }
      // the lane's own D->D chain (nothing entering): dp[j] <- D0[j]
      float sumD0 = 0.f;
#pragma unroll
      for (int j = 1; j < C; j++) { dp[j] = dp[j] + dp[j - 1] * rDD[j - 1]; sumD0 += dp[j]; }
      float P = dcv + dp[C - 1] * rDD[C - 1];
      float U = sumM + sumD0;
#pragma unroll
      for (int d = 0; d < 5; d++) {                                   // lane maps -> inclusive warp prefix
        const float Pl = __shfl_up_sync(0xffffffffu, P, 1 << d), Ul = __shfl_up_sync(0xffffffffu, U, 1 << d);
        if (lane >= (1 << d)) { U = (Ul + U) + s_kV[d][g] * Pl; P = P + s_kB[d][g] * Pl; }
      }
      float Z = __shfl_up_sync(0xffffffffu, P, 1);                    // enters this lane when nothing enters the warp
      if (lane == 0) Z = 0.f;
      if (W == 1) {                                                   // one warp: lane 31 holds the row sum, nothing enters
        xE = __shfl_sync(0xffffffffu, U, 31);
#pragma unroll
        for (int j = 0; j < C; j++) dp[j] = dp[j] + Z * s_pp[j][g];
      } else {
        const int par = i & 1;
        if (lane == 31) {
          float *pub = s_pub[par][w];
          pub[0] = P; pub[1] = U; pub[2] = mp[C - 1]; pub[3] = ip[C - 1]; pub[4] = dp[C - 1] + Z * s_pp[C - 1][g];
        }
        __syncthreads();                                              // ---- the row's only barrier ----
        float X = 0.f, Xmine = 0.f, Xprev = 0.f;
        xE = 0.f;
#pragma unroll
        for (int v = 0; v < W; v++) {                                 // X_v = what enters warp v's first node
          if (v == w) Xmine = X;
          if (v + 1 == w) Xprev = X;
          xE += s_pub[par][v][1] + s_wV[v] * X;
          X = s_pub[par][v][0] + s_wB[v] * X;
        }
        Z = Z + Bexc * Xmine;
#pragma unroll
        for (int j = 0; j < C; j++) dp[j] = dp[j] + Z * s_pp[j][g];
        if (w > 0) { const float *pub = s_pub[par][w - 1]; bm = pub[2]; bi = pub[3]; bd = pub[4] + s_wK[w - 1] * Xprev; }
      }
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
        bm *= inv; bi *= inv; bd *= inv;
        scale = xE;
        totscale = (float) ((double) totscale + log((double) xE));
        xE = 1.0f;
      }
      if (xrow && g == 0) mw_store_specials(xrow + (size_t) i * 6, xE, xN, xJ, xB, xC, scale);
    }
    if (g == 0) {
      int status = P7GPU_OK; float sc;
      if (isnan(xC) || (L > 0 && xC == 0.0f) || isinf(xC)) { status = P7GPU_ERANGE; sc = xC; }   // :459-461
      else sc = (float) ((double) totscale + log((double) (xC * pmove)));                         // :463
      p.c.sc[out] = sc;
      p.c.status[out] = status;
      if (p.totscale) p.totscale[out] = totscale;
    }
  }
}

template <int C, int W>
static int launch_fwd_mw_c(p7gpu_ctx *ctx, cudaStream_t stream, const FbMwParams &p)
{
  auto kern = fwd_mw_kernel<C, W>;
  int occ = 0;
  P7G_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * W, 0));
  if (occ < 1) { set_error("fwd_mw_kernel<%d,%d>: zero occupancy", C, W); return P7GPU_ESYS; }
  long long blocks = (long long) ctx->sm_count * occ;
  if (p.c.n < blocks) blocks = p.c.n;
  if (blocks < 1) blocks = 1;
  { LaunchTimer t(ctx, P7GPU_KERNEL_FWD, stream); kern<<<(unsigned) blocks, 32 * W, 0, stream>>>(p); }
  P7G_CUDA(cudaGetLastError());
  ctx->launches++;
  return P7GPU_OK;
}

int launch_fwd_mw(p7gpu_ctx *ctx, cudaStream_t stream, const p7gpu_profile *prof, const FbMwParams &p)
{
  const int C = prof->mwC, W = prof->mwW;
  switch (C * 100 + W) {
#define X(CV, WV) case CV * 100 + WV: return launch_fwd_mw_c<CV, WV>(ctx, stream, p);
    P7G_MW_CASES(X)
#undef X
  default: set_error("fwd_mw: no kernel for C=%d W=%d", C, W); return P7GPU_EINVAL;
  }
}

}  // namespace p7g
