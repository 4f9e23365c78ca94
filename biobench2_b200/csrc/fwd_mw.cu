// fwd_mw.cu — p7_ForwardParser for long models, one CTA of W warps per target (see fwdbwd_mw.cuh).
//
// One CTA barrier per DP row.  The serial part of a Forward row is the D->D chain D_k = P_k + tDD_{k-1} D_{k-1}
// (fwdback.c:351-399) followed by the row sum xE = sum_k M_k + D_k (:402-411) that feeds xB of the next row.  Both are
// affine in what enters a lane (Z) and all coefficients are products of tDD, i.e. ROW-INVARIANT, so:
//   lane      D[j] = D0[j] + Z * PP[j]            PP[j] = prod_{j'<j} tDD[j'],  D0 = the chain with nothing entering
//             leaves the lane   P + Bl * Z        P = M_last tMD_last + D0_last tDD_last,  Bl = prod_j tDD[j]
//             row-sum part      U + SPP * Z       U = sum_j M[j] + D0[j],  SPP = sum_j PP[j]
//   warp      ONE Kogge-Stone pass over (P, U) composes the lane maps (constants kB/kV per step, built once per CTA); one-warp
//             CTAs scan P alone and take the row sum as a plain warp sum of U + cU P with row-invariant weights cU
//   CTA       lane 31 publishes the warp's (P, U) and its last node's M, I, D-part; after the barrier every thread
//             runs the <= W-step recurrence X_{v+1} = P_v + B_v X_v and xE = sum_v U_v + V_v X_v from shared memory.
// The row is PIPELINED: the state carried from row to row is part[j] = M tMM + I tIM + D0 tDM into node j (everything of M(i+1,j)
// that does not wait for row i's scan) and I(i+1,j), both computed under the scan's shuffle latency; after the scan a cell is
// finished as ((part + xB tBM) + Z q[j]) e with q[j] = PP[j-1] tDM[j] (row-invariant, shared memory) — the correction pass
// D += Z PP is gone, and the first cell of a lane takes the chain of the lane below (Zp, q[0]).  DESIGN.md 4.4.
// Arithmetic: FTZ; every fused multiply-add is written out (fmaf) and the file is built without implicit contraction, so that
// variants can be compared bit for bit (probes/fwd_tm_experiment.cu); only the association order of the sums differs from fwdbwd.cu.
#include "fwdbwd_mw.cuh"

namespace p7g {

template <int C, int W>
__global__ void __launch_bounds__(32 * W) fwd_mw_kernel(const FbMwParams p)
{
  constexpr int G = 32 * W;
  __shared__ long long s_work[2];
  __shared__ float s_q[C][G];           // q[j]: what one unit entering the D chain of the lane that owns node k-1 adds to M(i+1,k)
  __shared__ float s_ppl[G];            // PP[C-1] per thread
  __shared__ float s_kB[5][G], s_kV[W == 1 ? 1 : 5][G];
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
  float Bexc, BexcPrev, ppLast, cU;
  {
    float pp = 1.0f, spp = 0.f;
#pragma unroll
    for (int j = 0; j < C; j++) {
      if (j + 1 < C) s_q[j + 1][g] = pp * rDM[j + 1];                 // D(i,k-1) = D0 + Z PP[j] feeds M(i+1,k) through tDM
      else { s_ppl[g] = pp; ppLast = pp; }
      spp += pp; pp = pp * rDD[j];
    }
    float B = pp, V = spp;
#pragma unroll
    for (int d = 0; d < 5; d++) {
      s_kB[d][g] = B; if (W > 1) s_kV[d][g] = V;                      // this lane's segment before step d
      const float Bl = __shfl_up_sync(0xffffffffu, B, 1 << d), Vl = __shfl_up_sync(0xffffffffu, V, 1 << d);
      if (lane >= (1 << d)) { V = Vl + V * Bl; B = B * Bl; }
    }
    // One-warp CTAs: the row sum is linear in the lanes' P with ROW-INVARIANT weights, xE = sum_g U_g + SPP_g Z_g with
    // Z_g = sum_{h<g} T(g,h) P_h, i.e. xE = sum_g (U_g + cU_g P_g), cU_g = SPP_{g+1} + Bl_{g+1} cU_{g+1} — one plain warp sum
    // instead of a second value riding through the Kogge-Stone pass (10 instructions fewer per row: M=100 0.61 -> 0.66 TCUPS,
    // M=400 1.15 -> 1.16; multi-warp CTAs measured slower with it, 1.085 -> 1.05 at M=2000, and keep the pair scan)
    cU = 0.f;
    if (W == 1) {
      float c = 0.f;
      const float Bl = pp, Sl = spp;                                  // this lane's whole-lane product and sum
#pragma unroll 1
      for (int t = 30; t >= 0; t--) {
        const float cn = __shfl_down_sync(0xffffffffu, c, 1), Bn = __shfl_down_sync(0xffffffffu, Bl, 1), Sn = __shfl_down_sync(0xffffffffu, Sl, 1);
        if (lane == t) c = fmaf(Bn, cn, Sn);
      }
      cU = c;
    }
    Bexc = __shfl_up_sync(0xffffffffu, B, 1);
    if (lane == 0) Bexc = 1.0f;
    BexcPrev = __shfl_up_sync(0xffffffffu, Bexc, 1);                  // of the lane below (unused by lane 0)
    if (lane == 31) { s_wB[w] = B; s_wV[w] = V; s_wK[w] = Bexc * ppLast; }
    __syncthreads();
    s_q[0][g] = lane ? s_ppl[g - 1] * rDM[0] : 0.0f;                  // first node: the D chain of the lane below (lane 0: see bd)
  }
  const float *rfg = p.rf + g;
  long long s, out;
  while (mw_next_target(p.c, s_work, &s, &out)) {         // (its barriers also order the coefficient stores above)
    const int L = p.c.len[s];
    const uint8_t *seq = p.c.res + p.c.off[s];
    const float pmove = p.c.lc[L].pmove, ploop = 1.0f - pmove;
    float *xrow = p.xmx ? p.xmx + p.xmx_off[out] : nullptr;
    // Row state, one row ahead of the recursion: part[j] = M tMM + I tIM + D0 tDM of row i-1 into node j (everything of
    // M(i,j) that does not wait for row i-1's scan), ip[j] = I(i,j), and the two scan results of row i-1: xB and what
    // entered the lane's D chain (Z; Zp for the lane below).
    float part[C], ip[C], e[C];
#pragma unroll
    for (int j = 0; j < C; j++) part[j] = ip[j] = e[j] = 0.0f;
    float xE = 0.f, xN = 1.f, xJ = 0.f, xB = pmove, xC = 0.f, totscale = 0.f;     // fwdback.c:287-294
    float Z = 0.f, Zp = 0.f;
    float bx = 0.f;                                       // lane 0 of warps > 0: M, I, D of the node below, times their transitions
    if (xrow && g == 0) mw_store_specials(xrow, xE, xN, xJ, xB, xC, 1.0f);
    // multi-warp CTAs fetch the residue of row i+1 a row before its emissions (the dependent byte load was found on the row's
    // critical path there: 1.03 -> 1.09 TCUPS at M=2000); one-warp CTAs are faster without the extra live register
    constexpr bool PF = W > 1;
    uint32_t xn = 0;
    if (L > 0) {
      const float *rfx = rfg + (size_t) __ldg(seq) * (C * G);
#pragma unroll
      for (int j = 0; j < C; j++) e[j] = __ldg(rfx + j * G);
      if (PF && L > 1) xn = __ldg(seq + 1);
    }

#pragma unroll 1
    for (int i = 1; i <= L; i++) {
      // ---- behind the scan of row i-1: finish M(i,.) (:318-345) ----
      float sv[C];
#pragma unroll
      for (int j = C - 1; j >= 0; j--) {
        float v = fmaf(xB, rBM[j], part[j]);
        v = fmaf(j ? Z : Zp, s_q[j][g], v);
        if (W > 1 && j == 0) v = v + bx;
        sv[j] = v * e[j];
      }
      const float pm_b = __shfl_up_sync(0xffffffffu, sv[C - 1], 1);
      const float pi_b = __shfl_up_sync(0xffffffffu, ip[C - 1], 1);
      const float ipLast = ip[C - 1];
      // the lane's own D->D chain with nothing entering (D0), the value leaving the lane (P) and the row-sum part (U)
      float d0[C];
      float sumM = sv[0], sumD0 = 0.f;
      d0[0] = 0.f;
#pragma unroll
      for (int j = 1; j < C; j++) {
        d0[j] = fmaf(d0[j - 1], rDD[j - 1], sv[j - 1] * rMD[j - 1]);   // one FMA per link of the chain
        sumM += sv[j]; sumD0 += d0[j];
      }
      float P = fmaf(d0[C - 1], rDD[C - 1], sv[C - 1] * rMD[C - 1]);
      float U = sumM + sumD0;
      const float pd_b = __shfl_up_sync(0xffffffffu, d0[C - 1], 1);
      // ---- independent of the scan: row i+1's part and I (these fill the scan's shuffle latency) ----
#pragma unroll
      for (int j = C - 1; j >= 1; j--) {
        part[j] = fmaf(d0[j - 1], rDM[j], fmaf(ip[j - 1], rIM[j], sv[j - 1] * rMM[j]));
        ip[j] = fmaf(ip[j], rII[j], sv[j] * rMI[j]);
      }
      {
        const float v = fmaf(pd_b, rDM[0], fmaf(pi_b, rIM[0], pm_b * rMM[0]));
        part[0] = lane ? v : 0.f;
        ip[0] = fmaf(ip[0], rII[0], sv[0] * rMI[0]);
      }
      if (i < L) {                                                    // next row's emissions
        const float *rfx = rfg + (size_t) (PF ? xn : (uint32_t) __ldg(seq + i)) * (C * G);
#pragma unroll
        for (int j = 0; j < C; j++) e[j] = __ldg(rfx + j * G);
        if (PF && i + 1 < L) xn = __ldg(seq + i + 1);
      }
      if (W == 1) {
        U = fmaf(cU, P, U);                                           // this lane's share of the row sum
#pragma unroll
        for (int d = 0; d < 5; d++) {                                 // lane maps -> inclusive warp prefix; the row sum beside it
          const float Pl = __shfl_up_sync(0xffffffffu, P, 1 << d);
          U += __shfl_xor_sync(0xffffffffu, U, 1 << d);
          if (lane >= (1 << d)) P = fmaf(s_kB[d][g], Pl, P);
        }
      } else {
#pragma unroll
        for (int d = 0; d < 5; d++) {                                 // lane maps -> inclusive warp prefix, the row sum riding along
          const float Pl = __shfl_up_sync(0xffffffffu, P, 1 << d), Ul = __shfl_up_sync(0xffffffffu, U, 1 << d);
          if (lane >= (1 << d)) { U = fmaf(s_kV[d][g], Pl, Ul + U); P = fmaf(s_kB[d][g], Pl, P); }
        }
      }
      Z = __shfl_up_sync(0xffffffffu, P, 1);                          // enters this lane when nothing enters the warp
      Zp = __shfl_up_sync(0xffffffffu, P, 2);                         // ... and the lane below
      if (lane == 0) Z = 0.f;
      if (lane <= 1) Zp = 0.f;
      if (W == 1) {                                                   // one warp: every lane holds the row sum, nothing enters
        xE = U;
      } else {
        const int par = i & 1;
        if (lane == 31) {
          float *pub = s_pub[par][w];
          pub[0] = P; pub[1] = U; pub[2] = sv[C - 1]; pub[3] = ipLast; pub[4] = fmaf(Z, ppLast, d0[C - 1]);
        }
        __syncthreads();                                              // ---- the row's only barrier ----
        float X = 0.f, Xmine = 0.f, Xprev = 0.f;
        xE = 0.f;
#pragma unroll
        for (int v = 0; v < W; v++) {                                 // X_v = what enters warp v's first node
          if (v == w) Xmine = X;
          if (v + 1 == w) Xprev = X;
          xE += fmaf(s_wV[v], X, s_pub[par][v][1]);
          X = fmaf(s_wB[v], X, s_pub[par][v][0]);
        }
        Z = fmaf(Bexc, Xmine, Z);
        if (lane) Zp = fmaf(BexcPrev, Xmine, Zp);
        bx = 0.f;
        if (w > 0 && lane == 0) {
          const float *pub = s_pub[par][w - 1];
          bx = fmaf(fmaf(s_wK[w - 1], Xprev, pub[4]), rDM[0], fmaf(pub[3], rIM[0], pub[2] * rMM[0]));
        }
      }
      xN = xN * ploop;                                                // :413-416
      xC = fmaf(xC, ploop, xE * p.s.xf_E_move);
      xJ = fmaf(xJ, ploop, xE * p.s.xf_E_loop);
      xB = fmaf(xJ, pmove, xN * pmove);
      float scale = 1.0f;
      if (xE > 1.0e4f) {                                              // :420-437 (everything carried forward is linear in row i)
        const float inv = (float) (1.0 / (double) xE);
        xN = xN / xE; xC = xC / xE; xJ = xJ / xE; xB = xB / xE;
#pragma unroll
        for (int j = 0; j < C; j++) { part[j] *= inv; ip[j] *= inv; }
        Z *= inv; Zp *= inv; bx *= inv;
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
