// bwd_mw.cu — p7_BackwardParser for long models, one CTA of W warps per target (see fwdbwd_mw.cuh).
//
// One CTA barrier per DP row.  A Backward row (fwdback.c:566-692) has two serial parts: the B-state sum
// xB = sum_k M(i+1,k) e_k(x_{i+1}) tBM_k, from which xE follows, and the D<-D chain D_k = P_k + xE + tDD_k D_{k+1}.
// The chain is affine in xE and in what enters a lane from above, with ROW-INVARIANT coefficients:
//   D_k = Dz_k + xE * Q_k,   Q_k = 1 + tDD_k Q_{k+1}  (whole model, built once per CTA),   Dz = the chain over P alone,
//   Dz[j] = Dploc[j] + PPs[j] * Yin,   PPs[j] = prod_{j'>=j} tDD[j'],   Dploc = the lane's own chain, Yin = Dz entering it,
// so the scan over P (one Kogge-Stone pass with constant multipliers, the xB partial sums riding along) runs BEFORE xE is
// known; lane 0 publishes the warp's map and sum, and after the barrier every thread finishes the row from shared memory.
// The k+1 neighbour across a warp boundary is rebuilt by the reader from three published numbers (Ma + xE*Mq + Mk*Yz).
// Arithmetic inside a cell is that of fwdbwd.cu (FTZ; these production kernels allow FMA contraction, csrc/Makefile); only the association order of the D<-D sums differs.
#include "fwdbwd_mw.cuh"

namespace p7g {

template <int C, int W>
__global__ void __launch_bounds__(32 * W) bwd_mw_kernel(const FbMwParams p)
{
  static_assert(C >= 2, "the warp-boundary reconstruction uses the lane's second node");
  constexpr int G = 32 * W;
  __shared__ long long s_work[2];
  __shared__ float2 s_pq[C][G];         // (PPs[j], Q[j]) per thread
  __shared__ float s_kB[5][G];
  __shared__ float s_qa[G], s_qb[G], s_qin[G + 1];
  __shared__ float s_wB[W], s_wMq[W], s_wMk[W];
  __shared__ float s_pub[2][W][4];      // per row parity: P (Dz of the warp's first node, nothing entering), sumB, Ma
  __shared__ float s_first[W];          // M of a warp's first node after the row-L initialisation
  __shared__ float s_sum[W];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = threadIdx.x;
  float rBM[C], rMD[C], rMI[C], rII[C], rDD[C], nMM[C], nIM[C], nDM[C];
#pragma unroll
  for (int j = 0; j < C; j++) {
    rBM[j] = __ldg(p.tf + (0 * C + j) * G + g); rMD[j] = __ldg(p.tf + (4 * C + j) * G + g);
    rMI[j] = __ldg(p.tf + (5 * C + j) * G + g); rII[j] = __ldg(p.tf + (6 * C + j) * G + g);
    rDD[j] = __ldg(p.tf + (7 * C + j) * G + g);
    // the transitions INTO M(k+1) seen from cell k (entry of node k+1), fwdback.c:586-588
    if (j + 1 < C) {
      nMM[j] = __ldg(p.tf + (1 * C + j + 1) * G + g); nIM[j] = __ldg(p.tf + (2 * C + j + 1) * G + g);
      nDM[j] = __ldg(p.tf + (3 * C + j + 1) * G + g);
    } else if (g + 1 < G) {
      nMM[j] = __ldg(p.tf + (1 * C) * G + g + 1); nIM[j] = __ldg(p.tf + (2 * C) * G + g + 1);
      nDM[j] = __ldg(p.tf + (3 * C) * G + g + 1);
    } else nMM[j] = nIM[j] = nDM[j] = 0.0f;
  }
  // ---- row-invariant coefficients ----
  float Bexc, Qn;                       // Bexc: product of tDD over the lanes above me in my warp; Qn: Q of the next lane's first node
  {
    float pps = 1.0f, qa = 0.f;         // lane map of Q: Q_first = qa + pps * Q_in
#pragma unroll
    for (int j = C - 1; j >= 0; j--) { pps = rDD[j] * pps; s_pq[j][g].x = pps; qa = 1.0f + rDD[j] * qa; }
    s_qa[g] = qa; s_qb[g] = pps;
    float B = pps;
#pragma unroll
    for (int d = 0; d < 5; d++) {
      s_kB[d][g] = B;                                                 // this lane's segment before step d
      const float Bn = __shfl_down_sync(0xffffffffu, B, 1 << d);
      if (lane + (1 << d) < 32) B = B * Bn;
    }
    Bexc = __shfl_down_sync(0xffffffffu, B, 1);
    if (lane == 31) Bexc = 1.0f;
    if (lane == 0) s_wB[w] = B;
    __syncthreads();
    if (g == 0) {                                                     // Q over the whole model, once per CTA
      float q = 0.f;
      for (int t = G - 1; t >= 0; t--) { s_qin[t + 1] = q; q = s_qa[t] + s_qb[t] * q; }
      s_qin[0] = q;
    }
    __syncthreads();
    Qn = s_qin[g + 1];
    float q = Qn;
#pragma unroll
    for (int j = C - 1; j >= 0; j--) { q = 1.0f + rDD[j] * q; s_pq[j][g].y = q; }
    if (lane == 0) {
      s_wMq[w] = 1.0f + rMD[0] * s_pq[1][g].y;
      s_wMk[w] = (rMD[0] * s_pq[1][g].x) * Bexc;
    }
  }
  const float *rfg = p.rf + g;
  const bool has_next = (g + 1 < G);
  long long s, out;
  while (mw_next_target(p.c, s_work, &s, &out)) {         // (its barriers also order the coefficient stores above)
    const int L = p.c.len[s];
    const uint8_t *seq = p.c.res + p.c.off[s];            // seq[i-1] = x_i
    const float pmove = p.c.lc[L].pmove, ploop = 1.0f - pmove;
    const float *frow = p.fwd_xmx + p.xmx_off[out];
    float *brow = p.xmx ? p.xmx + p.xmx_off[out] : nullptr;
    float mp[C], ip[C], dp[C], e[C];
    float xJ = 0.f, xB = 0.f, xN = 0.f, xC = pmove, xE = xC * p.s.xf_E_move, totscale;   // row L, :492-500
    float nm = 0.f, en = 0.f;           // M(i+1,.) and emission of the node after this lane's last node (lane 31 uses them)
    bool own_scales = false;
#pragma unroll
    for (int j = 0; j < C; j++) e[j] = 0.0f;
    if (L > 0) {                                          // emissions of x_L: used by row L-1 (or by the termination if L = 1)
      const float *rfx = rfg + (size_t) __ldg(seq + L - 1) * (C * G);
#pragma unroll
      for (int j = 0; j < C; j++) e[j] = __ldg(rfx + j * G);
      if (has_next) en = __ldg(rfx + 1);
    }

    // ---- row L: D_k = xE + tDD_k D_{k+1} = xE Q_k;  M_k = xE + tMD_k D_{k+1};  I_k = 0  (:501-533) ----
    {
      float carry = xE * Qn;
#pragma unroll
      for (int j = C - 1; j >= 0; j--) {
        mp[j] = xE + carry * rMD[j];
        dp[j] = xE * s_pq[j][g].y;
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
      if (brow && g == 0) mw_store_specials(brow + (size_t) L * 6, xE, xN, xJ, xB, xC, scale);
    }
    if (lane == 0) s_first[w] = mp[0];
    __syncthreads();
    if (w + 1 < W) nm = s_first[w + 1];

    // ---- rows L-1 .. 1 (:566-692) ----
    // the residue whose emissions row i fetches and the Forward scale of row i-1 are read one row ahead: both are dependent
    // global loads that the row's arithmetic was found waiting for (profiles/r4e_bwd_M400_full.txt)
    uint32_t xn = L >= 2 ? __ldg(seq + L - 2) : 0u;
    float fscale_n = L >= 2 ? frow[(size_t) (L - 1) * 6 + 5] : 1.0f;
#pragma unroll (C >= 5 ? 2 : 1)
    for (int i = L - 1; i >= 1; i--) {
      const float fscale = fscale_n;
      if (i >= 2) fscale_n = frow[(size_t) (i - 1) * 6 + 5];
      // e[j] <- M(i+1,k) * e_k(x_{i+1})
      float S = 0.f;
#pragma unroll
      for (int j = 0; j < C; j++) { e[j] = mp[j] * e[j]; S += e[j] * rBM[j]; }
      float me_next = __shfl_down_sync(0xffffffffu, e[0], 1);         // cell k+1 of the last cell of this lane
      if (lane == 31) me_next = nm * en;
#pragma unroll
      for (int j = 0; j < C; j++) {
        const float mpv = (j + 1 < C) ? e[j + 1] : me_next;
        const float ipv = ip[j];
        ip[j] = ipv * rII[j] + mpv * nIM[j];
        dp[j] = mpv * nDM[j];
        mp[j] = ipv * rMI[j] + mpv * nMM[j];
      }
      {                                                               // emissions of x_i for the next row / the termination
        const float *rfx = rfg + (size_t) xn * (C * G);
#pragma unroll
        for (int j = 0; j < C; j++) e[j] = __ldg(rfx + j * G);
        if (has_next) en = __ldg(rfx + 1);
        if (i >= 2) xn = __ldg(seq + i - 2);
      }
      // the lane's own chain over P alone: dp[j] <- Dploc[j]
#pragma unroll
      for (int j = C - 2; j >= 0; j--) dp[j] = dp[j] + rDD[j] * dp[j + 1];
      float P = dp[0];
#pragma unroll
      for (int d = 0; d < 5; d++) {                                   // lane maps -> inclusive warp suffix; xB partial sums ride along
        const float Pn = __shfl_down_sync(0xffffffffu, P, 1 << d), Sn = __shfl_down_sync(0xffffffffu, S, 1 << d);
        if (lane + (1 << d) < 32) { P = P + s_kB[d][g] * Pn; S = S + Sn; }
      }
      float Yin = __shfl_down_sync(0xffffffffu, P, 1);                // Dz of the next lane's first node when nothing enters the warp
      if (lane == 31) Yin = 0.f;
      float Ymine = 0.f, Ynext = 0.f;
      if (W == 1) {                                                   // one warp: lane 0 holds the B sum, nothing enters
        xB = __shfl_sync(0xffffffffu, S, 0);
      } else {
        const int par = i & 1;
        if (lane == 0) {
          float *pub = s_pub[par][w];
          pub[0] = P; pub[1] = S;
          pub[2] = mp[0] + rMD[0] * (dp[1] + s_pq[1][g].x * Yin);     // Ma: M of my first node without its xE and Yz_w terms
        }
        __syncthreads();                                              // ---- the row's only barrier ----
        xB = 0.f;
#pragma unroll
        for (int v = 0; v < W; v++) xB += s_pub[par][v][1];
        float Yz = 0.f;                                               // Yz_v = Dz entering warp v from above
#pragma unroll
        for (int v = W - 1; v >= 0; v--) {
          if (v == w) Ymine = Yz;
          if (v == w + 1) Ynext = Yz;
          Yz = s_pub[par][v][0] + s_wB[v] * Yz;
        }
      }
      xC = xC * ploop;
      xJ = (xB * pmove) + (xJ * ploop);
      xN = (xB * pmove) + (xN * ploop);
      xE = (xC * p.s.xf_E_move) + (xJ * p.s.xf_E_loop);
      Yin = Yin + Bexc * Ymine;
      {
        float carry = Yin + xE * Qn;                                  // full D of the node behind this lane's last node
#pragma unroll
        for (int j = C - 1; j >= 0; j--) {
          const float2 pq = s_pq[j][g];
          mp[j] = (mp[j] + xE) + carry * rMD[j];
          dp[j] = (dp[j] + pq.x * Yin) + xE * pq.y;
          carry = dp[j];
        }
      }
      if (w + 1 < W) nm = (s_pub[i & 1][w + 1][2] + xE * s_wMq[w + 1]) + s_wMk[w + 1] * Ynext;
      if (xB > 1.0e16f) own_scales = true;                            // :661
      float scale;
      if (own_scales) scale = (xB > 1.0e4f) ? xB : 1.0f;
      else            scale = fscale;
      if (scale > 1.0f) {
        const float inv = (float) (1.0 / (double) scale);
        xE /= scale; xN /= scale; xJ /= scale; xB /= scale; xC /= scale;
#pragma unroll
        for (int j = 0; j < C; j++) { mp[j] *= inv; dp[j] *= inv; ip[j] *= inv; }
        nm *= inv;
        totscale = (float) ((double) totscale + log((double) scale));
      }
      if (brow && g == 0) mw_store_specials(brow + (size_t) i * 6, xE, xN, xJ, xB, xC, scale);
    }

    // ---- termination at i = 0 (:697-733) ----
    {
      float sumB = 0.f;
      if (L >= 1) {
#pragma unroll
        for (int j = 0; j < C; j++) sumB += (mp[j] * e[j]) * rBM[j];
      }
      const float part = mw_warp_sum(sumB);
      if (lane == 0) s_sum[w] = part;
      __syncthreads();
      xB = s_sum[0];
#pragma unroll
      for (int v = 1; v < W; v++) xB += s_sum[v];
      xN = (xB * pmove) + (xN * ploop);
      if (brow && g == 0) mw_store_specials(brow, 0.f, xN, 0.f, xB, 0.f, 1.0f);
    }
    if (g == 0) {
      int status = P7GPU_OK; float sc;
      if (isnan(xN) || (L > 0 && xN == 0.0f) || isinf(xN)) { status = P7GPU_ERANGE; sc = xN; }
      else sc = (float) ((double) totscale + log((double) xN));
      p.c.sc[out] = sc;
      p.c.status[out] = status;
      if (p.totscale) p.totscale[out] = totscale;
      if (p.own) p.own[out] = own_scales ? 1 : 0;
    }
  }
}

template <int C, int W>
static int launch_bwd_mw_c(p7gpu_ctx *ctx, cudaStream_t stream, const FbMwParams &p)
{
  auto kern = bwd_mw_kernel<C, W>;
  int occ = 0;
  P7G_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * W, 0));
  if (occ < 1) { set_error("bwd_mw_kernel<%d,%d>: zero occupancy", C, W); return P7GPU_ESYS; }
  long long blocks = (long long) ctx->sm_count * occ;
  if (p.c.n < blocks) blocks = p.c.n;
  if (blocks < 1) blocks = 1;
  { LaunchTimer t(ctx, P7GPU_KERNEL_BWD, stream); kern<<<(unsigned) blocks, 32 * W, 0, stream>>>(p); }
  P7G_CUDA(cudaGetLastError());
  ctx->launches++;
  return P7GPU_OK;
}

int launch_bwd_mw(p7gpu_ctx *ctx, cudaStream_t stream, const p7gpu_profile *prof, const FbMwParams &p)
{
  const int C = prof->mwC, W = prof->mwW;
  switch (C * 100 + W) {
#define X(CV, WV) case CV * 100 + WV: return launch_bwd_mw_c<CV, WV>(ctx, stream, p);
    P7G_MW_CASES(X)
#undef X
  default: set_error("bwd_mw: no kernel for C=%d W=%d", C, W); return P7GPU_EINVAL;
  }
}

}  // namespace p7g
