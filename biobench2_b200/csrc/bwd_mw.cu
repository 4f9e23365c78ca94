// bwd_mw.cu — p7_BackwardParser for long models, one CTA of W warps per target (see fwdbwd_mw.cuh).
#include "fwdbwd_mw.cuh"

namespace p7g {

template <int C, int W>
__global__ void __launch_bounds__(32 * W) bwd_mw_kernel(const FbMwParams p)
{
  __shared__ long long s_work[2];
  __shared__ float s_me[W];          // M(i+1,k) e_k(x_{i+1}) of a warp's first node (the k+1 neighbour of the warp below)
  __shared__ float s_scan[W][2];     // D of a warp's first node as a map of what enters behind its last node
  __shared__ float s_sum[W];
  constexpr int G = 32 * W;
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
  const float *rfg = p.rf + g;
  long long s, out;
  while (mw_next_target(p.c, s_work, &s, &out)) {
    const int L = p.c.len[s];
    const uint8_t *seq = p.c.res + p.c.off[s];            // seq[i-1] = x_i
    const float pmove = p.c.lc[L].pmove, ploop = 1.0f - pmove;
    const float *frow = p.fwd_xmx + p.xmx_off[out];
    float *brow = p.xmx ? p.xmx + p.xmx_off[out] : nullptr;
    float mp[C], ip[C], dp[C], e[C];
    float xJ = 0.f, xB = 0.f, xN = 0.f, xC = pmove, xE = xC * p.s.xf_E_move, totscale;   // row L, :492-500
    bool own_scales = false;
#pragma unroll
    for (int j = 0; j < C; j++) e[j] = 0.0f;
    if (L > 0) {                                          // emissions of x_L: used by row L-1 (or by the termination if L = 1)
      const float *rfx = rfg + (size_t) __ldg(seq + L - 1) * (C * G);
#pragma unroll
      for (int j = 0; j < C; j++) e[j] = __ldg(rfx + j * G);
    }

    // ---- row L: D_k = xE + tDD_k D_{k+1};  M_k = xE + tMD_k D_{k+1};  I_k = 0  (:501-533) ----
    {
      float A = 0.f, B = 1.f;
#pragma unroll
      for (int j = C - 1; j >= 0; j--) { const float t = rDD[j]; A = xE + t * A; B = t * B; }
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float A2 = __shfl_down_sync(0xffffffffu, A, d), B2 = __shfl_down_sync(0xffffffffu, B, d);
        if (lane + d < 32) { A = A + B * A2; B = B * B2; }
      }
      if (lane == 0) { s_scan[w][0] = A; s_scan[w][1] = B; }
      __syncthreads();
      float Y = 0.f;                                                  // D of the first node of the warp above
      for (int v = W - 1; v > w; v--) Y = s_scan[v][0] + s_scan[v][1] * Y;
      float carry;
      {
        const float An = __shfl_down_sync(0xffffffffu, A, 1), Bn = __shfl_down_sync(0xffffffffu, B, 1);
        carry = (lane == 31) ? Y : An + Bn * Y;
      }
#pragma unroll
      for (int j = C - 1; j >= 0; j--) {
        mp[j] = xE + carry * rMD[j];
        dp[j] = xE + carry * rDD[j];
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

    // ---- rows L-1 .. 1 (:566-692) ----
    for (int i = L - 1; i >= 1; i--) {
      const float fscale = frow[(size_t) i * 6 + 5];
      // e[j] <- M(i+1,k) * e_k(x_{i+1})
      float sumB = 0.f;
#pragma unroll
      for (int j = 0; j < C; j++) { e[j] = mp[j] * e[j]; sumB += e[j] * rBM[j]; }
      {
        const float part = mw_warp_sum(sumB);
        if (lane == 0) { s_sum[w] = part; s_me[w] = e[0]; }
      }
      __syncthreads();                                                // ---- barrier 1 ----
      xB = s_sum[0];
      for (int v = 1; v < W; v++) xB += s_sum[v];
      float me_next = __shfl_down_sync(0xffffffffu, e[0], 1);         // cell k+1 of the last cell of this lane
      if (lane == 31) me_next = (w + 1 < W) ? s_me[w + 1] : 0.f;
#pragma unroll
      for (int j = 0; j < C; j++) {
        const float mpv = (j + 1 < C) ? e[j + 1] : me_next;
        const float ipv = ip[j];
        ip[j] = ipv * rII[j] + mpv * nIM[j];
        dp[j] = mpv * nDM[j];
        mp[j] = ipv * rMI[j] + mpv * nMM[j];
      }
      {                                                               // emissions of x_i for the next row / the termination
        const float *rfx = rfg + (size_t) __ldg(seq + i - 1) * (C * G);
#pragma unroll
        for (int j = 0; j < C; j++) e[j] = __ldg(rfx + j * G);
      }
      xC = xC * ploop;
      xJ = (xB * pmove) + (xJ * ploop);
      xN = (xB * pmove) + (xN * ploop);
      xE = (xC * p.s.xf_E_move) + (xJ * p.s.xf_E_loop);
      // D_k = P_k + xE + tDD_k D_{k+1}: suffix scan
      {
        float A = 0.f, B = 1.f;
#pragma unroll
        for (int j = C - 1; j >= 0; j--) { const float t = rDD[j]; A = dp[j] + (t * A + xE); B = t * B; }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const float A2 = __shfl_down_sync(0xffffffffu, A, d), B2 = __shfl_down_sync(0xffffffffu, B, d);
          if (lane + d < 32) { A = A + B * A2; B = B * B2; }
        }
        if (lane == 0) { s_scan[w][0] = A; s_scan[w][1] = B; }
        __syncthreads();                                              // ---- barrier 2 ----
        float Y = 0.f;
        for (int v = W - 1; v > w; v--) Y = s_scan[v][0] + s_scan[v][1] * Y;
        float carry;
        {
          const float An = __shfl_down_sync(0xffffffffu, A, 1), Bn = __shfl_down_sync(0xffffffffu, B, 1);
          carry = (lane == 31) ? Y : An + Bn * Y;
        }
#pragma unroll
        for (int j = C - 1; j >= 0; j--) {
          mp[j] = (mp[j] + xE) + carry * rMD[j];
          dp[j] = dp[j] + (carry * rDD[j] + xE);
          carry = dp[j];
        }
      }
      if (xB > 1.0e16f) own_scales = true;                            // :661
      float scale;
      if (own_scales) scale = (xB > 1.0e4f) ? xB : 1.0f;
      else            scale = fscale;
      if (scale > 1.0f) {
        const float inv = (float) (1.0 / (double) scale);
        xE /= scale; xN /= scale; xJ /= scale; xB /= scale; xC /= scale;
#pragma unroll
        for (int j = 0; j < C; j++) { mp[j] *= inv; dp[j] *= inv; ip[j] *= inv; }
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
      if (lane == 0) s_sum[w] = part;                                 // (row 1's readers of s_sum are behind its barrier 2)
      __syncthreads();
      xB = s_sum[0];
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
