// fwd_mw.cu — p7_ForwardParser for long models, one CTA of W warps per target (see fwdbwd_mw.cuh).
#include "fwdbwd_mw.cuh"

namespace p7g {

template <int C, int WMAX>
__global__ void __launch_bounds__(32 * WMAX) fwd_mw_kernel(const FbMwParams p)
{
  __shared__ long long s_work[2];
  __shared__ float s_bnd[WMAX][4];      // M, I, D of a warp's last node (row i-1, before that row's rescale)
  __shared__ float s_scan[WMAX][2];     // what leaves a warp towards the next one: X_{w+1} = s[0] + s[1] * X_w
  __shared__ float s_sum[WMAX];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int G = p.G, g = threadIdx.x;
  float rBM[C], rMM[C], rIM[C], rDM[C], rMD[C], rMI[C], rII[C], rDD[C];
#pragma unroll
  for (int j = 0; j < C; j++) {
    rBM[j] = __ldg(p.tf + (0 * C + j) * G + g); rMM[j] = __ldg(p.tf + (1 * C + j) * G + g);
    rIM[j] = __ldg(p.tf + (2 * C + j) * G + g); rDM[j] = __ldg(p.tf + (3 * C + j) * G + g);
    rMD[j] = __ldg(p.tf + (4 * C + j) * G + g); rMI[j] = __ldg(p.tf + (5 * C + j) * G + g);
    rII[j] = __ldg(p.tf + (6 * C + j) * G + g); rDD[j] = __ldg(p.tf + (7 * C + j) * G + g);
  }
  const float *rfg = p.rf + g;
  long long s, out;
  while (mw_next_target(p.c, s_work, &s, &out)) {
    const int L = p.c.len[s];
    const uint8_t *seq = p.c.res + p.c.off[s];
    const float pmove = p.c.lc[L].pmove, ploop = 1.0f - pmove;
    float *xrow = p.xmx ? p.xmx + p.xmx_off[out] : nullptr;
    float mp[C], ip[C], dp[C], e[C];
#pragma unroll
    for (int j = 0; j < C; j++) mp[j] = ip[j] = dp[j] = e[j] = 0.0f;
    float xE = 0.f, xN = 1.f, xJ = 0.f, xB = pmove, xC = 0.f, totscale = 0.f;     // fwdback.c:287-294
    float rs = 1.0f;                                                              // 1/scale of the previous row
    if (xrow && g == 0) mw_store_specials(xrow, xE, xN, xJ, xB, xC, 1.0f);
    if (L > 0) {
      const float *rfx = rfg + (size_t) __ldg(seq) * C * G;
#pragma unroll
      for (int j = 0; j < C; j++) e[j] = __ldg(rfx + j * G);
    }

    for (int i = 1; i <= L; i++) {
      float pm = __shfl_up_sync(0xffffffffu, mp[C - 1], 1);
      float pi = __shfl_up_sync(0xffffffffu, ip[C - 1], 1);
      float pd = __shfl_up_sync(0xffffffffu, dp[C - 1], 1);
      if (lane == 0) {
        if (w == 0 || i == 1) pm = pi = pd = 0.0f;
        else { pm = s_bnd[w - 1][0] * rs; pi = s_bnd[w - 1][1] * rs; pd = s_bnd[w - 1][2] * rs; }
      }
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
        dp[j] = dcv;                                                  // M(i,k-1) * tMD(k-1)
        dcv = sv * rMD[j];
        ip[j] = pm * rMI[j] + pi * rII[j];
      }
      if (i < L) {                                                    // next row's emissions: in flight during the scans
        const float *rfx = rfg + (size_t) __ldg(seq + i) * C * G;
#pragma unroll
        for (int j = 0; j < C; j++) e[j] = __ldg(rfx + j * G);
      }
      {
        const float d0 = __shfl_up_sync(0xffffffffu, dcv, 1);
        dp[0] = (lane == 0) ? 0.0f : d0;                              // a warp's first node gets its M->D term with X below
      }
      // D->D paths (:351-399): D_k = P_k + D_{k-1} * tDD_{k-1}.  Lane map: out = A + B * carry-in
      float A = 0.f, B = 1.f;
#pragma unroll
      for (int j = 0; j < C; j++) { const float t = rDD[j]; A = (dp[j] + A) * t; B = B * t; }
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float A2 = __shfl_up_sync(0xffffffffu, A, d), B2 = __shfl_up_sync(0xffffffffu, B, d);
        if (lane >= d) { A = A + B * A2; B = B * B2; }
      }
      if (lane == 31) { s_scan[w][0] = dcv + A; s_scan[w][1] = B; }
      __syncthreads();                                                // ---- barrier 1: warp maps published ----
      float X = 0.f;                                                  // what enters this warp's first node
      for (int v = 0; v < w; v++) X = s_scan[v][0] + s_scan[v][1] * X;
      float carry;
      {
        const float Ap = __shfl_up_sync(0xffffffffu, A, 1), Bp = __shfl_up_sync(0xffffffffu, B, 1);
        carry = (lane == 0) ? X : Ap + Bp * X;
      }
      float sumD = 0.f;
#pragma unroll
      for (int j = 0; j < C; j++) {
        dp[j] = dp[j] + carry;
        carry = dp[j] * rDD[j];
        sumD += dp[j];
      }
      if (lane == 31) { s_bnd[w][0] = mp[C - 1]; s_bnd[w][1] = ip[C - 1]; s_bnd[w][2] = dp[C - 1]; }
      {
        const float part = mw_warp_sum(sumM + sumD);
        if (lane == 0) s_sum[w] = part;
      }
      __syncthreads();                                                // ---- barrier 2: row sums published ----
      xE = s_sum[0];                                                  // :402-411
      for (int v = 1; v < W; v++) xE += s_sum[v];
      xN = xN * ploop;                                                // :413-416
      xC = (xC * ploop) + (xE * p.s.xf_E_move);
      xJ = (xJ * ploop) + (xE * p.s.xf_E_loop);
      xB = (xJ * pmove) + (xN * pmove);
      float scale = 1.0f;
      rs = 1.0f;
      if (xE > 1.0e4f) {                                              // :420-437
        const float inv = (float) (1.0 / (double) xE);
        xN = xN / xE; xC = xC / xE; xJ = xJ / xE; xB = xB / xE;
#pragma unroll
        for (int j = 0; j < C; j++) { mp[j] *= inv; dp[j] *= inv; ip[j] *= inv; }
        scale = xE; rs = inv;
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

template <int C, int WMAX>
static int launch_fwd_mw_c(p7gpu_ctx *ctx, cudaStream_t stream, const FbMwParams &p)
{
  auto kern = fwd_mw_kernel<C, WMAX>;
  const int threads = p.G;
  int occ = 0;
  P7G_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0));
  if (occ < 1) { set_error("fwd_mw_kernel<%d,%d>: zero occupancy at %d threads", C, WMAX, threads); return P7GPU_ESYS; }
  long long blocks = (long long) ctx->sm_count * occ;
  if (p.c.n < blocks) blocks = p.c.n;
  if (blocks < 1) blocks = 1;
  { LaunchTimer t(ctx, P7GPU_KERNEL_FWD, stream); kern<<<(unsigned) blocks, threads, 0, stream>>>(p); }
  P7G_CUDA(cudaGetLastError());
  ctx->launches++;
  return P7GPU_OK;
}

int launch_fwd_mw(p7gpu_ctx *ctx, cudaStream_t stream, const p7gpu_profile *prof, const FbMwParams &p)
{
  const int C = prof->mwC, W = prof->mwW, key = C * 100 + prof->mwWmax;
  switch (key) {
#define X(CV, WM) case CV * 100 + WM: return launch_fwd_mw_c<CV, WM>(ctx, stream, p);
    P7G_MW_CASES(X)
#undef X
  default: set_error("fwd_mw: no kernel for C=%d W=%d", C, W); return P7GPU_EINVAL;
  }
}

}  // namespace p7g
