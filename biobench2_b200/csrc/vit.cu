// vit.cu — sm_100a kernel for p7_ViterbiFilter (HMMER/src/impl_sse/vitfilter.c:83-249).
//
// One warp per target, lane l owns nodes k = l*C+1 .. l*C+C (dp_common.cuh).  The 16-bit saturating
// arithmetic of the reference is reproduced exactly in 32-bit lanes with DPX VIADDMNMX (max(a+b,c)):
// every adds_epi16(a,t) with t <= 0 equals max(a+t, -32768), and the only add that can saturate upward
// (the emission) makes xE >= 32767, which returns eslERANGE for the row before the value is used again.
// The lazy-F test of vitfilter.c:198-199 is evaluated as written; when it fires, the D->D fixed point
// of :203-226 is obtained with a max-plus prefix scan over the warp (exact: see DESIGN.md).
#include "dp_common.cuh"
#include <math_constants.h>

namespace p7g {

struct VitParams {
  DpCommon c;
  const int16_t *rw;    // [NTABROWS][C][32]
  const int16_t *tw;    // [8][C][32]
  VitScalars s;
};

constexpr int VNEG = -32768;
constexpr int VNOCARRY = -(1 << 28);

// TREG: the 7 per-node transition vectors BM,MM,IM,DM,MD,MI,II live in registers (7*C per lane) instead of being
// re-read from shared memory on every row: the row loop is then bound by the integer pipe (10 ops/cell), not by the
// 9 LDS per cell of the table-in-smem form.  Used when 7*C registers fit (C <= 16, M <= 512).
template <int C, bool SMEM, bool TREG>
__global__ void __launch_bounds__(DP_THREADS) vit_kernel(const VitParams p)
{
  extern __shared__ __align__(16) int16_t s_vit[];
  __shared__ int s_ptab[C * 32];
  const int16_t *tw = p.tw, *rw = p.rw;
  if (SMEM) {
    const int ntw = 8 * C * 32, nrw = NTABROWS * C * 32;
    for (int i = threadIdx.x; i < ntw; i += blockDim.x) s_vit[i] = p.tw[i];
    for (int i = threadIdx.x; i < nrw; i += blockDim.x) s_vit[ntw + i] = p.rw[i];
    __syncthreads();
    tw = s_vit; rw = s_vit + ntw;
  }
  const int lane = threadIdx.x & 31;
  const int16_t *tBM = tw + 0 * C * 32 + lane, *tMM = tw + 1 * C * 32 + lane, *tIM = tw + 2 * C * 32 + lane,
                *tDM = tw + 3 * C * 32 + lane, *tMD = tw + 4 * C * 32 + lane, *tMI = tw + 5 * C * 32 + lane,
                *tII = tw + 6 * C * 32 + lane, *tDD = tw + 7 * C * 32 + lane;

  int rBM[TREG ? C : 1], rMM[TREG ? C : 1], rIM[TREG ? C : 1], rDM[TREG ? C : 1], rMD[TREG ? C : 1], rMI[TREG ? C : 1], rII[TREG ? C : 1];
  if (TREG) {
#pragma unroll
    for (int j = 0; j < C; j++) {
      rBM[j] = tBM[j * 32]; rMM[j] = tMM[j * 32]; rIM[j] = tIM[j * 32]; rDM[j] = tDM[j * 32];
      rMD[j] = tMD[j * 32]; rMI[j] = tMI[j * 32]; rII[j] = tII[j * 32];
    }
  }
#define T_(name, j) (TREG ? r##name[j] : (int) t##name[(j) * 32])
  // row-invariant sums of tDD for the lazy-F scan: PT[j] = sum of this lane's tDD[0..j-1] (shared memory, per lane),
  // S_lane = sum over all earlier lanes' cells, S_next = S_lane + this lane's total
  const int *s_pt = s_ptab + lane;
  int S_lane, S_next;
  {
    int acc = 0;
#pragma unroll
    for (int j = 0; j < C; j++) { if (threadIdx.x < 32) s_ptab[j * 32 + lane] = acc; acc += (int) tDD[j * 32]; }
    int incl = acc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    S_next = incl; S_lane = incl - acc;
    __syncthreads();
  }

  long long s, out;
  while (next_target(p.c, lane, &s, &out)) {
    const int L = p.c.len[s];
    const uint8_t *seq = p.c.res + p.c.off[s];
    const int xw_move = p.c.lc[L].xw_move;
    int mp[C], ip[C], dp[C];
#pragma unroll
    for (int j = 0; j < C; j++) mp[j] = ip[j] = dp[j] = VNEG;
    int xN = p.s.base_w, xB = (short) (xN + xw_move), xJ = VNEG, xC = VNEG, xE = VNEG;
    int status = P7GPU_OK;

    for (int i = 0; i < L; i++) {
      const int x = __ldg(seq + i);
      const int16_t *rwx = rw + (size_t) x * C * 32 + lane;
      int pm = __shfl_up_sync(0xffffffffu, mp[C - 1], 1);
      int pi = __shfl_up_sync(0xffffffffu, ip[C - 1], 1);
      int pd = __shfl_up_sync(0xffffffffu, dp[C - 1], 1);
      if (lane == 0) pm = pi = pd = VNEG;
      int xEl = VNEG, Dml = VNEG, dcv = VNEG;
#pragma unroll
      for (int j = 0; j < C; j++) {
        int sv = __viaddmax_s32(xB, T_(BM, j), VNEG);                  // vitfilter.c:146-150
        sv = __viaddmax_s32(pm, T_(MM, j), sv);
        sv = __viaddmax_s32(pi, T_(IM, j), sv);
        sv = __viaddmax_s32(pd, T_(DM, j), sv);
        sv = __viaddmax_s32(sv, (int) rwx[j * 32], VNEG);
        xEl = max(xEl, sv);
        pm = mp[j]; pi = ip[j]; pd = dp[j];                           // :155-157
        mp[j] = sv;
        dp[j] = dcv;                                                  // D(i,k) = M(i,k-1)+tMD so far (:160-161)
        dcv = __viaddmax_s32(sv, T_(MD, j), VNEG);                    // :166
        Dml = max(Dml, dcv);
        ip[j] = __viaddmax_s32(pi, T_(II, j), __viaddmax_s32(pm, T_(MI, j), VNEG));   // :170-171
      }
      {
        int d0 = __shfl_up_sync(0xffffffffu, dcv, 1);
        dp[0] = (lane == 0) ? VNEG : d0;
      }
      xE = __reduce_max_sync(0xffffffffu, xEl);
      if (xE >= 32767) { status = P7GPU_ERANGE; break; }              // :177
      // specials: int arithmetic truncated to int16_t on assignment (:178-181); N/C/J loops are 0
      xC = (short) max(xC, xE + p.s.xw_E_move);
      xJ = (short) max(xJ, xE + p.s.xw_E_loop);
      xB = (short) max(xJ + xw_move, xN + xw_move);
      const int Dmax = __reduce_max_sync(0xffffffffu, Dml);
      if (Dmax + p.s.ddbound_w > xB) {                                // lazy F, :198-232
        // D_k = max(D_k, D_{k-1} + tDD_{k-1}) to its fixed point = a max-plus prefix over k.  (1) lane-local prefix;
        // (2) the carry a lane offers its successor, A, moved into a frame where the (row-invariant) sums of tDD
        // cancel: B = A - S_{l+1}; an inclusive prefix-max over lanes (shfl_up hands a lane its own value when the
        // source is out of range, so no guard); (3) c_l = PM_{l-1} + S_l enters cell j as c_l + PT[j].
#pragma unroll
        for (int j = 1; j < C; j++) dp[j] = __viaddmax_s32(dp[j - 1], (int) tDD[(j - 1) * 32], dp[j]);
        int B = dp[C - 1] + (int) tDD[(C - 1) * 32] - S_next;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) B = max(B, __shfl_up_sync(0xffffffffu, B, d));
        int carry = __shfl_up_sync(0xffffffffu, B, 1) + S_lane;
        if (lane == 0) carry = VNOCARRY;
#pragma unroll
        for (int j = 0; j < C; j++) dp[j] = __viaddmax_s32(carry, s_pt[j * 32], dp[j]);
      }
    }
    if (lane == 0) {
      float sc;
      if (status == P7GPU_ERANGE) sc = CUDART_INF_F;
      else if (xC > -32768) {                                         // :240-247
        sc = (float) xC + (float) xw_move - (float) p.s.base_w;
        sc = __fdiv_rn(sc, p.s.scale_w);
        sc = __fsub_rn(sc, 3.0f);
      } else sc = -CUDART_INF_F;
      p.c.sc[out] = sc;
      p.c.status[out] = status;
    }
  }
#undef T_
}

template <int C>
static int launch_vit_c(p7gpu_ctx *ctx, cudaStream_t stream, const VitParams &p)
{
  const size_t smem = (size_t) (8 + NTABROWS) * C * 32 * sizeof(int16_t);
  const bool use_smem = smem + (size_t) C * 32 * 4 + 1024 <= (size_t) ctx->smem_optin;
  constexpr bool TREG = (C <= 16);
  auto kern = use_smem ? vit_kernel<C, true, TREG> : vit_kernel<C, false, false>;
  const size_t dyn = use_smem ? smem : 0;
  if (use_smem) P7G_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dyn));
  int occ = 0;
  P7G_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, DP_THREADS, dyn));
  if (occ < 1) { set_error("vit_kernel<%d>: zero occupancy", C); return P7GPU_ESYS; }
  long long blocks = (long long) ctx->sm_count * occ;
  long long need = (p.c.n + (DP_THREADS / 32) - 1) / (DP_THREADS / 32);
  if (need < blocks) blocks = need;
  if (blocks < 1) blocks = 1;
  { LaunchTimer t(ctx, P7GPU_KERNEL_VIT, stream); kern<<<(unsigned) blocks, DP_THREADS, dyn, stream>>>(p); }
  P7G_CUDA(cudaGetLastError());
  ctx->launches++;
  return P7GPU_OK;
}

int launch_vit(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
               float *d_sc, int32_t *d_status)
{
  if (n <= 0) return P7GPU_OK;
  P7G_CHECK(ensure_len_table(prof, db.max_length));
  const Lane &ln = ctx->lanes[lane];
  VitParams p;
  p.c.res = db.res; p.c.off = db.off; p.c.len = db.len; p.c.order = d_idx ? nullptr : db.order; p.c.idx = d_idx; p.c.first = db.first;
  p.c.n = n; p.c.lc = prof->d_len; p.c.sc = d_sc; p.c.status = d_status;
  p.c.work_counter = (unsigned long long *) (ln.counters + CTR_VIT_WORK);
  p.rw = prof->d_rw; p.tw = prof->d_tw; p.s = prof->vit;
  P7G_CUDA(cudaMemsetAsync(ln.counters + CTR_VIT_WORK, 0, 2 * sizeof(unsigned int), ln.stream));
  switch (prof->dpC) {
#define X(CV) case CV: return launch_vit_c<CV>(ctx, ln.stream, p);
    P7G_DP_C_CASES(X)
#undef X
  default: set_error("vit: unsupported cells-per-lane %d", prof->dpC); return P7GPU_EINVAL;
  }
}

}  // namespace p7g
