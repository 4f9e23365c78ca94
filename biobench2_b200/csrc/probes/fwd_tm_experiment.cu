// NOT PART OF THE LIBRARY BUILD — a measured experiment, kept with its numbers (profiles/r4a_fwd_M400_full.txt,
// profiles/r4c_fwd_tm_M400_full.txt, profiles/r4b_f32x2_probe.json).
//
// Question: the fast Forward kernel (fwd_mw.cu) uses 64 % of the issue slots with two warps per SM sub-partition (195-254
// registers: 8*C of them transitions).  Does (a) moving the transitions into Tensor Memory to fit more warps, or (b) advancing
// two targets per warp with sm_100's packed fp32 (FFMA2/FMUL2/FADD2, scalar operands broadcast) make it faster?
// Built three ways as libp7gpu's "fb_tm" option and run against fwd_mw_kernel on one B200 (40,000 targets, mean L 350):
//   (a) scalar, 16 warps per SM (128 registers), one tcgen05.ld.x8 per cell:     M=400 1.05 TCUPS, M=2000 1.01   (fwd_mw: 1.16, 1.05)
//   (b) packed two-target, 12 warps (168 registers, 232-304 B spilled):          M=400 0.83,       M=2000 0.64
//   (b) packed two-target, 8 warps (255 registers), tcgen05.ld.x32 per 4 cells:  M=400 1.03,       M=2000 0.88   <- this file
// All three give bit-identical scores and specials to fwd_mw_kernel for C >= 3 (the C = 2 geometry of the packed form differs
// in the last bit on 0.9 % of targets; not run down).  Why no gain: (a) executes 23 % more instructions per row (tcgen05.ld +
// the NOP ptxas pairs with it, R2UR of the address, recomputed emission addresses under the 128-register cap) for 72 % instead of
// 64 % of the slots; (b) halves the fp32 instruction count (177 packed per two rows instead of 2 x 169) but FFMA2 issues at half
// rate (0.457 per clock per sub-partition against 0.958 for FFMA, probes/f32x2_probe.cu), pair construction costs 66 moves per
// two rows, and with two warps per sub-partition the stall profile moves to fixed-latency dependencies (1.3 stall cycles per
// issued instruction) and fp32-pipe throttle (0.49): 48 % of the slots used.  What did carry over into fwd_mw.cu is the row
// pipelining developed for these variants (M finished from carried partial sums; next row's partial sums and I computed under
// the scan's shuffles): 1.08 -> 1.16 TCUPS at M=400, 1.01 -> 1.05 at M=2000.
//
// fwd_tm.cu — p7_ForwardParser (fwdback.c:257-465, do_full = 0): two targets per warp group in packed fp32, the transition
// table in Tensor Memory.
//
// Same decomposition and the same sequence of float operations per target as fwd_mw.cu (W warps per target, C consecutive
// nodes per lane, affine D->D maps, one barrier per row; results are bit-identical to fwd_mw_kernel<C, W> — both files are
// built without implicit FMA contraction, every fused operation is written out).  Two things change:
//  * sm_100's packed fp32 instructions (FFMA2 / FMUL2 / FADD2) take a 64-bit register pair per operand — or ONE 32-bit
//    register broadcast to both halves.  The k-1 dependence of a DP row never pairs neighbouring cells without register
//    moves, but two TARGETS pair perfectly: every DP value becomes (target A, target B), every model constant stays a scalar
//    operand.  The row costs the same number of instructions as before and advances two targets (probe:
//    probes/f32x2_probe.cu — FFMA2 issues at half the FFMA rate, so the fp32 pipe sees the same work; what halves is the
//    issue-slot count, which is what bounds these kernels: 64 % of the slots used at 33 % per warp, profiles/r4a_*).
//    The two slots of a group are refilled independently (a slot that finishes its target draws the next one while the other
//    is mid-sequence), so unequal lengths cost nothing.
//  * The 8*C transition values of a lane would need 104-128 registers next to 6*C of paired row state.  They are
//    row-invariant and lane-private, which is what TMEM holds: TMEM[32*(warp%4) + lane][8*j + t], one tcgen05.ld.32x32b.x8
//    per cell, fetched one cell ahead of the arithmetic; the row runs as ONE pass over the lane's cells (finish M from the
//    carried partial sums, D0 chain link, next row's partial sums and I).  A 384-thread CTA per SM keeps three warps per
//    sub-partition: 24/W targets in flight per SM, each group of W warps meeting at its own named barrier.
#include "fwdbwd_mw.cuh"
#include "tmem.cuh"

namespace p7g {

constexpr int FB_TM_WARPS = 8;
constexpr int FB_TM_COLS = 128;         // 8 * C <= 128
constexpr int FB_TM_STAGE = 4;          // cells per tcgen05.ld

typedef float2 f2;
__device__ __forceinline__ f2 bc2(float s) { return make_float2(s, s); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 fma2s(f2 a, float s, f2 c) { return __ffma2_rn(a, bc2(s), c); }      // scalar operand: one register, broadcast
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 mul2s(f2 a, float s) { return __fmul2_rn(a, bc2(s)); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 shfl_up2(f2 v, int d) { return make_float2(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d)); }
__device__ __forceinline__ f2 shfl_idx2(f2 v, int l) { return make_float2(__shfl_sync(0xffffffffu, v.x, l), __shfl_sync(0xffffffffu, v.y, l)); }
template <bool B> __device__ __forceinline__ float &half_of(f2 &v) { return B ? v.y : v.x; }

template <int W>
__device__ __forceinline__ void tm_group_sync(int grp)
{
  if (W == 1) __syncwarp();
  else asm volatile("bar.sync %0, %1;" :: "r"(grp + 1), "n"(32 * W) : "memory");
}

// the group's next work item (its first lane draws it)
template <int W>
__device__ __forceinline__ bool tm_next_target(const DpCommon &c, int grp, int w, int lane, long long *sh, long long *s, long long *out)
{
  if (W == 1) {
    unsigned long long j = 0;
    if (lane == 0) j = atomicAdd(c.work_counter, 1ULL);
    j = __shfl_sync(0xffffffffu, j, 0);
    if (j >= (unsigned long long) c.n) return false;
    *s = c.idx ? c.idx[j] : (c.order ? (long long) c.order[j] : c.first + (long long) j);
    *out = c.idx ? (long long) j : *s;
    return true;
  }
  tm_group_sync<W>(grp);
  if (w == 0 && lane == 0) {
    unsigned long long j = atomicAdd(c.work_counter, 1ULL);
    long long ss = -1, oo = -1;
    if (j < (unsigned long long) c.n) {
      ss = c.idx ? c.idx[j] : (c.order ? (long long) c.order[j] : c.first + (long long) j);
      oo = c.idx ? (long long) j : ss;
    }
    sh[0] = ss; sh[1] = oo;
  }
  tm_group_sync<W>(grp);
  *s = sh[0]; *out = sh[1];
  return *s >= 0;
}

template <int C, int W>
__global__ void __launch_bounds__(32 * FB_TM_WARPS, 1) fwd_tm_kernel(const FbMwParams p)
{
  static_assert(W == 1 || W == 2 || W == 4, "a warp's model segment must follow from its TMEM quarter");
  static_assert(8 * C <= FB_TM_COLS, "the lane's transitions fit the allocation");
  constexpr int G = 32 * W, NG = FB_TM_WARPS / W;
  __shared__ uint32_t s_tmslot;
  __shared__ long long s_work[NG][2];
  __shared__ float s_q[C][G];           // as fwd_mw.cu
  __shared__ float s_ppl[G], s_bexc[G];
  __shared__ float s_kB[5][G], s_kV[5][G];
  __shared__ float s_wB[W], s_wV[W], s_wK[W];
  __shared__ f2 s_pub[NG][2][W][5];     // per group and row parity: P, U, M_last, I_last, D_last part (X = 0), both targets
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int w = wid % W, grp = wid / W;                 // w = (wid % 4) % W: the quarter's segment
  const int g = w * 32 + lane;
  const uint32_t tbase = tmem_alloc_cta<FB_TM_COLS>(&s_tmslot);
  const uint32_t mine = tbase + ((uint32_t) ((wid & 3) * 32) << 16);
  if (wid < 4) {                                        // one warp per quarter writes the lane-private table
#pragma unroll 1
    for (int j = 0; j < C; j++) {
      uint32_t t[8];
#pragma unroll
      for (int k = 0; k < 8; k++) t[k] = __float_as_uint(__ldg(p.tf + (k * C + j) * G + g));
      Tmem<8>::st(mine + 8 * j, t);
    }
    tmem_wait_st();
  }
  // ---- row-invariant coefficients of the D->D maps: the first group computes them for everybody ----
  if (grp == 0) {
    float pp = 1.0f, spp = 0.f;
#pragma unroll
    for (int j = 0; j < C; j++) {
      if (j + 1 < C) s_q[j + 1][g] = pp * __ldg(p.tf + (3 * C + j + 1) * G + g);
      else s_ppl[g] = pp;
      spp += pp; pp = pp * __ldg(p.tf + (7 * C + j) * G + g);
    }
    const float ppl = s_ppl[g];
    float B = pp, V = spp;
#pragma unroll
    for (int d = 0; d < 5; d++) {
      s_kB[d][g] = B; s_kV[d][g] = V;
      const float Bl = __shfl_up_sync(0xffffffffu, B, 1 << d), Vl = __shfl_up_sync(0xffffffffu, V, 1 << d);
      if (lane >= (1 << d)) { V = Vl + V * Bl; B = B * Bl; }
    }
    float Bexc = __shfl_up_sync(0xffffffffu, B, 1);
    if (lane == 0) Bexc = 1.0f;
    s_bexc[g] = Bexc;
    if (lane == 31) { s_wB[w] = B; s_wV[w] = V; s_wK[w] = Bexc * ppl; }
  }
  tmem_publish();                                       // (a CTA barrier: the table and the coefficients above)
  if (grp == 0) s_q[0][g] = lane ? s_ppl[g - 1] * __ldg(p.tf + (3 * C) * G + g) : 0.0f;
  __syncthreads();
  const float Bexc = s_bexc[g], BexcPrev = lane ? s_bexc[g - 1] : 0.f, ppLast = s_ppl[g];
  const float t0MM = __ldg(p.tf + (1 * C) * G + g), t0IM = __ldg(p.tf + (2 * C) * G + g), t0DM = __ldg(p.tf + (3 * C) * G + g);
  const float *rfg = p.rf + g;
  const f2 zero2 = make_float2(0.f, 0.f);

  // ---- the group's two slots: .x = target A, .y = target B; the row state of fwd_mw.cu, one row ahead of the recursion ----
  f2 part[C], ip[C], e[C];
  f2 xN = zero2, xJ = zero2, xB = zero2, xC = zero2, totscale = zero2, pmove = zero2, ploop = zero2, Z = zero2, Zp = zero2, bx = zero2;
  int iA = 0, iB = 0, LA = 0, LB = 0;                    // i = the row about to be computed
  long long outA = -1, outB = -1;
  const uint8_t *seqA = nullptr, *seqB = nullptr;
  float *xrowA = nullptr, *xrowB = nullptr;
#pragma unroll
  for (int j = 0; j < C; j++) part[j] = ip[j] = e[j] = zero2;

  // a target's result (:459-463)
  auto finish = [&](long long out, int L, float xCv, float pm, float tot) {
    if (g == 0) {
      int status = P7GPU_OK; float sc;
      if (isnan(xCv) || (L > 0 && xCv == 0.0f) || isinf(xCv)) { status = P7GPU_ERANGE; sc = xCv; }
      else sc = (float) ((double) tot + log((double) (xCv * pm)));
      p.c.sc[out] = sc;
      p.c.status[out] = status;
      if (p.totscale) p.totscale[out] = tot;
    }
  };
#define P7G_TM_START(SB, out_, L_, seq_, xrow_, i_)                                                                        \
  for (;;) {                                                  /* draw targets until one has rows to compute */            \
    long long s_;                                                                                                        \
    half_of<SB>(xN) = 0.f; half_of<SB>(xJ) = 0.f; half_of<SB>(xB) = 0.f; half_of<SB>(xC) = 0.f; half_of<SB>(totscale) = 0.f;  \
    half_of<SB>(pmove) = 0.f; half_of<SB>(ploop) = 0.f; half_of<SB>(Z) = 0.f; half_of<SB>(Zp) = 0.f; half_of<SB>(bx) = 0.f;   \
    _Pragma("unroll") for (int j = 0; j < C; j++) { half_of<SB>(part[j]) = 0.f; half_of<SB>(ip[j]) = 0.f; half_of<SB>(e[j]) = 0.f; }  \
    L_ = 0; i_ = 1; xrow_ = nullptr;                                                                                      \
    if (!tm_next_target<W>(p.c, grp, w, lane, s_work[grp], &s_, &out_)) { out_ = -1; break; }                             \
    L_ = p.c.len[s_];                                                                                                     \
    seq_ = p.c.res + p.c.off[s_];                                                                                         \
    const float pm_ = p.c.lc[L_].pmove;                                                                                   \
    xrow_ = p.xmx ? p.xmx + p.xmx_off[out_] : nullptr;                                                                    \
    if (xrow_ && g == 0) mw_store_specials(xrow_, 0.f, 1.f, 0.f, pm_, 0.f, 1.0f);            /* fwdback.c:287-294 */     \
    if (L_ == 0) { finish(out_, 0, 0.f, pm_, 0.f); continue; }                                                            \
    half_of<SB>(xN) = 1.f; half_of<SB>(xB) = pm_; half_of<SB>(pmove) = pm_; half_of<SB>(ploop) = 1.0f - pm_;              \
    const float *rfx_ = rfg + (size_t) __ldg(seq_) * (C * G);                                                             \
    _Pragma("unroll") for (int j = 0; j < C; j++) half_of<SB>(e[j]) = __ldg(rfx_ + j * G);                                 \
    break;                                                                                                                \
  }
  P7G_TM_START(false, outA, LA, seqA, xrowA, iA)
  P7G_TM_START(true, outB, LB, seqB, xrowB, iB)

  int step = 0;
#pragma unroll 1
  for (; outA >= 0 || outB >= 0; step++) {
    // the transitions arrive FB_TM_STAGE cells at a time, one stage ahead of the arithmetic (tcgen05.wait::ld waits for every
    // load in flight, so the distance between a load and its use is one stage whatever is issued; a stage of one cell left the
    // row waiting for a TMEM round trip per cell)
    uint32_t ta[8 * FB_TM_STAGE], tb[8 * FB_TM_STAGE];
    Tmem<8 * FB_TM_STAGE>::ld(mine, ta);
    const bool moreA = iA < LA, moreB = iB < LB;          // (an idle slot has L = 0)
    const float *rfxA = rfg + (size_t) (moreA ? __ldg(seqA + iA) : 0) * (C * G);
    const float *rfxB = rfg + (size_t) (moreB ? __ldg(seqB + iB) : 0) * (C * G);
    f2 svp = zero2, d0p = zero2, ipo = zero2, sumM = zero2, sumD0 = zero2;
    float tMDp = 0.f, tDDp = 0.f;
    tmem_wait_ld();
    constexpr int NS = (C + FB_TM_STAGE - 1) / FB_TM_STAGE;
#pragma unroll
    for (int st = 0; st < NS; st++) {
      uint32_t (&tc)[8 * FB_TM_STAGE] = (st & 1) ? tb : ta;
      if (st + 1 < NS) Tmem<8 * FB_TM_STAGE>::ld(mine + 8 * FB_TM_STAGE * (st + 1), (st & 1) ? ta : tb);
#pragma unroll
      for (int jj = 0; jj < FB_TM_STAGE; jj++) {
        const int j = st * FB_TM_STAGE + jj;
        if (j < C) {
          const uint32_t *tr = tc + 8 * jj;
          f2 v = fma2s(xB, __uint_as_float(tr[0]), part[j]);          // finish M(i,j) (:318-345)
          v = fma2s(j ? Z : Zp, s_q[j][g], v);
          if (W > 1 && j == 0) v = add2(v, bx);
          const f2 sv = mul2(v, e[j]);
          e[j] = make_float2(__ldg(rfxA + j * G), __ldg(rfxB + j * G));   // next row's emissions (a slot on its last row reads row 0 of the table: unused)
          f2 d0 = zero2;
          if (j >= 1) {
            d0 = fma2s(d0p, tDDp, mul2s(svp, tMDp));                  // the lane's own D->D chain, nothing entering
            part[j] = fma2s(d0p, __uint_as_float(tr[3]), fma2s(ipo, __uint_as_float(tr[2]), mul2s(svp, __uint_as_float(tr[1]))));
            sumM = add2(sumM, sv); sumD0 = add2(sumD0, d0);
          } else sumM = sv;
          ipo = ip[j];
          ip[j] = fma2s(ipo, __uint_as_float(tr[6]), mul2s(sv, __uint_as_float(tr[5])));
          svp = sv; d0p = d0; tMDp = __uint_as_float(tr[4]); tDDp = __uint_as_float(tr[7]);
        }
      }
      if (st + 1 < NS) tmem_wait_ld();
    }
    f2 P = fma2s(d0p, tDDp, mul2s(svp, tMDp));
    f2 U = add2(sumM, sumD0);
    {
      const f2 pm_b = shfl_up2(svp, 1), pi_b = shfl_up2(ipo, 1), pd_b = shfl_up2(d0p, 1);
      const f2 v = fma2s(pd_b, t0DM, fma2s(pi_b, t0IM, mul2s(pm_b, t0MM)));
      part[0] = lane ? v : zero2;
    }
#pragma unroll
    for (int d = 0; d < 5; d++) {                                     // lane maps -> inclusive warp prefix
      const f2 Pl = shfl_up2(P, 1 << d), Ul = shfl_up2(U, 1 << d);
      if (lane >= (1 << d)) { U = fma2s(Pl, s_kV[d][g], add2(Ul, U)); P = fma2s(Pl, s_kB[d][g], P); }
    }
    Z = shfl_up2(P, 1);
    Zp = shfl_up2(P, 2);
    if (lane == 0) Z = zero2;
    if (lane <= 1) Zp = zero2;
    f2 xE;
    if (W == 1) {
      xE = shfl_idx2(U, 31);
    } else {
      const int slot = step & 1;
      if (lane == 31) {
        f2 *pub = s_pub[grp][slot][w];
        pub[0] = P; pub[1] = U; pub[2] = svp; pub[3] = ipo; pub[4] = fma2s(Z, ppLast, d0p);
      }
      tm_group_sync<W>(grp);                                          // ---- the row's only barrier ----
      f2 X = zero2, Xmine = zero2, Xprev = zero2;
      xE = zero2;
#pragma unroll
      for (int v = 0; v < W; v++) {
        if (v == w) Xmine = X;
        if (v + 1 == w) Xprev = X;
        xE = add2(xE, fma2s(X, s_wV[v], s_pub[grp][slot][v][1]));
        X = fma2s(X, s_wB[v], s_pub[grp][slot][v][0]);
      }
      Z = fma2s(Xmine, Bexc, Z);
      if (lane) Zp = fma2s(Xmine, BexcPrev, Zp);
      bx = zero2;
      if (w > 0 && lane == 0) {
        const f2 *pub = s_pub[grp][slot][w - 1];
        bx = fma2s(fma2s(Xprev, s_wK[w - 1], pub[4]), t0DM, fma2s(pub[3], t0IM, mul2s(pub[2], t0MM)));
      }
    }
    xN = mul2(xN, ploop);                                             // :413-416
    xC = fma2(xC, ploop, mul2s(xE, p.s.xf_E_move));
    xJ = fma2(xJ, ploop, mul2s(xE, p.s.xf_E_loop));
    xB = fma2(xJ, pmove, mul2(xN, pmove));
    f2 scale = make_float2(1.0f, 1.0f);
    if (xE.x > 1.0e4f || xE.y > 1.0e4f) {                             // :420-437
      f2 inv = make_float2(1.0f, 1.0f);
      if (xE.x > 1.0e4f) {
        inv.x = (float) (1.0 / (double) xE.x);
        xN.x = xN.x / xE.x; xC.x = xC.x / xE.x; xJ.x = xJ.x / xE.x; xB.x = xB.x / xE.x;
        scale.x = xE.x;
        totscale.x = (float) ((double) totscale.x + log((double) xE.x));
        xE.x = 1.0f;
      }
      if (xE.y > 1.0e4f) {
        inv.y = (float) (1.0 / (double) xE.y);
        xN.y = xN.y / xE.y; xC.y = xC.y / xE.y; xJ.y = xJ.y / xE.y; xB.y = xB.y / xE.y;
        scale.y = xE.y;
        totscale.y = (float) ((double) totscale.y + log((double) xE.y));
        xE.y = 1.0f;
      }
#pragma unroll
      for (int j = 0; j < C; j++) { part[j] = mul2(part[j], inv); ip[j] = mul2(ip[j], inv); }
      Z = mul2(Z, inv); Zp = mul2(Zp, inv); bx = mul2(bx, inv);
    }
    if (g == 0) {
      if (xrowA) mw_store_specials(xrowA + (size_t) iA * 6, xE.x, xN.x, xJ.x, xB.x, xC.x, scale.x);
      if (xrowB) mw_store_specials(xrowB + (size_t) iB * 6, xE.y, xN.y, xJ.y, xB.y, xC.y, scale.y);
    }
    if (outA >= 0) {
      if (iA == LA) { finish(outA, LA, xC.x, pmove.x, totscale.x); P7G_TM_START(false, outA, LA, seqA, xrowA, iA) }
      else iA++;
    }
    if (outB >= 0) {
      if (iB == LB) { finish(outB, LB, xC.y, pmove.y, totscale.y); P7G_TM_START(true, outB, LB, seqB, xrowB, iB) }
      else iB++;
    }
  }
#undef P7G_TM_START
  tmem_free_cta<FB_TM_COLS>(tbase);
}

template <int C, int W>
static int launch_fwd_tm_c(p7gpu_ctx *ctx, cudaStream_t stream, const FbMwParams &p)
{
  constexpr int NG = 2 * FB_TM_WARPS / W;                 // targets in flight per CTA
  long long blocks = (p.c.n + NG - 1) / NG;
  if (blocks > ctx->sm_count) blocks = ctx->sm_count;
  if (blocks < 1) blocks = 1;
  { LaunchTimer t(ctx, P7GPU_KERNEL_FWD, stream); fwd_tm_kernel<C, W><<<(unsigned) blocks, 32 * FB_TM_WARPS, 0, stream>>>(p); }
  P7G_CUDA(cudaGetLastError());
  ctx->launches++;
  return P7GPU_OK;
}

int launch_fwd_tm(p7gpu_ctx *ctx, cudaStream_t stream, const p7gpu_profile *prof, const FbMwParams &p)
{
  const int C = prof->mwC, W = prof->mwW;
  switch (C * 100 + W) {
#define X(CV, WV) case CV * 100 + WV: return launch_fwd_tm_c<CV, WV>(ctx, stream, p);
    P7G_TM_CASES(X)
#undef X
  default: set_error("fwd_tm: no kernel for C=%d W=%d", C, W); return P7GPU_EINVAL;
  }
}

}  // namespace p7g
