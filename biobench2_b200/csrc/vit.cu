// vit.cu — sm_100a kernel for p7_ViterbiFilter (HMMER/src/impl_sse/vitfilter.c:83-249).
//
// One warp per target, lane l owns nodes k = l*C+1 .. l*C+C (dp_common.cuh).  The 16-bit saturating
// arithmetic of the reference is reproduced exactly in 32-bit lanes with DPX VIADDMNMX (max(a+b,c)):
// every adds_epi16(a,t) with t <= 0 equals max(a+t, -32768), and the only add that can saturate upward
// (the emission) makes xE >= 32767, which returns eslERANGE for the row before the value is used again.
// The lazy-F test of vitfilter.c:198-199 is evaluated as written; when it fires, the D->D fixed point
// of :203-226 is obtained with a max-plus prefix scan over the warp (exact: see DESIGN.md).
#include "dp_common.cuh"
#include "tmem.cuh"
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
  if (p.c.pairs && *p.c.pair_count == 0u) return;                     // redo pass with nothing to redo
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


// ------------------------------------------------------------------------------------------------
// Packed fast path: two model nodes per 32-bit register (DPX VIADDMNMX.S16x2 / VIMNMX3.S16x2).
//
// Lane l, register j, half h holds node k = 1 + h*32*C2 + l*C2 + j.  The reference's saturating int16 arithmetic
// needs 17 bits of intermediate range, so values are kept in a per-target window instead: w = v - SH with
// w in [W0, W1], W0 = -32768 + T, W1 = W0 + T, every transition / emission clamped at -T, so that no packed add can
// wrap and a clamped ("minus infinity") add always lands on the floor W0.  The floor corresponds to the reference
// value Fl = base_w + xw_move(L) + floff, chosen (p7gpu_host.cu: build_vit_packed) below everything that can influence
// a later cell, xE, or the lazy-F test; stored cells then equal max(reference cell, Fl) and xE, xC, xJ, xB and the lazy-F
// decisions are those of the reference (DESIGN.md 4.3 gives the induction).  A target whose row maximum leaves
// (W0, W1] — a very high score, possible overflow, or a row of impossible residues — is appended to the redo list and
// recomputed by the exact 32-bit kernel above.
struct VitpParams {
  DpCommon c;
  const uint32_t *rwp, *twp, *scanw;
  VitScalars s;
  VitPacked w;
  long long *redo_list; unsigned int *redo_count;
  uint32_t k16;         // 65536, as a kernel parameter so that ptxas keeps x * k16 / mulhi(x, k16) on the multiplier pipe
};

__device__ __forceinline__ uint32_t pack2(int v) { return __byte_perm((uint32_t) v, 0u, 0x1010); }
// max over the warp of both halves of v.  v * 65536 puts the low half on top, so the high half of the packed maximum is
// max(lo, hi); a 32-bit signed maximum over the warp is decided by the high halves, and a multiply-high by 65536 brings it
// down (arithmetic >> 16).  k16 is 65536 in a register the compiler cannot see through: both multiplies stay on the
// multiplier pipe instead of turning into shifts on the integer pipe.
__device__ __forceinline__ int warp_max16x2(uint32_t v, uint32_t k16)
{
  const uint32_t m = __vmaxs2(v, v * k16);
  int r = __reduce_max_sync(0xffffffffu, (int) m), e;
  asm("mul.hi.s32 %0, %1, %2;" : "=r"(e) : "r"(r), "r"(k16));
  return e;
}

// The integer pipe of sm_100 issues every instruction (plain or packed DPX) at 16 lanes per clock, and the row loop is bound
// by it (ncu: ALU 83 % busy), so everything that is not one of the 9 max-plus operations per register is kept off that pipe:
// residues arrive as byte loads one row ahead (no shift/mask), the rotation by one lane fixes lane 0 with one multiply-add
// (r * 65536 + floor moves lane 31's low half up; r * 1 + 0 elsewhere), the specials live in the window domain (no
// int16 re-truncation: max(xC, .) and max(xJ, .) never leave [-32768, 32766], xB >= xN + tNB), and the >> 16 after the warp
// reductions is a multiply-high.
constexpr int VITP_BIG_THREADS = 256;   // C2 > 14: few CTAs per SM (the emission table fills shared memory), so more warps share it
// C2 > 14 (M > 896): the 8 transition vectors of a lane do not fit its registers.  They are row-invariant and lane-private, which
// is what Tensor Memory holds best: TMEM[lane][8*j + t], one tcgen05.ld.32x32b.x8 per register-row instead of 8 LDS.32 (the
// shared-memory form was LDS-bound at 86 %, profiles/r2n_vitp_M2000_full.txt); emissions and the lazy-F copy of tDD stay in
// shared memory.
__host__ __device__ constexpr int vitp_tm_cols(int C2) { return 8 * C2 <= 128 ? 128 : (8 * C2 <= 256 ? 256 : 512); }

template <int C2, bool TREG>
__global__ void __launch_bounds__(TREG ? DP_THREADS : VITP_BIG_THREADS, TREG ? (C2 <= 7 ? 4 : 2) : 1) vitp_kernel(const VitpParams p)
{
  extern __shared__ __align__(16) uint32_t s_vp[];
  __shared__ uint32_t s_tmslot;
  constexpr int TCOLS = vitp_tm_cols(C2);
  uint32_t tbase = 0, mine = 0;
  {
    const int nrw = NTABROWS * C2 * 32;
    for (int i = threadIdx.x; i < nrw; i += blockDim.x) s_vp[i] = __ldg(p.rwp + i);
    if (!TREG) { for (int i = threadIdx.x; i < C2 * 32; i += blockDim.x) s_vp[nrw + i] = __ldg(p.twp + 7 * C2 * 32 + i); }   // tDD for lazy F
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  if (!TREG) {
    tbase = tmem_alloc_cta<TCOLS>(&s_tmslot);
    mine = tbase + ((uint32_t) (((threadIdx.x >> 5) & 3) * 32) << 16);
    if ((threadIdx.x >> 5) < 4) {                                      // one warp per quarter writes the lane-private table
      for (int j = 0; j < C2; j++) {
        uint32_t t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = __ldg(p.twp + (k * C2 + j) * 32 + lane);
        Tmem<8>::st(mine + 8 * j, t);
      }
      tmem_wait_st();
    }
    tmem_publish();
  }
  const int prevl = (lane + 31) & 31;
  const uint32_t *s_rw = s_vp + lane;
  const uint32_t *s_dd = s_vp + NTABROWS * C2 * 32 + lane;
  uint32_t rBM[TREG ? C2 : 1], rMM[TREG ? C2 : 1], rIM[TREG ? C2 : 1], rDM[TREG ? C2 : 1], rMD[TREG ? C2 : 1],
           rMI[TREG ? C2 : 1], rII[TREG ? C2 : 1], rDD[TREG ? C2 : 1];
  if (TREG) {
#pragma unroll
    for (int j = 0; j < C2; j++) {
      rBM[j] = __ldg(p.twp + (0 * C2 + j) * 32 + lane); rMM[j] = __ldg(p.twp + (1 * C2 + j) * 32 + lane);
      rIM[j] = __ldg(p.twp + (2 * C2 + j) * 32 + lane); rDM[j] = __ldg(p.twp + (3 * C2 + j) * 32 + lane);
      rMD[j] = __ldg(p.twp + (4 * C2 + j) * 32 + lane); rMI[j] = __ldg(p.twp + (5 * C2 + j) * 32 + lane);
      rII[j] = __ldg(p.twp + (6 * C2 + j) * 32 + lane); rDD[j] = __ldg(p.twp + (7 * C2 + j) * 32 + lane);
    }
  }
#define TP_(name, t, j) (TREG ? r##name[j] : tr[t])            /* inside the j loop of a row: tr = this register-row's 8 transitions */
#define TDD_(j) (TREG ? rDD[j] : s_dd[(j) * 32])
  uint32_t wd[5];
#pragma unroll
  for (int d = 0; d < 5; d++) wd[d] = __ldg(p.scanw + d * 32 + lane);
  const uint32_t wh = __ldg(p.scanw + 5 * 32 + lane);
  const int W0 = p.w.W0;
  const uint32_t FLv = pack2(W0);
  // s_pt[j][lane] = (clamped) sum of the lane's tDD[0..j-1], both halves: what a carry entering the lane has paid by cell j
  __shared__ uint32_t s_ptp[C2 * 32];
  if (threadIdx.x < 32) {
    int lo = 0, hi = 0;
#pragma unroll
    for (int j = 0; j < C2; j++) {
      s_ptp[j * 32 + lane] = ((uint32_t) max(lo, -p.w.T) & 0xffffu) | ((uint32_t) max(hi, -p.w.T) << 16);
      const uint32_t t = TDD_(j);
      lo += (int) (short) (t & 0xffffu); hi += (int) t >> 16;
    }
  }
  __syncthreads();
  const uint32_t *s_pt = s_ptp + lane;
  // lane l receives lane l-1's register; lane 0 receives (floor, low half of lane 31) = nodes 0 and 32*C2
  uint32_t rmul = lane == 0 ? 0x10000u : 1u;
  const uint32_t radd = lane == 0 ? (FLv & 0xffffu) : 0u;
  const uint32_t k16 = p.k16;
  asm("" : "+r"(rmul));                                                // keep it a multiply (a shift would go to the integer pipe)
#define ROT_(v) (__shfl_sync(0xffffffffu, (v), prevl) * rmul + radd)
  const int cEm = p.s.xw_E_move, cEl = p.s.xw_E_loop, ddb = p.s.ddbound_w;

  long long s, out;
  while (next_target(p.c, lane, &s, &out)) {
    const int L = p.c.len[s];
    const uint8_t *seq = p.c.res + p.c.off[s];
    const int xw_move = p.c.lc[L].xw_move;
    const int Fl = p.s.base_w + xw_move + p.w.floff;
    const int SH = Fl - W0;                                            // reference value = window value + SH
    const unsigned span = (unsigned) max(0, min(p.w.W1, 32766 - SH) - W0);   // row maxima must stay in (W0, W0 + span]: inside the
    bool redo = Fl < -32768;                                           // window and below the reference's overflow test
    uint32_t mp[C2], ip[C2], dp[C2];
#pragma unroll
    for (int j = 0; j < C2; j++) mp[j] = ip[j] = dp[j] = FLv;
    // specials in window units.  vitfilter.c:178-181 truncates to int16 on assignment; max(xC, .) and max(xJ, .) start at
    // -32768 and are fed by xE + (a cost <= 0) with xE <= 32766, and xB >= xN + tNB, so no truncation ever changes a value
    const int xNB = p.s.base_w + xw_move - SH;
    int xB = xNB, xJ = VNEG - SH, xC = VNEG - SH;

    // one DP row; false when the row maximum leaves the window (or reaches the reference's overflow test, :177)
    uint32_t pmN = FLv, piN = FLv;                                     // row i-1's M and I, already rotated by one lane
    auto row = [&](const uint32_t x) -> bool {
      const uint32_t *rwx = reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(s_rw) + x * (C2 * 128));
      const uint32_t xBv = pack2(xB);
      uint32_t pm = pmN;
      uint32_t pi = piN;
      uint32_t pd = ROT_(dp[C2 - 1]);
      uint32_t xEl = FLv, Dml = FLv, dcv = FLv, pendE = 0u, pendD = 0u;
      uint32_t tr[8], tn[8];
      if (!TREG) { Tmem<8>::ld(mine, tr); tmem_wait_ld(); }
#pragma unroll
      for (int j = 0; j < C2; j++) {
        if (!TREG && j + 1 < C2) Tmem<8>::ld(mine + 8 * (j + 1), tn);  // next register-row's transitions, under this one's arithmetic
        uint32_t sv = __viaddmax_s16x2(xBv, TP_(BM, 0, j), FLv);       // vitfilter.c:146-150
        sv = __viaddmax_s16x2(pm, TP_(MM, 1, j), sv);
        sv = __viaddmax_s16x2(pi, TP_(IM, 2, j), sv);
        sv = __viaddmax_s16x2(pd, TP_(DM, 3, j), sv);
        sv = __viaddmax_s16x2(sv, rwx[j * 32], FLv);
        pm = mp[j]; pi = ip[j]; pd = dp[j];                            // :155-157
        mp[j] = sv;
        dp[j] = dcv;                                                   // D(i,k) = M(i,k-1)+tMD so far (:160-161)
        dcv = __viaddmax_s16x2(sv, TP_(MD, 4, j), FLv);                // :166
        ip[j] = __viaddmax_s16x2(pi, TP_(II, 6, j), __viaddmax_s16x2(pm, TP_(MI, 5, j), FLv));   // :170-171
        // row maxima of M and of the provisional D, two new values per three-input max
        if (j == 0) { xEl = sv; Dml = dcv; }
        else if (j & 1) { pendE = sv; pendD = dcv; }
        else { xEl = __vimax3_s16x2(xEl, pendE, sv); Dml = __vimax3_s16x2(Dml, pendD, dcv); }
        if (!TREG && j + 1 < C2) {
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 8; k++) tr[k] = tn[k];
        }
      }
      if (C2 > 1 && !(C2 & 1)) { xEl = __vmaxs2(xEl, pendE); Dml = __vmaxs2(Dml, pendD); }
      dp[0] = ROT_(dcv);
      pmN = ROT_(mp[C2 - 1]); piN = ROT_(ip[C2 - 1]);                  // for the next row, under the reductions' latency
      const int xE = warp_max16x2(xEl, k16);
      const int Dmax = warp_max16x2(Dml, k16);
      if ((unsigned) (xE - W0 - 1) >= span) return false;
      xC = __viaddmax_s32(xE, cEm, xC);                                // :178-181
      xJ = __viaddmax_s32(xE, cEl, xJ);
      xB = __viaddmax_s32(xJ, xw_move, xNB);
      if (Dmax + ddb > xB) {                                           // lazy F, :198-232
        // max-plus prefix over k: (1) inside the lane; (2) X = value leaving the lane, Kogge-Stone over lanes with the
        // row-invariant lane-span sums wd[] (both halves at once); (3) the low half's total enters the high half (wh);
        // (4) the carry entering each lane reaches cell j after paying the lane's tDD[0..j-1] (s_pt).
#pragma unroll
        for (int j = 1; j < C2; j++) dp[j] = __viaddmax_s16x2(dp[j - 1], TDD_(j - 1), dp[j]);
        uint32_t X = __viaddmax_s16x2(dp[C2 - 1], TDD_(C2 - 1), FLv);
#pragma unroll
        for (int d = 0; d < 5; d++) X = __viaddmax_s16x2(__shfl_up_sync(0xffffffffu, X, 1 << d), wd[d], X);
        const uint32_t u = __shfl_sync(0xffffffffu, X, 31) * 0x10000u + (FLv & 0xffffu);   // (floor, X_lo(31))
        X = __viaddmax_s16x2(u, wh, X);
        const uint32_t carry = ROT_(X);                                // lane 0: (floor, X_lo(31)) = u again (step 3 left X_lo alone)
#pragma unroll
        for (int j = 0; j < C2; j++) dp[j] = __viaddmax_s16x2(carry, s_pt[j * 32], dp[j]);
      }
      return true;
    };

    // residues as byte loads, fetched one group of rows ahead (seq[L] is the sentinel byte, so reading it is safe)
    if (!redo) {
      int i = 0;
      uint32_t xn = __ldg(seq);
      if (TREG) {
        for (; i + 4 <= L; i += 4) {
          const uint32_t x0 = xn, x1 = __ldg(seq + i + 1), x2 = __ldg(seq + i + 2), x3 = __ldg(seq + i + 3);
          xn = __ldg(seq + i + 4);
          if (!(row(x0) && row(x1) && row(x2) && row(x3))) { redo = true; break; }
        }
      }
      if (!redo)
        for (; i < L; i++) {
          const uint32_t x = xn;
          xn = __ldg(seq + i + 1);
          if (!row(x)) { redo = true; break; }
        }
    }
    if (lane == 0) {
      if (redo) {
        const unsigned int t = atomicAdd(p.redo_count, 1u);
        p.redo_list[2 * (long long) t] = s; p.redo_list[2 * (long long) t + 1] = out;
      } else {
        float sc;
        if (xC + SH > -32768) {                                        // :240-247
          sc = (float) (xC + SH) + (float) xw_move - (float) p.s.base_w;
          sc = __fdiv_rn(sc, p.s.scale_w);
          sc = __fsub_rn(sc, 3.0f);
        } else sc = -CUDART_INF_F;
        p.c.sc[out] = sc;
        p.c.status[out] = P7GPU_OK;
      }
    }
  }
#undef TP_
#undef TDD_
#undef ROT_
  if (!TREG) tmem_free_cta<TCOLS>(tbase);
}

template <int C2>
static int launch_vitp_c(p7gpu_ctx *ctx, cudaStream_t stream, const VitpParams &p, bool *launched)
{
  constexpr bool TREG = (C2 <= 14);
  constexpr int NT = TREG ? DP_THREADS : VITP_BIG_THREADS;
  const size_t smem = (size_t) (NTABROWS + (TREG ? 0 : 1)) * C2 * 32 * sizeof(uint32_t);
  *launched = false;
  if (smem + 1024 > (size_t) ctx->smem_optin) return P7GPU_OK;
  auto kern = vitp_kernel<C2, TREG>;
  static int occ_dev[64];
  int &occ = occ_dev[ctx->device & 63];
  if (occ <= 0) {
    P7G_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    P7G_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    if (occ < 1) { set_error("vitp_kernel<%d>: zero occupancy", C2); return P7GPU_ESYS; }
    if (!TREG) occ = std::min(occ, 512 / vitp_tm_cols(C2));           // Tensor Memory columns per SM
  }
  long long blocks = (long long) ctx->sm_count * occ;
  long long need = (p.c.n + (NT / 32) - 1) / (NT / 32);
  if (need < blocks) blocks = need;
  if (blocks < 1) blocks = 1;
  { LaunchTimer t(ctx, P7GPU_KERNEL_VIT, stream); kern<<<(unsigned) blocks, NT, smem, stream>>>(p); }
  P7G_CUDA(cudaGetLastError());
  ctx->launches++;
  *launched = true;
  return P7GPU_OK;
}

#define P7G_VITP_C2_CASES(X) \
  X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(16) X(18) X(20) X(24) X(28) X(32) X(40) X(46)

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

static int vitp_tables_c2(int C2)
{
  static const int cs[] = {1,2,3,4,5,6,7,8,9,10,11,12,13,14,16,18,20,24,28,32,40,46};
  for (int c : cs) if (c >= C2) return c;
  return -1;
}
int vit_packed_c2(int M) { return vitp_tables_c2((M + 63) / 64); }

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
  // CTR_VIT_WORK (2 words) .. CTR_VIT_REDO (2) .. CTR_VIT_WORK2 (2) are not contiguous with each other: clear both ranges
  P7G_CUDA(cudaMemsetAsync(ln.counters + CTR_VIT_WORK, 0, 2 * sizeof(unsigned int), ln.stream));

  long long *d_redo = nullptr;
  if (prof->vitp.enabled && !ctx->vit_exact_only) {
    // fast path over every work item; what it cannot decide comes back as (target, slot) pairs for the exact kernel
    P7G_CUDA(cudaMemsetAsync(ln.counters + CTR_VIT_REDO, 0, 4 * sizeof(unsigned int), ln.stream));
    P7G_CUDA(cudaMallocAsync((void **) &d_redo, (size_t) n * 2 * sizeof(long long), ln.stream));
    VitpParams q;
    q.c = p.c; q.rwp = prof->d_rwp; q.twp = prof->d_twp; q.scanw = prof->d_scanw; q.s = prof->vit; q.w = prof->vitp;
    q.redo_list = d_redo; q.redo_count = ln.counters + CTR_VIT_REDO; q.k16 = 0x10000u;
    bool launched = false;
    int rc = P7GPU_OK;
    switch (prof->vitp.C2) {
#define X(CV) case CV: rc = launch_vitp_c<CV>(ctx, ln.stream, q, &launched); break;
      P7G_VITP_C2_CASES(X)
#undef X
    default: break;
    }
    if (rc != P7GPU_OK) { cudaFreeAsync(d_redo, ln.stream); return rc; }
    if (launched) {
      p.c.pairs = d_redo; p.c.pair_count = ln.counters + CTR_VIT_REDO;
      p.c.work_counter = (unsigned long long *) (ln.counters + CTR_VIT_WORK2);
    }
  }
  int rc;
  switch (prof->dpC) {
#define X(CV) case CV: rc = launch_vit_c<CV>(ctx, ln.stream, p); break;
    P7G_DP_C_CASES(X)
#undef X
  default: set_error("vit: unsupported cells-per-lane %d", prof->dpC); rc = P7GPU_EINVAL;
  }
  if (d_redo) cudaFreeAsync(d_redo, ln.stream);
  return rc;
}

}  // namespace p7g
