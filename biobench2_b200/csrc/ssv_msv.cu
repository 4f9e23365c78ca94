// ssv_msv.cu — sm_100a kernels for the 8-bit MSV stage of p7_Pipeline():
//   ssv_kernel  : p7_SSVFilter  (HMMER/src/impl_sse/ssvfilter.c:833-926), all targets
//   msvj_kernel : the J-state recursion of p7_MSVFilter (impl_sse/msvfilter.c:108-207) for the targets
//                 where SSV answers eslENORESULT (msvfilter.c:105-106)
//
// One warp per target sequence, model states across lanes.  The DP lives in registers as 64*Q running
// diagonals ("slots", 2 x int16 per register, DPX VIADDMNMX.S16x2 / VIMNMX3.S16x2); because the SSV
// recurrence v(i,k) = max(floor, v(i-1,k-1) + gain[x_i][k]) has no row-to-row scalar, a slot simply
// follows its diagonal and wraps through >= 1 pad column (gain = PADGAIN) that resets it to the floor:
// no shuffles, no permutes.  The per-profile gain table is staged once per CTA in shared memory as
// (lo,hi) int16 pairs laid out so that every LDS.32 of a warp is bank-conflict free (lane stride Q odd).
#include "p7gpu_internal.h"
#include "tmem.cuh"
#include <math_constants.h>

namespace p7g {

// Where the SSV gain table lives: 0 = global memory (L1/L2), 1 = all 30 rows in shared memory, 2 = "compact": the 20
// canonical residues + the pad row in shared memory, the rare degenerate codes read from global (warp-uniform branch).
constexpr int SSV_GLOBAL = 0, SSV_SMEM = 1, SSV_COMPACT = 2;
constexpr int SSV_COMPACT_ROWS = 21;
__host__ __device__ constexpr int ssv_lane_stride(int Q)  { return (Q & 1) ? Q : 1; }          // words between the slots of neighbouring lanes
__host__ __device__ constexpr int ssv_slot_stride(int Q)  { return (Q & 1) ? 1 : 32; }         // words between the registers of one lane
constexpr int ssv_mode_for(int Q)    { return (Q & 1) ? (Q <= 27 ? SSV_SMEM : (Q <= 41 ? SSV_COMPACT : SSV_GLOBAL)) : (Q <= 8 ? SSV_SMEM : SSV_COMPACT); }
constexpr int ssv_threads_for(int Q) { return (Q & 1) ? (Q <= 13 ? 256 : 512) : (Q <= 4 ? 256 : 512); }   // big tables leave room for few CTAs per SM
constexpr int MSVJ_THREADS = 128;

struct SsvParams {
  const uint32_t *tab;          // [NTABROWS][rowWords]
  int rowWords, P, M;
  const uint8_t *res; const int64_t *off; const int32_t *len; const int32_t *order; const int64_t *idx;
  long long first;
  long long n;
  const LenConst *lc;
  MsvScalars s;
  float *sc; int32_t *status;
  long long *nores_list;        // pairs (target, output slot), may be null
  unsigned int *nores_count;
  int32_t *pass_list;           // targets passing the MSV P-value test (p7_pipeline.c:691-693), may be null
  unsigned int *pass_count;
  unsigned long long *work_counter;
};

// Word b (bytes 4b..4b+3) of a 16-byte chunk whose first `head` and last `tail` bytes lie outside the target: those bytes
// become the pad residue (an all-PADGAIN table row), so that the chunk body needs no per-row bounds test.
__device__ __forceinline__ uint32_t patch_word(uint32_t w, int b, int head, int tail)
{
  const int nh = min(max(head - 4 * b, 0), 4), nt = min(max(tail - 4 * (3 - b), 0), 4);
  const uint32_t mh = nh >= 4 ? 0xffffffffu : ((1u << (8 * nh)) - 1u);
  const uint32_t mt = nt >= 4 ? 0xffffffffu : ~(0xffffffffu >> (8 * nt));
  const uint32_t m = mh | mt;
  return (w & ~m) | ((uint32_t) PADRES * 0x01010101u & m);
}

// 16 residues per load; the L2::128B hint makes the L2 fetch the whole 128-byte line from HBM once, so that the
// seven later 16-byte loads of the same line (16..112 DP rows later) hit in L2 instead of re-reading DRAM sectors
// (ncu r1f: 7.4 GB read per launch for 3.7 GB of database without the hint).
__device__ __forceinline__ uint4 ldg_res16(const uint8_t *ptr)
{
  uint4 v;
  asm("ld.global.nc.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
  return v;
}

__device__ __forceinline__ int hmax16(uint32_t x)
{
  int lo = (int) (short) (x & 0xffff), hi = (int) (short) (x >> 16);
  return lo > hi ? lo : hi;
}

// The decision tree of p7_SSVFilter after get_xE (ssvfilter.c:884-925), on one lane.
__device__ __forceinline__ int ssv_epilogue(int maxu, int tjb, const MsvScalars &s, float *ret_sc, int *ret_xJ)
{
  *ret_xJ = 0;
  int xE = 128 + maxu; if (xE > 255) xE = 255;
  if (xE >= 255 - s.bias_b) {
    *ret_sc = CUDART_INF_F;
    return (s.base_b - tjb - s.tbm_b < 128) ? P7GPU_ENORESULT : P7GPU_ERANGE;
  }
  unsigned short xE16 = (unsigned short) (xE + s.base_b - tjb - s.tbm_b);
  xE16 = (unsigned short) (xE16 - 128);
  if (xE16 >= 255 - s.bias_b) { *ret_sc = CUDART_INF_F; return P7GPU_ERANGE; }
  unsigned short xJ = (unsigned short) (xE16 - s.tec_b);
  if (xJ > s.base_b) { *ret_sc = 0.0f; return P7GPU_ENORESULT; }
  *ret_xJ = (int) xJ;
  float sc = ((float) ((int) xJ - tjb) - (float) s.base_b);
  sc = __fdiv_rn(sc, s.scale_b);
  sc = __fsub_rn(sc, 3.0f);
  *ret_sc = sc;
  return P7GPU_OK;
}

template <int Q, int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) ssv_kernel(const SsvParams p)
{
  extern __shared__ __align__(16) uint32_t s_tab[];
  constexpr bool SMEM = MODE != SSV_GLOBAL;
  if (MODE == SSV_SMEM) {
    const int nwords = NTABROWS * p.rowWords;
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) s_tab[i] = __ldg(p.tab + i);
    __syncthreads();
  } else if (MODE == SSV_COMPACT) {
    const int nwords = 20 * p.rowWords;                                   // rows 0..19 as they are, the pad row as row 20
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) s_tab[i] = __ldg(p.tab + i);
    for (int i = threadIdx.x; i < p.rowWords; i += blockDim.x) s_tab[nwords + i] = __ldg(p.tab + PADRES * p.rowWords + i);
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int P = p.P, rowWords = p.rowWords;
  const uint32_t rowBytes = 4u * (uint32_t) rowWords;
  const int thresh = 127 - p.s.bias_b;          // u >= thresh  <=>  xE >= 255 - bias

  for (;;) {
    unsigned long long j = 0;
    if (lane == 0) j = atomicAdd(p.work_counter, 1ULL);
    j = __shfl_sync(0xffffffffu, j, 0);
    if (j >= (unsigned long long) p.n) break;
    const long long s   = p.idx ? p.idx[j] : (p.order ? (long long) p.order[j] : p.first + (long long) j);
    const long long out = p.idx ? (long long) j : s;
    const int       L   = p.len[s];
    const long long o   = p.off[s];
    const LenConst  lc  = p.lc[L];
    const int       tjb = lc.tjb_b;

    int status; float sc = 0.0f; int xJ = 0;
    if (tjb + p.s.tbm_b + p.s.tec_b + p.s.bias_b >= 127) {      // ssvfilter.c:884
      status = P7GPU_ENORESULT;
    } else {
      uint32_t v[Q];
#pragma unroll
      for (int q = 0; q < Q; q++) v[q] = 0u;
      uint32_t xE = 0u;
      int c = lane * ssv_lane_stride(Q);
      // rows are counted relative to the target's first residue (32-bit): chunk k covers rows r0 .. r0+15, r0 = 16k - head,
      // where head = bytes between the 16-byte boundary below the target and its first residue (0 in the aligned layout)
      const int head0 = (int) (o & 15LL);
      const uint8_t *ptr = p.res + (o - head0);
      uint4 cur = ldg_res16(ptr);
      int since_check = 0;
#pragma unroll (Q <= 14 ? 2 : 1)
      for (int r0 = -head0; r0 < L; r0 += DIAG_CHUNK) {
        ptr += DIAG_CHUNK;
        uint4 nxt = cur;
        if (r0 + DIAG_CHUNK < L) nxt = ldg_res16(ptr);
        uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
        if (r0 < 0 || r0 + DIAG_CHUNK > L) {
          const int head = -r0, tail = r0 + DIAG_CHUNK - L;
#pragma unroll
          for (int b = 0; b < 4; b++) w[b] = patch_word(w[b], b, head, tail);
        }
        // byte address of this lane's first slot in table row 0 for this chunk; row x adds x*rowBytes (one IMAD),
        // the in-chunk row r and the slot q are immediate offsets of the LDS
        const uint32_t *s_lane = s_tab + c;
        uint32_t pend = 0u; bool have_pend = false;
#pragma unroll
        for (int r = 0; r < DIAG_CHUNK; r++) {
          const uint32_t x = __byte_perm(w[r >> 2], 0u, 0x4440u | (r & 3));
          uint32_t g[Q];
          if (SMEM && (MODE == SSV_SMEM || x < 20u || x == (uint32_t) PADRES)) {
            const uint32_t xr = (MODE == SSV_COMPACT) ? min(x, 20u) : x;
            // one multiply-add per row for the address; the in-chunk row and the slot are immediate offsets of the LDS
            const uint32_t *rowp = reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(s_lane) + xr * rowBytes);
#pragma unroll
            for (int q = 0; q < Q; q++) g[q] = rowp[r + q * ssv_slot_stride(Q)];
          } else {
            const uint32_t *rowp = p.tab + x * rowWords + c + r;
#pragma unroll
            for (int q = 0; q < Q; q++) g[q] = __ldg(rowp + q * ssv_slot_stride(Q));
          }
          // v >= 0 always, so max(v+g, g, 0) == max(v+g, 0): the .RELU form needs no zero operand register
#pragma unroll
          for (int q = 0; q < Q; q++) v[q] = __viaddmax_s16x2_relu(v[q], g[q], g[q]);
          // running maximum, two cells per VIMNMX3; an odd leftover waits for the next row (flag folds at compile time)
#pragma unroll
          for (int q = 0; q < Q; q++) {
            if (have_pend) { xE = __vimax3_s16x2(xE, pend, v[q]); have_pend = false; }
            else           { pend = v[q]; have_pend = true; }
          }
        }
        if (have_pend) { xE = __vmaxs2(xE, pend); have_pend = false; }
        c += DIAG_CHUNK; if (c >= P) c -= P;
        cur = nxt;
        // once any cell reaches 255-bias the outcome is fixed (overflow); leaving early also keeps
        // every int16 lane far from wrap-around (<= thresh + 64*bias < 2^15).
        if (++since_check == 4) {
          since_check = 0;
          if (__any_sync(0xffffffffu, hmax16(xE) >= thresh)) break;
        }
      }
      const int maxu = __reduce_max_sync(0xffffffffu, hmax16(xE));
      status = ssv_epilogue(maxu, tjb, p.s, &sc, &xJ);
    }
    if (lane == 0) {
      p.sc[out] = sc;
      p.status[out] = status;
      if (p.pass_list && L > 0 && (status == P7GPU_ERANGE || (status == P7GPU_OK && xJ >= (int) lc.msv_xj_min)))
        p.pass_list[atomicAdd(p.pass_count, 1u)] = (int32_t) s;
      if (status == P7GPU_ENORESULT && p.nores_list) {
        unsigned int t = atomicAdd(p.nores_count, 1u);
        p.nores_list[2 * (long long) t]     = s;
        p.nores_list[2 * (long long) t + 1] = out;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// ssv_tm_kernel: the same filter with the gain table in Tensor Memory (M <= 1024).
//
// ssv_kernel above is bound by shared-memory bandwidth (Q LDS.32 per DP row per warp against 128 B/clk per SM; the integer
// pipe idles ~16 % of the time).  TMEM is read at 64 B/clk per SM sub-partition (probes/tmem_probe.cu: 237 B/clk per SM), by a
// single instruction per row: tcgen05.ld.32x32b.xN hands every lane N consecutive columns of its own TMEM lane, and the column
// offset (residue * N) is a run-time operand.  A lane can only read its own TMEM lane, so the slots cannot follow their
// diagonals as they do above (that needs a different table column every row); the layout is column-stationary instead:
// lane l, register q, half h holds node k = 1 + h*32*Q + l*Q + q, a row moves every value one node up — a register rename
// inside the lane plus ONE warp rotation for the lane boundary (lane 0 takes node 0 = floor into its low half and lane 31's
// low half, node 32*Q, into its high half: one multiply-add, as in vit.cu).  Per row: 1 TMEM load, 1 SHFL, Q packed add-max,
// Q/2 three-input max.  Arithmetic and every decision are those of ssv_kernel (same gains, same early exit, same epilogue).
// TMEM: 30 table rows x N columns, N = 1/2/4/8/16 >= Q, in each of the four 32-lane quarters (warp w reads quarter w % 4).
struct SsvTmParams {
  const uint32_t *tab;          // [NTABROWS][NX][32]
  const uint8_t *res; const int64_t *off; const int32_t *len; const int32_t *order; const int64_t *idx;
  long long first;
  long long n;
  const LenConst *lc;
  MsvScalars s;
  float *sc; int32_t *status;
  long long *nores_list; unsigned int *nores_count;
  int32_t *pass_list; unsigned int *pass_count;
  unsigned long long *work_counter;
};

#ifndef P7G_SSV_TM16_THREADS
#define P7G_SSV_TM16_THREADS 768
#endif
constexpr int SSV_TM16_THREADS = P7G_SSV_TM16_THREADS;   // 513 <= M <= 1024: the table takes all 512 TMEM columns, one CTA per SM

// one residue, zero-extended by the load itself (the compiler masks a uint8_t again after widening it)
__device__ __forceinline__ uint32_t ldg_res1(const uint8_t *ptr)
{
  uint32_t x;
  asm("ld.global.nc.u8 %0, [%1];" : "=r"(x) : "l"(ptr));
  return x;
}

__host__ __device__ constexpr int ssv_tm_nx(int Q)      { return Q <= 1 ? 1 : Q <= 2 ? 2 : Q <= 4 ? 4 : Q <= 8 ? 8 : 16; }
__host__ __device__ constexpr int ssv_tm_cols(int NX)   { return NX == 16 ? 512 : NX == 8 ? 256 : NX == 4 ? 128 : NX == 2 ? 64 : 32; }      // power of two >= 30*NX
__host__ __device__ constexpr int ssv_tm_threads(int NX) { return NX == 16 ? SSV_TM16_THREADS : NX == 8 ? 512 : 256; }
__host__ __device__ constexpr int ssv_tm_ctas(int NX)   { return NX == 16 ? 1 : NX == 8 ? 2 : 4; }                         // per SM: TMEM columns and registers allow this many

template <int Q, int NX, int THREADS, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS) ssv_tm_kernel(const SsvTmParams p)
{
  constexpr int NCOLS = ssv_tm_cols(NX);
  __shared__ uint32_t s_slot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "l"((uint64_t) __cvta_generic_to_shared(&s_slot)), "n"(NCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = s_slot;
  // (through a REDUX so that the compiler knows the address is warp-uniform and keeps the per-row address arithmetic on the
  //  uniform datapath)
  const uint32_t mine = __reduce_or_sync(0xffffffffu, tbase + ((uint32_t) ((warp & 3) * 32) << 16));
  if (warp < 4) {                                                          // one warp per quarter writes the table
    for (int x = 0; x < NTABROWS; x++) {
      uint32_t t[NX];
#pragma unroll
      for (int q = 0; q < NX; q++) t[q] = __ldg(p.tab + ((size_t) x * NX + q) * 32 + lane);
      Tmem<NX>::st(mine + x * NX, t);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");

  const int prevl = (lane + 31) & 31;
  uint32_t rmul = lane == 0 ? 0x10000u : 1u;
  asm("" : "+r"(rmul));                                                    // keep the rotation's fix-up a multiply (FMA pipe)
  const int thresh = 127 - p.s.bias_b;          // u >= thresh  <=>  xE >= 255 - bias

  for (;;) {
    unsigned long long j = 0;
    if (lane == 0) j = atomicAdd(p.work_counter, 1ULL);
    j = __shfl_sync(0xffffffffu, j, 0);
    if (j >= (unsigned long long) p.n) break;
    const long long s   = p.idx ? p.idx[j] : (p.order ? (long long) p.order[j] : p.first + (long long) j);
    const long long out = p.idx ? (long long) j : s;
    const int       L   = p.len[s];
    const long long o   = p.off[s];
    const LenConst  lc  = p.lc[L];
    const int       tjb = lc.tjb_b;

    int status; float sc = 0.0f; int xJ = 0;
    if (tjb + p.s.tbm_b + p.s.tec_b + p.s.bias_b >= 127) {      // ssvfilter.c:884
      status = P7GPU_ENORESULT;
    } else {
      uint32_t v[Q];
#pragma unroll
      for (int q = 0; q < Q; q++) v[q] = 0u;
      uint32_t xE = 0u;
      // Residues arrive as byte loads (one broadcast LDG.U8 per row on the otherwise idle load pipe) instead of 16-byte vectors
      // taken apart with PRMT: the integer pipe is what binds this kernel, and the rows then need no alignment or patching.
      const uint8_t *seq = p.res + o;
      auto row = [&](const uint32_t (&g)[NX]) {
        const uint32_t carry = __shfl_sync(0xffffffffu, v[Q - 1], prevl) * rmul;
        // v >= 0 always, so max(v+g, g, 0) == max(v+g, 0): the .RELU form needs no zero operand register
#pragma unroll
        for (int q = Q - 1; q >= 1; q--) v[q] = __viaddmax_s16x2_relu(v[q - 1], g[q], g[q]);
        v[0] = __viaddmax_s16x2_relu(carry, g[0], g[0]);
      };
      int i = 0, since_check = 0;
      bool over = false;
      // rows up to the first 4-byte boundary of the residue array, one at a time (none in the 16-byte-aligned layout)
      for (; i < L && (reinterpret_cast<uintptr_t>(seq + i) & 3); i++) {
        uint32_t g[NX];
        Tmem<NX>::ld(mine + ldg_res1(seq + i) * NX, g);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        row(g);
#pragma unroll
        for (int q = 0; q < Q; q++) xE = __vmaxs2(xE, v[q]);
      }
      for (; i + DIAG_CHUNK <= L; i += DIAG_CHUNK) {
        // Four residues per 32-bit load, made warp-uniform by a REDUX (all lanes hold the same word): the byte extraction and the
        // table address then run on the uniform datapath, and the integer pipe — what binds this kernel — loses the per-row R2UR
        uint32_t xs[DIAG_CHUNK];
#pragma unroll
        for (int r4 = 0; r4 < DIAG_CHUNK / 4; r4++) {
          const uint32_t w = __reduce_or_sync(0xffffffffu, __ldg(reinterpret_cast<const uint32_t *>(seq + i) + r4));
          xs[4 * r4] = w & 0xffu; xs[4 * r4 + 1] = (w >> 8) & 0xffu; xs[4 * r4 + 2] = (w >> 16) & 0xffu; xs[4 * r4 + 3] = w >> 24;
        }
        // the table row of DP row r+1 is fetched while row r is computed; tcgen05.wait::ld only orders the loads already issued
        uint32_t g[NX], gn[NX];
        Tmem<NX>::ld(mine + xs[0] * NX, g);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t pend = 0u; bool have_pend = false;
#pragma unroll
        for (int r = 0; r < DIAG_CHUNK; r++) {
          if (r + 1 < DIAG_CHUNK) Tmem<NX>::ld(mine + xs[r + 1] * NX, gn);
          row(g);
          // running maximum, two cells per VIMNMX3; an odd leftover waits for the next row (flag folds at compile time)
#pragma unroll
          for (int q = 0; q < Q; q++) {
            if (have_pend) { xE = __vimax3_s16x2(xE, pend, v[q]); have_pend = false; }
            else           { pend = v[q]; have_pend = true; }
          }
          if (r + 1 < DIAG_CHUNK) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < NX; q++) g[q] = gn[q];
          }
        }
        if (have_pend) { xE = __vmaxs2(xE, pend); have_pend = false; }
        // once any cell reaches 255-bias the outcome is fixed (overflow); leaving early also keeps
        // every int16 lane far from wrap-around (<= thresh + 64*bias < 2^15).
        if (++since_check == 4) {
          since_check = 0;
          if (__any_sync(0xffffffffu, hmax16(xE) >= thresh)) { over = true; break; }
        }
      }
      if (!over)
        for (; i < L; i++) {                                     // the last L mod 16 rows, one at a time
          uint32_t g[NX];
          Tmem<NX>::ld(mine + ldg_res1(seq + i) * NX, g);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          row(g);
#pragma unroll
          for (int q = 0; q < Q; q++) xE = __vmaxs2(xE, v[q]);
        }
      const int maxu = __reduce_max_sync(0xffffffffu, hmax16(xE));
      status = ssv_epilogue(maxu, tjb, p.s, &sc, &xJ);
    }
    if (lane == 0) {
      p.sc[out] = sc;
      p.status[out] = status;
      if (p.pass_list && L > 0 && (status == P7GPU_ERANGE || (status == P7GPU_OK && xJ >= (int) lc.msv_xj_min)))
        p.pass_list[atomicAdd(p.pass_count, 1u)] = (int32_t) s;
      if (status == P7GPU_ENORESULT && p.nores_list) {
        unsigned int t = atomicAdd(p.nores_count, 1u);
        p.nores_list[2 * (long long) t]     = s;
        p.nores_list[2 * (long long) t + 1] = out;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "n"(NCOLS));
}

template <int Q>
static int launch_ssv_tm_q(p7gpu_ctx *ctx, cudaStream_t stream, const SsvTmParams &p)
{
  constexpr int NX = ssv_tm_nx(Q), THREADS = ssv_tm_threads(NX), CTAS = ssv_tm_ctas(NX);
  auto kern = ssv_tm_kernel<Q, NX, THREADS, CTAS>;
  // the grid never asks an SM for more CTAs than its 512 TMEM columns serve: a CTA beyond that would sit in tcgen05.alloc
  // until another one of this persistent grid has drained the queue
  long long blocks = (long long) ctx->sm_count * CTAS;
  long long need_blocks = (p.n + (THREADS / 32) - 1) / (THREADS / 32);
  if (need_blocks < blocks) blocks = need_blocks;
  if (blocks < 1) blocks = 1;
  { LaunchTimer t(ctx, P7GPU_KERNEL_SSV, stream); kern<<<(unsigned) blocks, THREADS, 0, stream>>>(p); }
  P7G_CUDA(cudaGetLastError());
  ctx->launches++;
  return P7GPU_OK;
}

// ------------------------------------------------------------------------------------------------
struct MsvjParams {
  const uint32_t *tab; int rowWords, P, M;
  const uint8_t *res; const int64_t *off; const int32_t *len;
  const long long *list; const unsigned int *count; long long fixed_n;   // list of (target, out) pairs
  const LenConst *lc; MsvScalars s;
  float *sc; int32_t *status;
  int32_t *pass_list; unsigned int *pass_count;
  unsigned long long *work_counter;
};

template <int Q>
__global__ void __launch_bounds__(MSVJ_THREADS) msvj_kernel(const MsvjParams p)
{
  const int lane = threadIdx.x & 31;
  const int P = p.P, rowWords = p.rowWords;
  const long long n = p.count ? (long long) *p.count : p.fixed_n;
  const uint32_t bias2 = (uint32_t) p.s.bias_b * 0x00010001u;
  const uint32_t c255  = 255u * 0x00010001u;
  for (;;) {
    unsigned long long j = 0;
    if (lane == 0) j = atomicAdd(p.work_counter, 1ULL);
    j = __shfl_sync(0xffffffffu, j, 0);
    if ((long long) j >= n) break;
    const long long s = p.list[2 * j], out = p.list[2 * j + 1];
    const int L = p.len[s];
    const long long o = p.off[s];
    const LenConst lc = p.lc[L];
    const int tjb = lc.tjb_b;
    const int tjbm = (int) (unsigned char) ((int) (signed char) tjb + (int) (signed char) p.s.tbm_b);   // msvfilter.c:118
    uint32_t v[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) v[q] = 0u;
    int xJ = 0;
    int xB = p.s.base_b - tjbm; if (xB < 0) xB = 0;
    int c = lane * ssv_lane_stride(Q);
    int status = P7GPU_OK;
    for (int i = 0; i < L; i++) {
      const uint32_t x = __ldg(p.res + o + i);
      const uint32_t *rowp = p.tab + x * rowWords + c;
      const uint32_t xB2 = (uint32_t) xB * 0x00010001u;
      uint32_t xEv = 0u;
#pragma unroll
      for (int q = 0; q < Q; q++) {
        uint32_t a = __vmaxs2(v[q], xB2);                      // max(M(i-1,k-1), xB)
        a = __viaddmin_s16x2(a, bias2, c255);                  // adds_epu8(., bias)
        { const uint32_t g = __ldg(rowp + q * ssv_slot_stride(Q)); v[q] = __viaddmax_s16x2_relu(a, g, g); }   // subs_epu8(., rbv); a >= 0
        xEv = __vmaxs2(xEv, v[q]);
      }
      c += 1; if (c >= P) c -= P;
      int xE = __reduce_max_sync(0xffffffffu, hmax16(xEv));
      if (xE + p.s.bias_b >= 255) { status = P7GPU_ERANGE; break; }       // msvfilter.c:157-182
      xE -= p.s.tec_b; if (xE < 0) xE = 0;
      if (xE > xJ) xJ = xE;
      xB = (p.s.base_b > xJ ? p.s.base_b : xJ) - tjbm; if (xB < 0) xB = 0;
    }
    if (lane == 0) {
      float sc;
      if (status == P7GPU_ERANGE) sc = CUDART_INF_F;
      else {
        sc = ((float) (xJ - tjb) - (float) p.s.base_b);
        sc = __fdiv_rn(sc, p.s.scale_b);
        sc = __fsub_rn(sc, 3.0f);
      }
      p.sc[out] = sc;
      p.status[out] = status;
      if (p.pass_list && L > 0 && (status == P7GPU_ERANGE || xJ >= (int) lc.msv_xj_min))
        p.pass_list[atomicAdd(p.pass_count, 1u)] = (int32_t) s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
template <int Q, int MODE, int THREADS>
static int launch_ssv_q(p7gpu_ctx *ctx, cudaStream_t stream, const SsvParams &p, size_t smem)
{
  auto kern = ssv_kernel<Q, MODE, THREADS>;
  static int occ_dev[64];                         // per template instance (one geometry), per device; 0 = unknown
  int &occ_cache = occ_dev[ctx->device & 63];
  if (occ_cache <= 0) {
    if (smem) P7G_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int occ = 0;
    P7G_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
    if (occ < 1) { set_error("ssv_kernel<%d,%d,%d>: zero occupancy (smem %zu)", Q, MODE, THREADS, smem); return P7GPU_ESYS; }
    occ_cache = occ;
  }
  long long blocks = (long long) ctx->sm_count * occ_cache;
  long long need_blocks = (p.n + (THREADS / 32) - 1) / (THREADS / 32);
  if (need_blocks < blocks) blocks = need_blocks;
  if (blocks < 1) blocks = 1;
  { LaunchTimer t(ctx, P7GPU_KERNEL_SSV, stream); kern<<<(unsigned) blocks, THREADS, smem, stream>>>(p); }
  P7G_CUDA(cudaGetLastError());
  ctx->launches++;
  return P7GPU_OK;
}

// the preferred placement for this Q when its shared-memory footprint fits the device, else the global-memory form
template <int Q>
static int launch_ssv_pick(p7gpu_ctx *ctx, cudaStream_t stream, const SsvParams &p)
{
  constexpr int MODE = ssv_mode_for(Q), THREADS = ssv_threads_for(Q);
  const size_t rows = MODE == SSV_SMEM ? NTABROWS : (MODE == SSV_COMPACT ? SSV_COMPACT_ROWS : 0);
  const size_t smem = rows * (size_t) p.rowWords * 4;
  if (MODE != SSV_GLOBAL && smem + 1024 <= (size_t) ctx->smem_optin) return launch_ssv_q<Q, MODE, THREADS>(ctx, stream, p, smem);
  return launch_ssv_q<Q, SSV_GLOBAL, 256>(ctx, stream, p, 0);
}

#define P7G_ODD_Q_CASES(X) \
  X(2) X(4) X(6) X(8) X(10) X(12) X(14) \
  X(1) X(3) X(5) X(7) X(9) X(11) X(13) X(15) X(17) X(19) X(21) X(23) X(25) X(27) X(29) X(31) X(33) \
  X(35) X(37) X(39) X(41) X(43) X(45) X(47) X(49) X(51) X(53) X(55) X(57) X(59) X(61) X(63) X(65)

int launch_ssv(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
               float *d_sc, int32_t *d_status, int64_t *d_noresult_list, unsigned int *d_noresult_count,
               int32_t *d_pass_list, unsigned int *d_pass_count)
{
  if (n <= 0) return P7GPU_OK;
  P7G_CHECK(ensure_len_table(prof, db.max_length));
  const Lane &ln = ctx->lanes[lane];
  SsvParams p;
  p.tab = prof->d_ssv_tab; p.rowWords = prof->geom.rowWords; p.P = prof->geom.P; p.M = prof->M;
  p.res = db.res; p.off = db.off; p.len = db.len; p.order = d_idx ? nullptr : db.order; p.idx = d_idx; p.first = db.first;
  p.n = n; p.lc = prof->d_len; p.s = prof->msv; p.sc = d_sc; p.status = d_status;
  p.nores_list = (long long *) d_noresult_list; p.nores_count = d_noresult_count;
  p.pass_list = d_pass_list; p.pass_count = d_pass_count;
  p.work_counter = (unsigned long long *) (ln.counters + CTR_SSV_WORK);
  // work counter (2 words) + the lane's eslENORESULT count (CTR_NORES_COUNT = CTR_SSV_WORK + 2)
  P7G_CUDA(cudaMemsetAsync(ln.counters + CTR_SSV_WORK, 0, 4 * sizeof(unsigned int), ln.stream));
  // 193 <= M <= 1024 (or option ssv_tmem = 2: any M <= 1024): gain table in Tensor Memory.  Below Q = 4 the per-row fixed cost
  // (rotation, table address, LDTM) outweighs the saved LDS: measured 8.05 against 9.2 TCUPS at M = 100
  if (prof->d_ssv_tm && ctx->ssv_tmem && (prof->ssv_tm_Q >= 4 || ctx->ssv_tmem >= 2)) {
    SsvTmParams t;
    t.tab = prof->d_ssv_tm; t.res = p.res; t.off = p.off; t.len = p.len; t.order = p.order; t.idx = p.idx; t.first = p.first; t.n = p.n;
    t.lc = p.lc; t.s = p.s; t.sc = p.sc; t.status = p.status; t.nores_list = p.nores_list; t.nores_count = p.nores_count;
    t.pass_list = p.pass_list; t.pass_count = p.pass_count; t.work_counter = p.work_counter;
    switch (prof->ssv_tm_Q) {
#define X(QV) case QV: return launch_ssv_tm_q<QV>(ctx, ln.stream, t);
      X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)
#undef X
    default: break;
    }
  }
  switch (prof->geom.Q) {
#define X(QV) case QV: return launch_ssv_pick<QV>(ctx, ln.stream, p);
    P7G_ODD_Q_CASES(X)
#undef X
  default: set_error("SSV: unsupported Q=%d (M=%d too large; max M is %d)", prof->geom.Q, prof->M, 64 * MAX_Q - 1); return P7GPU_EINVAL;
  }
}

template <int Q>
static int launch_msvj_q(p7gpu_ctx *ctx, cudaStream_t stream, const MsvjParams &p, long long max_n)
{
  long long blocks = (max_n + (MSVJ_THREADS / 32) - 1) / (MSVJ_THREADS / 32);
  long long cap = (long long) ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  { LaunchTimer t(ctx, P7GPU_KERNEL_MSVJ, stream); msvj_kernel<Q><<<(unsigned) blocks, MSVJ_THREADS, 0, stream>>>(p); }
  P7G_CUDA(cudaGetLastError());
  ctx->launches++;
  return P7GPU_OK;
}

// d_list holds (target, output slot) pairs.  If d_count != nullptr the number of pairs is read on the device.
int launch_msv_j(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_list,
                 const unsigned int *d_count, int64_t max_n, float *d_sc, int32_t *d_status,
                 int32_t *d_pass_list, unsigned int *d_pass_count)
{
  if (max_n <= 0) return P7GPU_OK;
  P7G_CHECK(ensure_len_table(prof, db.max_length));
  const Lane &ln = ctx->lanes[lane];
  MsvjParams p;
  p.tab = prof->d_msv_tab; p.rowWords = prof->geom.rowWords; p.P = prof->geom.P; p.M = prof->M;
  p.res = db.res; p.off = db.off; p.len = db.len;
  p.list = (const long long *) d_list; p.count = d_count; p.fixed_n = max_n;
  p.lc = prof->d_len; p.s = prof->msv; p.sc = d_sc; p.status = d_status;
  p.pass_list = d_pass_list; p.pass_count = d_pass_count;
  p.work_counter = (unsigned long long *) (ln.counters + CTR_MSVJ_WORK);
  P7G_CUDA(cudaMemsetAsync(ln.counters + CTR_MSVJ_WORK, 0, 2 * sizeof(unsigned int), ln.stream));
  switch (prof->geom.Q) {
#define X(QV) case QV: return launch_msvj_q<QV>(ctx, ln.stream, p, max_n);
    P7G_ODD_Q_CASES(X)
#undef X
  default: set_error("MSV: unsupported Q=%d", prof->geom.Q); return P7GPU_EINVAL;
  }
}

// ------------------------------------------------------------------------------------------------
// ssv_scan_kernel: hmmscan's MSV stage for a whole block of models at once.
//
// hmmscan streams thousands of models past a few query sequences (hmmscan.c:1176-1198); one launch set per model is launch-
// bound there.  Here the work item is a (model, query) pair and the gain table is picked per item: the tables of a block of
// models lie side by side in global memory (L1/L2-resident: a warp re-reads ~1 KB of its model's table per DP row).  One
// kernel serves every model length, because the SSV recurrence has no coupling between diagonals: a warp walks the query once
// per group of 256 diagonals ("pass": 4 registers x 32 lanes x 2 halves) and keeps one running maximum over all passes.
// The scan table of a model is laid out for that (scan.cu): P = 256 * npass >= M + 1 columns, word w of a row holds the gains
// of columns (w mod P, (w + 128) mod P), rowWords = P + 96 so that a lane's four words c, c+32, c+64, c+96 never wrap.
// Arithmetic, early exit and the decision tree are those of ssv_kernel (same gains, ssv_epilogue); the output is, per model,
// the number of queries that pass the MSV threshold or need the J-state recursion (eslENORESULT): a model with none is one
// for which p7gpu_pipeline_run(stop_after_msv) would report no survivor, so hmmscan never has to upload it.
constexpr int SCAN_QB = 4, SCAN_THREADS = 128;

__global__ void __launch_bounds__(SCAN_THREADS) ssv_scan_kernel(const ScanParams p)
{
  const int lane = threadIdx.x & 31;
  for (;;) {
    unsigned long long j = 0;
    if (lane == 0) j = atomicAdd(p.work_counter, 1ULL);
    j = __shfl_sync(0xffffffffu, j, 0);
    if (j >= (unsigned long long) p.nitems) break;
    const long long m = (long long) (j / (unsigned long long) p.nq), s = (long long) (j - (unsigned long long) m * (unsigned long long) p.nq);
    const ScanModel md = p.models[m];
    const int       L   = p.len[s];
    const long long o   = p.off[s];
    const uint32_t  lcw = __ldg(p.lc + (size_t) m * p.nlens + p.lenidx[s]);
    const int       tjb = (int) (lcw & 0xffu), xjmin = (int) (lcw >> 16);
    int status; float sc = 0.0f; int xJ = 0;
    if (tjb + md.s.tbm_b + md.s.tec_b + md.s.bias_b >= 127) {    // ssvfilter.c:884
      status = P7GPU_ENORESULT;
    } else {
      const int thresh = 127 - md.s.bias_b;
      const uint32_t *tab = p.tabs + md.tab_off;
      const uint8_t *seq = p.res + o;
      uint32_t xE = 0u;
      bool over = false;
      for (int pass = 0; pass < md.npass && !over; pass++) {
        uint32_t v[SCAN_QB];
#pragma unroll
        for (int q = 0; q < SCAN_QB; q++) v[q] = 0u;
        int c = pass * (64 * SCAN_QB) + lane;
        for (int i = 0; i < L; i++) {
          const uint32_t x = __ldg(seq + i);
          const uint32_t *rowp = tab + x * md.rowWords + c;
#pragma unroll
          for (int q = 0; q < SCAN_QB; q++) {
            const uint32_t g = __ldg(rowp + 32 * q);
            v[q] = __viaddmax_s16x2_relu(v[q], g, g);
            xE = __vmaxs2(xE, v[q]);
          }
          c += 1; if (c >= md.P) c -= md.P;
          // as in ssv_kernel: once a cell reaches 255-bias the outcome is fixed, and leaving keeps the int16 lanes from wrapping
          if ((i & 63) == 63 && __any_sync(0xffffffffu, hmax16(xE) >= thresh)) { over = true; break; }
        }
      }
      const int maxu = __reduce_max_sync(0xffffffffu, hmax16(xE));
      status = ssv_epilogue(maxu, tjb, md.s, &sc, &xJ);
    }
    if (lane == 0 && L > 0 && (status != P7GPU_OK || xJ >= xjmin)) atomicAdd(p.npossible + m, 1);
  }
}

int launch_ssv_scan(p7gpu_ctx *ctx, cudaStream_t stream, const ScanParams &p)
{
  if (p.nitems <= 0) return P7GPU_OK;
  long long blocks = (p.nitems + (SCAN_THREADS / 32) - 1) / (SCAN_THREADS / 32);
  const long long cap = (long long) ctx->sm_count * 16;
  if (blocks > cap) blocks = cap;
  { LaunchTimer t(ctx, P7GPU_KERNEL_SSV, stream); ssv_scan_kernel<<<(unsigned) blocks, SCAN_THREADS, 0, stream>>>(p); }
  P7G_CUDA(cudaGetLastError());
  ctx->launches++;
  return P7GPU_OK;
}

}  // namespace p7g
