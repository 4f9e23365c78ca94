// fwdbwd_mw.cuh — one CTA of W warps per target: p7_ForwardParser / p7_BackwardParser
// (HMMER/src/impl_sse/fwdback.c:257-465 and :469-735, do_full = 0) with the model split over W warps (W = 1 up to M = 512).
// These are the production kernels for every model length; fwdbwd.cu keeps the first-generation one-warp kernels as a
// cross-check (ctx option "fb_one_warp").
//
// The one-warp kernels of fwdbwd.cu keep C = ceil(M/32) cells per lane; beyond C = 16 (M > 512) their per-node
// transition vectors no longer fit in registers and the tables no longer fit in shared memory, so every cell costs
// nine table loads (ncu r1f, M=2000: LSU 61%, FP32 pipe 26%/17%, 255 registers).  Here G = 32*W "virtual lanes" own
// C <= 16 consecutive nodes each (node k = g*C + j + 1), all eight transition vectors of a lane stay in registers, and
// the only table traffic is one emission load per cell (L1-resident, coalesced [x][j][g]).  Per row the warps meet at
// two CTA barriers: Forward — (1) the affine D->D maps of the warps are composed in warp order, (2) the row sum xE;
// Backward — (1) the B-state sum xB and the k+1 neighbour across the warp boundary, (2) the D<-D maps.
// Arithmetic inside a lane is that of fwdbwd.cu with FMA (FTZ; bwd_mw.cu by contraction, fwd_mw.cu written out — csrc/Makefile); the reference's rescaling rule and xmx layout are kept.
#pragma once
#include "dp_common.cuh"
#include <math_constants.h>

namespace p7g {

struct FbMwParams {
  DpCommon c;
  const float *rf;      // [NTABROWS][C][G]
  const float *tf;      // [8][C][G]
  FwdScalars s;
  float *xmx;           // output specials of this direction (may be null)
  const float *fwd_xmx; // backward only: forward specials (scale factors)
  const int64_t *xmx_off;
  float *totscale;      // per-target total log scale (may be null)
  int32_t *own;         // backward only: has_own_scales per target (may be null)
};

__device__ __forceinline__ float mw_warp_sum(float v)
{
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

__device__ __forceinline__ void mw_store_specials(float *row, float xE, float xN, float xJ, float xB, float xC, float scale)
{
  float2 *r2 = reinterpret_cast<float2 *>(row);
  r2[0] = make_float2(xE, xN); r2[1] = make_float2(xJ, xB); r2[2] = make_float2(xC, scale);
}

// the CTA's next work item (thread 0 draws it; two barriers so that the slot can be reused)
__device__ __forceinline__ bool mw_next_target(const DpCommon &c, long long *sh, long long *s, long long *out)
{
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long j = atomicAdd(c.work_counter, 1ULL);
    long long ss = -1, oo = -1;
    if (j < (unsigned long long) c.n) {
      ss = c.idx ? c.idx[j] : (c.order ? (long long) c.order[j] : c.first + (long long) j);
      oo = c.idx ? (long long) j : ss;
    }
    sh[0] = ss; sh[1] = oo;
  }
  __syncthreads();
  *s = sh[0]; *out = sh[1];
  return *s >= 0;
}

// (cells per lane C, warps per CTA W) instances; G = 32*W is a compile-time constant so that table loads have immediate offsets
#define P7G_MW_CASES(X) \
  X(2, 1) X(3, 1) X(4, 1) X(5, 1) X(6, 1) X(7, 1) X(8, 1) X(9, 1) X(10, 1) X(11, 1) X(12, 1) X(13, 1) X(14, 1) X(15, 1) X(16, 1) \
  X(5, 2) X(6, 2) X(7, 2) X(8, 2) X(9, 2) X(10, 2) X(11, 2) X(12, 2) X(13, 2) X(14, 2) X(15, 2) X(16, 2) \
  X(9, 4) X(10, 4) X(11, 4) X(12, 4) X(13, 4) X(14, 4) X(15, 4) X(16, 4) \
  X(9, 8) X(10, 8) X(11, 8) X(12, 8) X(13, 8) X(14, 8) \
  X(10, 12) X(11, 12)

int launch_fwd_mw(p7gpu_ctx *ctx, cudaStream_t stream, const p7gpu_profile *prof, const FbMwParams &p);
int launch_bwd_mw(p7gpu_ctx *ctx, cudaStream_t stream, const p7gpu_profile *prof, const FbMwParams &p);

}  // namespace p7g
