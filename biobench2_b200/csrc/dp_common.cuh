// dp_common.cuh — shared pieces of the row-synchronous kernels (Viterbi filter, Forward/Backward parsers).
//
// Layout: one warp per target; lane l owns the C consecutive model nodes k = l*C + j + 1, j = 0..C-1
// ("blocked" layout: the k-1 neighbour is a register of the same lane except for j = 0, which comes from
// lane l-1 by one warp shuffle per state per row).  Profile tables are stored lane-transposed,
// tab[row][j][lane], so a warp-wide load of cell j is one conflict-free / coalesced access.
#pragma once
#include "p7gpu_internal.h"

namespace p7g {

constexpr int DP_THREADS = 128;

struct DpCommon {
  const uint8_t *res = nullptr; const int64_t *off = nullptr; const int32_t *len = nullptr;
  const int32_t *order = nullptr; const int64_t *idx = nullptr;
  long long first = 0;
  long long n = 0;
  const LenConst *lc = nullptr;
  float *sc = nullptr; int32_t *status = nullptr;
  unsigned long long *work_counter = nullptr;
  // optional device-side work list of (target, output slot) pairs whose length is read on the device
  // (the packed Viterbi kernel's "redo" list); when set it replaces idx/order/first and n
  const long long *pairs = nullptr; const unsigned int *pair_count = nullptr;
};

// fetch the next work item for this warp; returns false when the queue is empty
__device__ __forceinline__ bool next_target(const DpCommon &p, int lane, long long *s, long long *out)
{
  unsigned long long j = 0;
  if (lane == 0) j = atomicAdd(p.work_counter, 1ULL);
  j = __shfl_sync(0xffffffffu, j, 0);
  if (p.pairs) {
    if (j >= (unsigned long long) *p.pair_count) return false;
    *s = p.pairs[2 * j]; *out = p.pairs[2 * j + 1];
    return true;
  }
  if (j >= (unsigned long long) p.n) return false;
  *s   = p.idx ? p.idx[j] : (p.order ? (long long) p.order[j] : p.first + (long long) j);
  *out = p.idx ? (long long) j : *s;
  return true;
}

// cells per lane: smallest supported C with 32*C >= M
#define P7G_DP_C_CASES(X) \
  X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(16) X(18) X(20) X(24) X(28) X(32) \
  X(40) X(48) X(56) X(64) X(80) X(96) X(112) X(130)

// the first-generation one-warp Forward/Backward kernels (fwdbwd.cu) are a cross-check only (option "fb_one_warp"); beyond
// C = 32 cells per lane (M > 1024) they spilled to 3-5 KB of local memory per thread, so they are not instantiated there —
// the exact-order kernels (fbx.cu) against the bit-exact striped oracle are the stronger check at every M
#define P7G_FB_ONEWARP_C_CASES(X) \
  X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(16) X(18) X(20) X(24) X(28) X(32)

}  // namespace p7g
