// bias.cu — sm_100a kernel for the bias filter score of p7_Pipeline() (p7_pipeline.c:697-703):
// p7_bg_FilterScore (HMMER/src/p7_bg.c:472-484) = esl_hmm_Forward on the 2-state filter HMM
// (HMMER/easel/esl_hmm.c:351-409) plus the length term.  It runs on the ~2% of targets that pass the MSV filter.
//
// bias_kernel, one thread per target (throughput form, large survivor lists): the recursion is a 2x2 matrix-vector product per
// residue with a per-row renormalisation and a double-precision log() of the row maximum (float sc[i] = log(max), accumulated in
// float).  bias_warp_kernel, one warp per target (latency form, few survivors): same arithmetic, the logarithms of 32 rows in parallel.  Float operations are issued in the reference's order with explicit
// round-to-nearest intrinsics (no FMA contraction); the L-dependent logf() terms come from the host-built per-L table.
//
// The one libm call is (float) log((double) max).  CUDA's log() and glibc's are both within 1 ulp of the true value but not
// bit-identical, and the float they round to differs only when the true value lies within that distance of a float rounding
// boundary.  The kernel checks exactly that — does rounding d*(1 +- 2^-48) give the same float as rounding d? — and reports
// the (about three per 10^7-target pass) targets where it cannot be sure; host_bias_filtersc (p7gpu_host.cu) recomputes those
// with the host libm.  filtersc is therefore the reference's bit for bit by construction, not by probability.
#include "p7gpu_internal.h"

namespace p7g {

struct BiasParams {
  const uint8_t *res; const int64_t *off; const int32_t *len; const int64_t *idx; long long first; long long n;
  const LenConst *lc;
  BiasHmm h;
  float *filtersc;
  double window;                 // relative half-width of the "cannot be sure" test (2^-48; tests widen it)
  int32_t *unsure_list; unsigned int *unsure_count;   // work items whose score the host must recompute
};

// (float) log((double) v); *unsure is set when a 1-ulp difference in the double logarithm could change the float
__device__ __forceinline__ float log_checked(float v, double window, bool *unsure)
{
  const double d = log((double) v);
  const float f = (float) d;
  if ((float) (d * (1.0 + window)) != f || (float) (d * (1.0 - window)) != f) *unsure = true;
  return f;
}

__global__ void __launch_bounds__(128) bias_kernel(const BiasParams p)
{
  const long long j = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p.n) return;
  const long long s = p.idx ? p.idx[j] : p.first + j;
  const int L = p.len[s];
  const uint8_t *x = p.res + p.off[s];                 // x[i-1] = residue i
  const float p1 = __fdiv_rn((float) L, (float) (L + 1));
  const float t00 = p1, t01 = __fsub_rn(1.0f, p1);      // p7_bg_SetLength, p7_bg.c:193-196
  const float t10 = p.h.t[1][0], t11 = p.h.t[1][1], t02 = p.h.t[0][2], t12 = p.h.t[1][2];
  const float len_term_a = __fmul_rn((float) L, p.lc[L].logf_p1), len_term_b = p.lc[L].logf_1mp1;
  float logsc = 0.0f;
  bool unsure = false;
  if (L == 0) {                                        // esl_hmm.c:362-368 (not reached by the pipeline, :682)
    p.filtersc[j] = __fadd_rn(__fadd_rn((float) log((double) p.h.pi[2]), len_term_a), len_term_b);
    return;
  }
  float dp0 = __fmul_rn(p.h.eo[x[0]][0], p.h.pi[0]);
  float dp1 = __fmul_rn(p.h.eo[x[0]][1], p.h.pi[1]);
  float mx = 0.0f; if (dp0 > mx) mx = dp0; if (dp1 > mx) mx = dp1;
  dp0 = __fdiv_rn(dp0, mx); dp1 = __fdiv_rn(dp1, mx);
  logsc = __fadd_rn(logsc, log_checked(mx, p.window, &unsure));
  for (int i = 2; i <= L; i++) {
    const int r = x[i - 1];
    float n0 = __fadd_rn(0.0f, __fmul_rn(dp0, t00)); n0 = __fadd_rn(n0, __fmul_rn(dp1, t10)); n0 = __fmul_rn(n0, p.h.eo[r][0]);
    float n1 = __fadd_rn(0.0f, __fmul_rn(dp0, t01)); n1 = __fadd_rn(n1, __fmul_rn(dp1, t11)); n1 = __fmul_rn(n1, p.h.eo[r][1]);
    mx = 0.0f; if (n0 > mx) mx = n0; if (n1 > mx) mx = n1;
    dp0 = __fdiv_rn(n0, mx); dp1 = __fdiv_rn(n1, mx);
    logsc = __fadd_rn(logsc, log_checked(mx, p.window, &unsure));
  }
  float last = __fadd_rn(0.0f, __fmul_rn(dp0, t02)); last = __fadd_rn(last, __fmul_rn(dp1, t12));
  logsc = __fadd_rn(logsc, log_checked(last, p.window, &unsure));
  p.filtersc[j] = __fadd_rn(__fadd_rn(logsc, len_term_a), len_term_b);          // p7_bg.c:481
  if (unsure && p.unsure_list) p.unsure_list[atomicAdd(p.unsure_count, 1u)] = (int32_t) j;
}

// Latency form, one warp per target, for the few-survivor launches (hmmscan: a handful of queries per model, where the one-thread
// kernel's ~180 us serial chain was the longest step of a model's pipeline).  The recursion itself stays serial — every lane
// carries it redundantly, 32 rows at a time, lane k keeping row k's maximum — but the 32 double-precision logarithms of a tile,
// which are most of the chain, are taken in parallel, one per lane, and then added to the score in row order.  Every float
// operation, its operands and the order of the additions are those of bias_kernel (and of esl_hmm_Forward), so the result is the
// same bit for bit (tests/test_gpu_pipeline.py::test_bias_warp_kernel_equals_thread_kernel); the certainty test is per logarithm.
__global__ void __launch_bounds__(128) bias_warp_kernel(const BiasParams p)
{
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const long long j = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= p.n) return;                                // the whole warp leaves together
  const long long s = p.idx ? p.idx[j] : p.first + j;
  const int L = p.len[s];
  const uint8_t *x = p.res + p.off[s];
  const float p1 = __fdiv_rn((float) L, (float) (L + 1));
  const float t00 = p1, t01 = __fsub_rn(1.0f, p1);
  const float t10 = p.h.t[1][0], t11 = p.h.t[1][1], t02 = p.h.t[0][2], t12 = p.h.t[1][2];
  const float len_term_a = __fmul_rn((float) L, p.lc[L].logf_p1), len_term_b = p.lc[L].logf_1mp1;
  if (L == 0) {
    if (lane == 0) p.filtersc[j] = __fadd_rn(__fadd_rn((float) log((double) p.h.pi[2]), len_term_a), len_term_b);
    return;
  }
  float logsc = 0.0f, dp0 = 0.0f, dp1 = 0.0f;
  bool unsure = false;
  for (int base = 0; base < L; base += 32) {
    const int nrow = min(32, L - base);
    const int r = lane < nrow ? (int) x[base + lane] : 0;
    const float e0l = p.h.eo[r][0], e1l = p.h.eo[r][1];
    float mymx = 1.0f;
    for (int k = 0; k < nrow; k++) {
      const float e0 = __shfl_sync(full, e0l, k), e1 = __shfl_sync(full, e1l, k);
      float n0, n1;
      if (base + k == 0) {
        n0 = __fmul_rn(e0, p.h.pi[0]);
        n1 = __fmul_rn(e1, p.h.pi[1]);
      } else {
        n0 = __fadd_rn(0.0f, __fmul_rn(dp0, t00)); n0 = __fadd_rn(n0, __fmul_rn(dp1, t10)); n0 = __fmul_rn(n0, e0);
        n1 = __fadd_rn(0.0f, __fmul_rn(dp0, t01)); n1 = __fadd_rn(n1, __fmul_rn(dp1, t11)); n1 = __fmul_rn(n1, e1);
      }
      float mx = 0.0f; if (n0 > mx) mx = n0; if (n1 > mx) mx = n1;
      dp0 = __fdiv_rn(n0, mx); dp1 = __fdiv_rn(n1, mx);
      if (k == lane) mymx = mx;
    }
    float f = 0.0f;
    if (lane < nrow) f = log_checked(mymx, p.window, &unsure);
    for (int k = 0; k < nrow; k++) logsc = __fadd_rn(logsc, __shfl_sync(full, f, k));
  }
  float last = __fadd_rn(0.0f, __fmul_rn(dp0, t02)); last = __fadd_rn(last, __fmul_rn(dp1, t12));
  logsc = __fadd_rn(logsc, log_checked(last, p.window, &unsure));
  unsure = __any_sync(full, unsure);
  if (lane == 0) {
    p.filtersc[j] = __fadd_rn(__fadd_rn(logsc, len_term_a), len_term_b);
    if (unsure && p.unsure_list) p.unsure_list[atomicAdd(p.unsure_count, 1u)] = (int32_t) j;
  }
}

int launch_bias(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
                float *d_filtersc, int32_t *d_unsure_list, unsigned int *d_unsure_count)
{
  if (n <= 0) return P7GPU_OK;
  if (!prof->have_bias) { set_error("bias filter: the profile has no filter HMM (p7gpu_profile_set_biasfilter)"); return P7GPU_EINVAL; }
  P7G_CHECK(ensure_len_table(prof, db.max_length));
  const Lane &ln = ctx->lanes[lane];
  BiasParams p;
  p.res = db.res; p.off = db.off; p.len = db.len; p.idx = d_idx; p.first = db.first; p.n = n;
  p.lc = prof->d_len; p.h = prof->bias; p.filtersc = d_filtersc;
  p.window = ctx->bias_window > 0 ? (double) ctx->bias_window : 3.552713678800501e-15;   // 2^-48
  p.unsure_list = d_unsure_list; p.unsure_count = d_unsure_count;
  if (d_unsure_count) P7G_CUDA(cudaMemsetAsync(d_unsure_count, 0, sizeof(unsigned int), ln.stream));
  // up to one warp per target for every warp slot of the device the launch is latency-bound: the warp form (option "bias_warp":
  // 0 = never, 1 = automatic, 2 = always)
  const bool warp_form = ctx->bias_warp >= 2 || (ctx->bias_warp == 1 && n <= (int64_t) ctx->sm_count * 32);
  {
    LaunchTimer t(ctx, P7GPU_KERNEL_BIAS, ln.stream);
    if (warp_form) bias_warp_kernel<<<(unsigned) ((n * 32 + 127) / 128), 128, 0, ln.stream>>>(p);
    else           bias_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, ln.stream>>>(p);
  }
  P7G_CUDA(cudaGetLastError());
  ctx->launches++;
  return P7GPU_OK;
}

}  // namespace p7g
