// decode.cu — the region trigger of the domain-definition step, on the device.
//
// p7_domaindef_ByPosteriorHeuristics() first calls p7_DomainDecoding() (HMMER/src/impl_sse/decoding.c:166-200) to turn the
// Forward/Backward specials into mocc[i] = 1 - P(residue i emitted by N, J or C), then walks i = 1..L and opens a region
// only where mocc[i] >= rt1 = 0.25 (p7_domaindef.c:109,419).  A target on which that never happens has nregions = 0 and
// p7_Pipeline() drops it (p7_pipeline.c:739).  Under --max every target reaches this point, so evaluating the trigger
// here means the (L+1)*6*2 floats of specials travel to the host only for the few targets that can have a domain.
//
// mocc[i] is computed with the reference's own float/double expression order (no FMA), from the same xmx values the host
// would read, so "max_i mocc[i] >= rt1" is exactly the reference's trigger condition.  Targets whose Backward pass used
// its own scale factors (fwdback.c:661), or whose scale product is not finite (decoding.c:196), are always flagged:
// the unchanged host code decides those.
#include "p7gpu_internal.h"
#include <math_constants.h>

namespace p7g {

struct TrigParams {
  const float *fx, *bx; const int64_t *xoff;
  const int64_t *idx; const int32_t *len; const LenConst *lc;
  const int32_t *own;            // Backward's has_own_scales per work item, may be null
  long long n;
  float rt1;                     // p7_domaindef.c:109 minus the slack below
  float *mocc_max; int32_t *flag;
  unsigned int *nflag;           // optional count of flagged items
};

__global__ void __launch_bounds__(256) region_trigger_kernel(const TrigParams p)
{
  const int lane = threadIdx.x & 31;
  const long long j = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= p.n) return;
  const int L = p.len[p.idx ? p.idx[j] : j];
  const float *f = p.fx + p.xoff[j], *b = p.bx + p.xoff[j];
  const float ploop = 1.0f - p.lc[L].pmove;                           // om->xf[N|J|C][LOOP], p7_oprofile.c:1112-1116
  const float sp = (float) (1.0 / (double) b[1]);                     // decoding.c:169: 1.0 / oxb->xmx[p7X_N]
  float mx = -CUDART_INF_F;
  for (int i = 1 + lane; i <= L; i += 32) {
    const float *fr = f + (size_t) (i - 1) * 6, *br = b + (size_t) i * 6;
    float njcp =             __fmul_rn(__fmul_rn(__fmul_rn(fr[1], br[1]), ploop), sp);    // :186-188
    njcp = __fadd_rn(njcp,   __fmul_rn(__fmul_rn(__fmul_rn(fr[2], br[2]), ploop), sp));
    njcp = __fadd_rn(njcp,   __fmul_rn(__fmul_rn(__fmul_rn(fr[4], br[4]), ploop), sp));
    const float mocc = (float) (1.0 - (double) njcp);                 // :189
    mx = fmaxf(mx, mocc);                                             // a NaN row never satisfies mocc >= rt1 either
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  if (lane == 0) {
    const bool undecided = (p.own && p.own[j]) || isinf(sp) || isnan(sp);
    const int fl = (undecided || mx >= p.rt1) ? 1 : 0;
    p.mocc_max[j] = mx;
    p.flag[j] = fl;
    if (fl && p.nflag) atomicAdd(p.nflag, 1u);
  }
}

int launch_region_trigger(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
                          const float *d_fx, const float *d_bx, const int64_t *d_xoff, const int32_t *d_own,
                          float *d_mocc_max, int32_t *d_flag, unsigned int *d_nflag, float slack)
{
  if (n <= 0) return P7GPU_OK;
  P7G_CHECK(ensure_len_table(prof, db.max_length));
  const Lane &ln = ctx->lanes[lane];
  TrigParams p;
  p.fx = d_fx; p.bx = d_bx; p.xoff = d_xoff; p.idx = d_idx; p.len = db.len; p.lc = prof->d_len; p.own = d_own; p.n = n;
  // rt1 = 0.25 (p7_domaindef.c:109).  On exact-order specials mocc is the reference's bit for bit and slack is 0; on the fast
  // kernels' specials (<= 1e-4 relative) a target within 1e-4 of the trigger is flagged, i.e. left to the host to decide
  p.rt1 = 0.25f - slack;
  p.mocc_max = d_mocc_max; p.flag = d_flag; p.nflag = d_nflag;
  const long long blocks = (n * 32 + 255) / 256;
  { LaunchTimer t(ctx, P7GPU_KERNEL_TRIG, ln.stream); region_trigger_kernel<<<(unsigned) blocks, 256, 0, ln.stream>>>(p); }
  P7G_CUDA(cudaGetLastError());
  ctx->launches++;
  return P7GPU_OK;
}

// ------------------------------------------------------------------------------------------------
// gather the specials of selected targets into a dense buffer (one warp per target): src/dst are float offsets
__global__ void __launch_bounds__(256) gather_xmx_kernel(const float *fx, const float *bx, const int64_t *src, const int64_t *dst,
                                                         const int32_t *nfloats, long long n, float *ofx, float *obx)
{
  const int lane = threadIdx.x & 31;
  const long long j = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= n) return;
  const float2 *sf = reinterpret_cast<const float2 *>(fx + src[j]), *sb = reinterpret_cast<const float2 *>(bx + src[j]);
  float2 *df = reinterpret_cast<float2 *>(ofx + dst[j]), *dbb = reinterpret_cast<float2 *>(obx + dst[j]);
  const int n2 = nfloats[j] >> 1;                                     // (L+1)*6 floats: even, and every block starts 8-byte aligned
  for (int t = lane; t < n2; t += 32) { df[t] = sf[t]; dbb[t] = sb[t]; }
}

int launch_gather_xmx(p7gpu_ctx *ctx, int lane, const float *d_fx, const float *d_bx, const int64_t *d_src, const int64_t *d_dst,
                      const int32_t *d_nfloats, int64_t n, float *d_ofx, float *d_obx)
{
  if (n <= 0) return P7GPU_OK;
  const Lane &ln = ctx->lanes[lane];
  const long long blocks = (n * 32 + 255) / 256;
  gather_xmx_kernel<<<(unsigned) blocks, 256, 0, ln.stream>>>(d_fx, d_bx, d_src, d_dst, d_nfloats, n, d_ofx, d_obx);
  P7G_CUDA(cudaGetLastError());
  ctx->launches++;
  return P7GPU_OK;
}


// Forward's SCALE column and xC(L) of each item, 4 bytes per row instead of the 24 of its specials: all the host needs to
// recompute totscale and the Forward score with its own libm (fwdback.c:434,:463).  Block of item j at out + xoff[j]/6:
// [0] = xC(L), [i] = SCALE(i) for i = 1..L.
__global__ void __launch_bounds__(256) gather_scale_kernel(const float *fx, const int64_t *xoff, const int64_t *idx, const int32_t *len,
                                                           long long n, float *out)
{
  const int lane = threadIdx.x & 31;
  const long long j = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= n) return;
  const int L = len[idx[j]];
  const float *src = fx + xoff[j];
  float *dst = out + xoff[j] / 6;
  for (int i = lane; i <= L; i += 32) dst[i] = i == 0 ? src[(size_t) L * 6 + 4] : src[(size_t) i * 6 + 5];
}

int launch_gather_scale(p7gpu_ctx *ctx, int lane, const DbView &db, const int64_t *d_idx, int64_t n, const float *d_fx,
                        const int64_t *d_xoff, float *d_out)
{
  if (n <= 0) return P7GPU_OK;
  const Lane &ln = ctx->lanes[lane];
  const long long blocks = (n * 32 + 255) / 256;
  gather_scale_kernel<<<(unsigned) blocks, 256, 0, ln.stream>>>(d_fx, d_xoff, d_idx, db.len, n, d_out);
  P7G_CUDA(cudaGetLastError());
  ctx->launches++;
  return P7GPU_OK;
}

// ------------------------------------------------------------------------------------------------
// p7_DomainDecoding itself (impl_sse/decoding.c:166-200) for the flagged survivors: btot[i], etot[i], mocc[i], i = 0..L, in the
// reference's own order of float operations — btot and etot are running sums, so one lane chain does the adds while the 32
// lanes compute the terms of 32 rows at a time.  The host then needs 12 bytes per row instead of the 48 of both sets of
// specials: p7_domaindef_ByPosteriorHeuristics reads nothing else of the two parser matrices (p7_domaindef.c:392-407).
// Each block of the output also carries Forward's SCALE column and xC(L), from which the host recomputes totscale and the
// Forward score with its own libm (fwdback.c:434,:463).  Layout per item at dec + doff[j]: btot[L+1], etot[L+1], mocc[L+1],
// fscale[L+1], xC(L), 0.
struct DecParams {
  const float *fx, *bx; const int64_t *xoff, *doff;
  const int64_t *idx; const int32_t *len; const LenConst *lc;
  long long n;
  float *dec;
};

__global__ void __launch_bounds__(128) domain_decoding_kernel(const DecParams p)
{
  const int lane = threadIdx.x & 31;
  const long long j = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= p.n) return;
  const int L = p.len[p.idx ? p.idx[j] : j];
  const float *f = p.fx + p.xoff[j], *b = p.bx + p.xoff[j];
  float *btot = p.dec + p.doff[j], *etot = btot + (L + 1), *mocc = etot + (L + 1), *fsc = mocc + (L + 1);
  const float ploop = 1.0f - p.lc[L].pmove;
  const float sp = (float) (1.0 / (double) b[1]);                     // decoding.c:169 (has_own_scales targets never come here)
  float bt = 0.0f, et = 0.0f;
  if (lane == 0) { btot[0] = 0.0f; etot[0] = 0.0f; mocc[0] = 0.0f; fsc[0] = f[5]; fsc[L + 1] = f[(size_t) L * 6 + 4]; fsc[L + 2] = 0.0f; }
  for (int i0 = 1; i0 <= L; i0 += 32) {
    const int i = i0 + lane;
    float tb = 0.0f, te = 0.0f;
    if (i <= L) {
      const float *fp = f + (size_t) (i - 1) * 6, *fr = f + (size_t) i * 6, *bp = b + (size_t) (i - 1) * 6, *br = b + (size_t) i * 6;
      tb = __fmul_rn(__fmul_rn(__fmul_rn(fp[3], bp[3]), fp[5]), sp);                        // :177-178
      te = __fmul_rn(__fmul_rn(__fmul_rn(fr[0], br[0]), fr[5]), sp);                        // :183-184
      float njcp =           __fmul_rn(__fmul_rn(__fmul_rn(fp[1], br[1]), ploop), sp);      // :186-188
      njcp = __fadd_rn(njcp, __fmul_rn(__fmul_rn(__fmul_rn(fp[2], br[2]), ploop), sp));
      njcp = __fadd_rn(njcp, __fmul_rn(__fmul_rn(__fmul_rn(fp[4], br[4]), ploop), sp));
      mocc[i] = (float) (1.0 - (double) njcp);                                              // :189
      fsc[i] = fr[5];
    }
    const int m = min(32, L - i0 + 1);
    float myb = 0.0f, mye = 0.0f;
    for (int k = 0; k < m; k++) {                                      // the running sums, in row order (every lane carries them)
      bt = __fadd_rn(bt, __shfl_sync(0xffffffffu, tb, k));
      et = __fadd_rn(et, __shfl_sync(0xffffffffu, te, k));
      if (k == lane) { myb = bt; mye = et; }
    }
    if (i <= L) { btot[i] = myb; etot[i] = mye; }
  }
}

int launch_domain_decoding(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
                           const float *d_fx, const float *d_bx, const int64_t *d_xoff, const int64_t *d_doff, float *d_dec)
{
  if (n <= 0) return P7GPU_OK;
  P7G_CHECK(ensure_len_table(prof, db.max_length));
  const Lane &ln = ctx->lanes[lane];
  DecParams p;
  p.fx = d_fx; p.bx = d_bx; p.xoff = d_xoff; p.doff = d_doff; p.idx = d_idx; p.len = db.len; p.lc = prof->d_len; p.n = n; p.dec = d_dec;
  const long long blocks = (n * 32 + 127) / 128;
  { LaunchTimer t(ctx, P7GPU_KERNEL_TRIG, ln.stream); domain_decoding_kernel<<<(unsigned) blocks, 128, 0, ln.stream>>>(p); }
  P7G_CUDA(cudaGetLastError());
  ctx->launches++;
  return P7GPU_OK;
}

}  // namespace p7g
