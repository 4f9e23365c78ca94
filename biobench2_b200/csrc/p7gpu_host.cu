// p7gpu_host.cu — host side of libp7gpu.so: context, profile repack, packed database, per-L tables and
// the batched stage entry points of include/p7gpu.h.  No CPU compute fallback lives here: every stage
// entry launches a CUDA kernel or fails with P7GPU_ESYS.
#include "p7gpu_internal.h"
#include <thread>
#include <unordered_set>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <iterator>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <algorithm>

namespace p7g {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...)
{
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}

DiagGeom diag_geom(int M)
{
  DiagGeom g; g.M = M;
  int Q = (M + 1 + 63) / 64;          // 64*Q >= M+1: at least one pad column
  // odd Q: lane stride Q words is conflict-free by itself.  Even Q <= MAX_EVEN_Q: lanes take consecutive words and the
  // registers of a lane are 32 words apart (conflict-free for any Q, at the price of a longer wrap-free row).  Larger even
  // Q are rounded up to odd (<= 6% of pad columns there).
  if ((Q & 1) == 0 && Q > MAX_EVEN_Q) Q++;
  g.Q = Q; g.P = 64 * Q;
  g.rowWords = (Q & 1) ? g.P + DIAG_CHUNK + Q : g.P + DIAG_CHUNK + 32 * (Q - 1) + 1;
  return g;
}

int ensure_scratch(p7gpu_ctx *ctx, size_t bytes)
{
  if (bytes <= ctx->d_scratch_bytes) return P7GPU_OK;
  if (ctx->d_scratch) { P7G_CUDA(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_scratch); ctx->d_scratch = nullptr; ctx->d_scratch_bytes = 0; }
  size_t want = std::max(bytes, (size_t) 1 << 20);
  P7G_CUDA(cudaMalloc(&ctx->d_scratch, want));
  ctx->d_scratch_bytes = want;
  return P7GPU_OK;
}
int ensure_scratch2(p7gpu_ctx *ctx, size_t bytes)
{
  if (bytes <= ctx->d_scratch2_bytes) return P7GPU_OK;
  if (ctx->d_scratch2) { P7G_CUDA(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_scratch2); ctx->d_scratch2 = nullptr; ctx->d_scratch2_bytes = 0; }
  size_t want = std::max(bytes, (size_t) 1 << 20);
  P7G_CUDA(cudaMalloc(&ctx->d_scratch2, want));
  ctx->d_scratch2_bytes = want;
  return P7GPU_OK;
}
int ensure_pinned(p7gpu_ctx *ctx, size_t bytes)
{
  if (bytes <= ctx->h_pinned_bytes) return P7GPU_OK;
  if (ctx->h_pinned) { P7G_CUDA(cudaStreamSynchronize(ctx->stream)); cudaFreeHost(ctx->h_pinned); ctx->h_pinned = nullptr; ctx->h_pinned_bytes = 0; }
  size_t want = std::max(bytes, (size_t) 1 << 20);
  P7G_CUDA(cudaMallocHost(&ctx->h_pinned, want));
  ctx->h_pinned_bytes = want;
  return P7GPU_OK;
}

// ---- per-L constants (host libm; p7_oprofile.c:676-699,1085-1127; p7_bg.c:190-199,358-363) ----------
static uint8_t byteify_cost(float scale_b, float sc)
{
  sc = -1.0f * roundf(scale_b * sc);
  return (sc > 255.) ? 255 : (uint8_t) sc;
}
static int16_t wordify(float scale_w, float sc)
{
  sc = roundf(scale_w * sc);
  if      (sc >=  32767.0) return  32767;
  else if (sc <= -32768.0) return -32768;
  else return (int16_t) sc;
}
// totscale and score of a Forward matrix from its specials, exactly as forward_engine accumulates them (fwdback.c:434, :463):
// float += double log of every scale factor > 1 in row order; score = totscale + log(xC * xf[C][MOVE]) in double, stored to float
void fbx_host_fwd_score(const float *xmx, int L, float pmove, float *totscale, float *sc, int *status)
{
  float tot = 0.0f;
  for (int i = 1; i <= L; i++) { const float s = xmx[(size_t) i * 6 + 5]; if (s != 1.0f) tot = (float) ((double) tot + log((double) s)); }
  const float xC = xmx[(size_t) L * 6 + 4];
  *totscale = tot;
  if (std::isnan(xC) || (L > 0 && xC == 0.0f) || std::isinf(xC)) { *status = P7GPU_ERANGE; *sc = xC; }   // :456-458
  else { *status = P7GPU_OK; *sc = (float) ((double) tot + log((double) (xC * pmove))); }
}
// the same from the compact block gather_scale_kernel writes: blk[0] = xC(L), blk[i] = SCALE(i)
void fbx_host_fwd_score_compact(const float *blk, int L, float pmove, float *totscale, float *sc, int *status)
{
  float tot = 0.0f;
  for (int i = 1; i <= L; i++) { const float s = blk[i]; if (s != 1.0f) tot = (float) ((double) tot + log((double) s)); }
  const float xC = blk[0];
  *totscale = tot;
  if (std::isnan(xC) || (L > 0 && xC == 0.0f) || std::isinf(xC)) { *status = P7GPU_ERANGE; *sc = xC; }   // :456-458
  else { *status = P7GPU_OK; *sc = (float) ((double) tot + log((double) (xC * pmove))); }
}
// same for backward_engine (:553, :689, :735): row L first, then rows L-1 .. 1
void fbx_host_bwd_score(const float *xmx, int L, float *totscale, float *sc, int *status)
{
  float tot = (float) log((double) xmx[(size_t) L * 6 + 5]);
  for (int i = L - 1; i >= 1; i--) { const float s = xmx[(size_t) i * 6 + 5]; if (s > 1.0f) tot = (float) ((double) tot + log((double) s)); }
  const float xN = xmx[1];
  *totscale = tot;
  if (std::isnan(xN) || (L > 0 && xN == 0.0f) || std::isinf(xN)) { *status = P7GPU_ERANGE; *sc = xN; }   // :731-733
  else { *status = P7GPU_OK; *sc = (float) ((double) tot + log((double) xN)); }
}

float host_bias_filtersc(const BiasHmm &h, const uint8_t *x, int L)
{
  const float p1 = (float) L / (float) (L + 1);                                  // p7_bg_SetLength, p7_bg.c:193-196
  const float t00 = p1, t01 = 1.0f - p1, t10 = h.t[1][0], t11 = h.t[1][1], t02 = h.t[0][2], t12 = h.t[1][2];
  const float la = (float) L * logf(p1), lb = logf(1.0f - p1);                   // p7_bg.c:481
  float logsc = 0.0f;
  if (L == 0) return ((float) log((double) h.pi[2]) + la) + lb;
  float dp0 = h.eo[x[0]][0] * h.pi[0], dp1 = h.eo[x[0]][1] * h.pi[1];
  float mx = 0.0f; if (dp0 > mx) mx = dp0; if (dp1 > mx) mx = dp1;
  dp0 = dp0 / mx; dp1 = dp1 / mx;
  logsc = logsc + (float) log((double) mx);
  for (int i = 2; i <= L; i++) {
    const int r = x[i - 1];
    float n0 = 0.0f + dp0 * t00; n0 = n0 + dp1 * t10; n0 = n0 * h.eo[r][0];
    float n1 = 0.0f + dp0 * t01; n1 = n1 + dp1 * t11; n1 = n1 * h.eo[r][1];
    mx = 0.0f; if (n0 > mx) mx = n0; if (n1 > mx) mx = n1;
    dp0 = n0 / mx; dp1 = n1 / mx;
    logsc = logsc + (float) log((double) mx);
  }
  float last = 0.0f + dp0 * t02; last = last + dp1 * t12;
  logsc = logsc + (float) log((double) last);
  return (logsc + la) + lb;
}

int fix_unsure_filtersc(p7gpu_ctx *ctx, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, float *d_filtersc,
                        const int32_t *d_unsure_list, const unsigned int *d_unsure_count, float *h_filtersc, unsigned *nfixed)
{
  unsigned n = 0;
  if (nfixed) *nfixed = 0;
  if (!d_unsure_list || !d_unsure_count) return P7GPU_OK;
  cudaStream_t st = ctx->stream;
  P7G_CUDA(cudaMemcpyAsync(&n, d_unsure_count, sizeof(n), cudaMemcpyDeviceToHost, st));
  P7G_CUDA(cudaStreamSynchronize(st));
  if (n == 0) return P7GPU_OK;
  std::vector<int32_t> list(n);
  P7G_CUDA(cudaMemcpyAsync(list.data(), d_unsure_list, 4 * (size_t) n, cudaMemcpyDeviceToHost, st));
  P7G_CUDA(cudaStreamSynchronize(st));
  std::vector<uint8_t> seq;
  for (unsigned k = 0; k < n; k++) {
    const int32_t j = list[k];
    int64_t s = db.first + j, off = 0; int32_t L = 0;
    if (d_idx) P7G_CUDA(cudaMemcpy(&s, d_idx + j, 8, cudaMemcpyDeviceToHost));
    P7G_CUDA(cudaMemcpy(&off, db.off + s, 8, cudaMemcpyDeviceToHost));
    P7G_CUDA(cudaMemcpy(&L, db.len + s, 4, cudaMemcpyDeviceToHost));
    seq.resize((size_t) std::max(L, 1));
    if (L > 0) P7G_CUDA(cudaMemcpy(seq.data(), db.res + off, (size_t) L, cudaMemcpyDeviceToHost));
    const float f = host_bias_filtersc(prof->bias, seq.data(), L);
    P7G_CUDA(cudaMemcpy(d_filtersc + j, &f, 4, cudaMemcpyHostToDevice));
    if (h_filtersc) h_filtersc[j] = f;
  }
  ctx->bias_fixed += n;
  if (nfixed) *nfixed = n;
  return P7GPU_OK;
}

float host_nullsc(int L)                                   // p7_bg_NullOne, p7_bg.c:358-363 (double math, float result)
{
  float p1 = (float) L / (float) (L + 1);
  return (float) ((double) (float) L * log((double) p1) + log(1. - (double) p1));
}
double host_gumbel_surv(double x, double mu, double lambda)   // easel/esl_gumbel.c:129-138
{
  double y = lambda * (x - mu), ey = -exp(-y);
  if (fabs(ey) < 5e-9) return -ey;
  return 1 - exp(ey);
}
double host_exp_surv(double x, double mu, double lambda)      // easel/esl_exponential.c:127-132
{
  if (x < mu) return 1.0;
  return exp(-lambda * (x - mu));
}
// seq_score = (sc - nullsc) / eslCONST_LOG2 is a float in the reference (p7_pipeline.c:691,699,720,728)
double host_filter_pvalue_gumbel(float sc, float nullsc, float mu, float lambda)
{
  float seq_score = (sc - nullsc) / 0.69314718055994529;
  return host_gumbel_surv((double) seq_score, (double) mu, (double) lambda);
}
double host_filter_pvalue_exp(float sc, float filtersc, float tau, float lambda)
{
  float seq_score = (sc - filtersc) / 0.69314718055994529;
  return host_exp_surv((double) seq_score, (double) tau, (double) lambda);
}

// MSV score as a function of the final xJ byte (ssvfilter.c:921-923 == msvfilter.c:205-207)
static float msv_score_of_xj(const p7gpu_profile *prof, int xJ, int tjb)
{
  float sc = ((float) (xJ - tjb) - (float) prof->msv.base_b);
  sc /= prof->msv.scale_b;
  sc -= 3.0f;
  return sc;
}

static LenConst len_const(const p7gpu_profile *prof, int L, double F1)
{
  LenConst c;
  c.tjb_b   = byteify_cost(prof->scale_b, logf(3.0f / (float) (L + 3)));
  c.pmove   = (2.0f + prof->nj) / ((float) L + 2.0f + prof->nj);
  c.xw_move = wordify(prof->scale_w, logf(c.pmove));
  float p1  = (float) L / (float) (L + 1);
  c.nullsc  = host_nullsc(L);
  c.pad = 0; c.pad2 = 0;
  c.logf_p1 = logf(p1);
  c.logf_1mp1 = logf(1. - p1);
  // smallest xJ whose P-value passes: P(xJ) is non-increasing in xJ, so bisect, then verify the boundary pair with
  // the reference expression itself (p7_pipeline.c:691-693: fail iff P > F1)
  if (L == 0) c.msv_xj_min = 256;                                       // p7_pipeline.c:682: empty targets never enter
  else {
    auto passes = [&](int xJ) { return !(host_filter_pvalue_gumbel(msv_score_of_xj(prof, xJ, c.tjb_b), c.nullsc, prof->evparam[0], prof->evparam[1]) > F1); };
    int lo = 0, hi = 256;                                               // invariant: everything >= hi passes (256 = none)
    while (lo < hi) { int mid = (lo + hi) / 2; if (passes(mid)) hi = mid; else lo = mid + 1; }
    while (hi > 0 && passes(hi - 1)) hi--;                              // (never iterates unless libm is non-monotone)
    c.msv_xj_min = (uint16_t) hi;
  }
  return c;
}

// hmmscan's block prefilter (scan.cu): tjb_b and the smallest passing xJ of a model that has no p7gpu_profile, for the given
// target lengths; the same len_const() as every uploaded profile uses, so the decisions are the same
void scan_len_words(const p7gpu_oprofile_view *om, const int32_t *lengths, int64_t n, double F1, uint32_t *out)
{
  p7gpu_profile tmp;
  tmp.M = om->M; tmp.Kp = om->Kp; tmp.nj = om->nj;
  memcpy(tmp.evparam, om->evparam, sizeof(tmp.evparam));
  tmp.msv.tbm_b = om->tbm_b; tmp.msv.tec_b = om->tec_b; tmp.msv.base_b = om->base_b; tmp.msv.bias_b = om->bias_b;
  tmp.msv.scale_b = om->scale_b; tmp.scale_b = om->scale_b;
  tmp.scale_w = 1.0f;                                                   // the Viterbi constant of LenConst is not used by the prefilter
  for (int64_t k = 0; k < n; k++) {
    const LenConst c = len_const(&tmp, lengths[k], F1);
    out[k] = (uint32_t) c.tjb_b | ((uint32_t) c.msv_xj_min << 16);
  }
}

static void sync_all_lanes(p7gpu_ctx *ctx)
{
  for (int l = 0; l < NLANES; l++) if (ctx->lanes[l].stream || l == 0) cudaStreamSynchronize(ctx->lanes[l].stream);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
}

int ensure_len_table(const p7gpu_profile *prof, int Lmax, double F1)
{
  std::lock_guard<std::mutex> lock(prof->len_mutex);
  if (F1 < 0) F1 = prof->len_F1 >= 0 ? prof->len_F1 : DEFAULT_F1;
  if (F1 != prof->len_F1) {                                             // thresholds changed: rebuild everything
    if (prof->d_len) { sync_all_lanes(prof->ctx); cudaFreeAsync(prof->d_len, prof->ctx->stream); prof->d_len = nullptr; }
    prof->len_cap = 0; prof->h_len.clear(); prof->len_F1 = F1; prof->len_sparse = false;
  }
  if (Lmax < prof->len_cap && !prof->len_sparse) return P7GPU_OK;
  if (prof->len_sparse) {                                               // prepared for a fixed set of lengths (p7gpu_profile_prepare_lengths)
    if (Lmax < prof->len_cap) return P7GPU_OK;                          // the caller vouched for the lengths; only the bound is checked here
    prof->len_cap = 0; prof->h_len.clear(); prof->len_sparse = false;   // longer targets than prepared for: build the dense table
    if (prof->d_len) { sync_all_lanes(prof->ctx); cudaFreeAsync(prof->d_len, prof->ctx->stream); prof->d_len = nullptr; }
  }
  int cap = std::max(Lmax + 1, std::max(4096, prof->len_cap * 2));
  prof->h_len.resize(cap);
  {                                                                     // ~0.7 us per length (bisection over exp()): all host threads
    const int L0 = prof->len_cap, nL = cap - L0;
    unsigned hw = std::thread::hardware_concurrency();
    const int T = nL < 2048 ? 1 : (int) std::min<unsigned>(hw ? hw : 1, 16);
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
      th.emplace_back([=]() { for (int L = L0 + (int) ((long long) nL * t / T); L < L0 + (int) ((long long) nL * (t + 1) / T); L++) prof->h_len[L] = len_const(prof, L, F1); });
    for (auto &x : th) x.join();
  }
  LenConst *d_new = nullptr;
  P7G_CUDA(cudaMallocAsync((void **) &d_new, sizeof(LenConst) * cap, prof->ctx->stream));
  P7G_CUDA(cudaMemcpyAsync(d_new, prof->h_len.data(), sizeof(LenConst) * cap, cudaMemcpyHostToDevice, prof->ctx->stream));
  P7G_CUDA(cudaStreamSynchronize(prof->ctx->stream));
  if (prof->d_len) { sync_all_lanes(prof->ctx); cudaFreeAsync(prof->d_len, prof->ctx->stream); }
  prof->d_len = d_new;
  prof->len_cap = cap;
  return P7GPU_OK;
}

}  // namespace p7g
// hmmscan: the targets (queries) are few and fixed while the profiles stream past, so each profile's per-length constants are
// needed for the distinct query lengths only — 0.7 us each instead of >= 4096 of them per profile.
extern "C" int p7gpu_profile_prepare_lengths(p7gpu_profile *prof, const int32_t *lengths, int64_t n, double F1)
{
  using namespace p7g;
  if (!prof || n < 0 || (n > 0 && !lengths)) return P7GPU_EINVAL;
  std::lock_guard<std::mutex> lock(prof->len_mutex);
  if (F1 < 0) F1 = DEFAULT_F1;
  int32_t Lmax = 0;
  for (int64_t k = 0; k < n; k++) { if (lengths[k] < 0) return P7GPU_EINVAL; Lmax = std::max(Lmax, lengths[k]); }
  P7G_CUDA(cudaSetDevice(prof->ctx->device));
  if (prof->d_len) { sync_all_lanes(prof->ctx); cudaFreeAsync(prof->d_len, prof->ctx->stream); prof->d_len = nullptr; }
  const int cap = Lmax + 1;
  prof->h_len.assign((size_t) cap, LenConst{});
  for (int64_t k = 0; k < n; k++) prof->h_len[(size_t) lengths[k]] = len_const(prof, lengths[k], F1);
  P7G_CUDA(cudaMallocAsync((void **) &prof->d_len, sizeof(LenConst) * (size_t) cap, prof->ctx->stream));
  P7G_CUDA(cudaMemcpyAsync(prof->d_len, prof->h_len.data(), sizeof(LenConst) * (size_t) cap, cudaMemcpyHostToDevice, prof->ctx->stream));
  P7G_CUDA(cudaStreamSynchronize(prof->ctx->stream));
  prof->len_cap = cap; prof->len_F1 = F1; prof->len_sparse = true;
  return P7GPU_OK;
}
namespace p7g {

int dp_cells_per_lane(int M)
{
  static const int cs[] = {1,2,3,4,5,6,7,8,9,10,11,12,13,14,16,18,20,24,28,32,40,48,56,64,80,96,112,130};
  for (int c : cs) if (32 * c >= M) return c;
  return -1;
}

// CTA-per-target geometry of the long-model Forward/Backward kernels (fwdbwd_mw.cuh): W warps x 32 lanes x C cells >= M.
// Instances exist for W = 1 (2 <= C <= 16), W = 2 (5 <= C <= 16), W = 4 (9 <= C <= 16), W = 8 (9 <= C <= 14), W = 12 (10 <= C <= 11);
// the smallest W that fits is used (W = 1 up to M = 512).  Environment (tuning and tests only): P7GPU_MW_MINM = smallest M that
// uses these kernels (default 1: all models; below it the first-generation one-warp kernels of fwdbwd.cu run),
// P7GPU_MW_CMAX = upper bound on C when choosing W.
bool mw_geometry(int M, int *C, int *W)
{
  static const int ws[5] = {1, 2, 4, 8, 12}, cmax[5] = {16, 16, 16, 14, 11}, cmin[5] = {2, 5, 9, 9, 10};
  const char *e1 = getenv("P7GPU_MW_MINM"), *e2 = getenv("P7GPU_MW_CMAX");
  const int minM = e1 ? atoi(e1) : 1, cap = e2 ? atoi(e2) : 0;
  if (M < minM) return false;
  for (int t = 0; t < 5; t++) {
    int cm = cmax[t]; if (cap > 0 && cap < cm) cm = cap;
    const int c = (M + 32 * ws[t] - 1) / (32 * ws[t]);
    if (c > cm) continue;
    *C = c < cmin[t] ? cmin[t] : c; *W = ws[t];
    return true;
  }
  return false;
}

// profile tables live in the context's stream-ordered pool: an upload is cudaMallocAsync + cudaMemcpyAsync on the main stream
// (the host vectors are pageable, so the copy has left them when the call returns), a destroy is cudaFreeAsync — hmmscan
// uploads and drops one profile per model, and cudaMalloc/cudaFree round trips were most of its time per model
// All tables of one profile travel as ONE block: every pool_upload() of a p7gpu_profile_upload() call appends its bytes to a
// host arena (256-byte aligned) and remembers where the device pointer has to go; arena_commit() then does one cudaMallocAsync,
// one cudaMemcpyAsync and fixes the pointers up.  (Seventeen allocate + copy pairs per profile were most of what a model costs
// hmmscan once the block prefilter has decided that it must be uploaded.)
struct UploadArena {
  std::vector<char> bytes;
  std::vector<std::pair<void **, size_t>> fix;       // (where the device pointer is stored, offset inside the block)
};
static thread_local UploadArena *g_arena = nullptr;

template <class T>
static int pool_upload(p7gpu_ctx *ctx, T **dptr, const void *src, size_t bytes)
{
  if (g_arena) {
    const size_t off = (g_arena->bytes.size() + 255) & ~(size_t) 255;
    g_arena->bytes.resize(off + std::max<size_t>(bytes, 16));
    if (bytes) memcpy(g_arena->bytes.data() + off, src, bytes);
    g_arena->fix.emplace_back((void **) dptr, off);
    return P7GPU_OK;
  }
  P7G_CUDA(cudaMallocAsync((void **) dptr, std::max<size_t>(bytes, 16), ctx->stream));
  if (bytes) P7G_CUDA(cudaMemcpyAsync(*dptr, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return P7GPU_OK;
}

static int arena_commit(p7gpu_ctx *ctx, UploadArena &a, void **d_block)
{
  *d_block = nullptr;
  if (a.bytes.empty()) return P7GPU_OK;
  P7G_CUDA(cudaMallocAsync(d_block, a.bytes.size(), ctx->stream));
  P7G_CUDA(cudaMemcpyAsync(*d_block, a.bytes.data(), a.bytes.size(), cudaMemcpyHostToDevice, ctx->stream));
  for (auto &f : a.fix) *f.first = (char *) *d_block + f.second;
  return P7GPU_OK;
}

static int nq_of(int M, int w) { int q = (M - 1) / w + 1; return q < 2 ? 2 : q; }   // impl_sse.h:25-27

// Packed (2 x int16 per register) Viterbi tables and clamp window (vit.cu: vitp_kernel; proof in DESIGN.md 4.3).
// Node k = 1 + h*32*C2 + lane*C2 + j for half h, register j.  Every transition / emission is clamped at -T; values live
// in a per-target window [Fl, Fl+T] whose floor is Fl = base_w + xw_move(L) + floff; targets that leave the window are
// redone by the exact 32-bit kernel.
static int build_vit_packed(p7gpu_profile *prof, const p7gpu_oprofile_view *om, int Q8)
{
  const int M = prof->M, Kp = prof->Kp;
  const int C2 = vit_packed_c2(M);
  VitPacked &vp = prof->vitp;
  vp.enabled = 0; vp.C2 = C2;
  if (C2 < 1) return P7GPU_OK;
  auto rw_at = [&](int x, int k) { return (int) om->rwv[((size_t) x * Q8 + (k - 1) % Q8) * 8 + (k - 1) / Q8]; };
  auto tw_at = [&](int t, int k) {
    const int q8 = (k - 1) % Q8, z8 = (k - 1) / Q8;
    return (int) (t < 7 ? om->twv[((size_t) 7 * q8 + t) * 8 + z8] : om->twv[((size_t) 7 * Q8 + q8) * 8 + z8]);
  };
  int tbm_min = 32767, rw_max = 0, min_rwstar = 32767;
  for (int k = 1; k <= M; k++) tbm_min = std::min(tbm_min, tw_at(0, k));
  for (int x = 0; x < Kp; x++) {
    int star = -32768;
    for (int k = 1; k <= M; k++) { star = std::max(star, rw_at(x, k)); rw_max = std::max(rw_max, rw_at(x, k)); }
    if (star > -32768) min_rwstar = std::min(min_rwstar, star);
  }
  for (int t = 0; t < 8; t++) for (int k = 1; k <= M; k++) if (tw_at(t, k) > 0) return P7GPU_OK;   // not log-probabilities: keep the exact kernel only
  const int T = (65535 - rw_max) / 2;
  int floff = std::min(0, tbm_min);
  if (min_rwstar < 32767) floff = std::min(floff, tbm_min + min_rwstar - 1);
  floff = std::min(floff, -(int) om->ddbound_w);
  vp.T = T; vp.W0 = -32768 + T; vp.W1 = vp.W0 + T; vp.floff = floff;
  if (tbm_min <= -32768 || -floff > T - 64 || C2 > 46) return P7GPU_OK;      // window cannot be placed / tables too large
  auto cl = [&](int v) { return (uint32_t) (uint16_t) (int16_t) std::max(v, -T); };
  const int H = 32 * C2;
  std::vector<uint32_t> rwp((size_t) NTABROWS * C2 * 32), twp((size_t) 8 * C2 * 32), scanw((size_t) 6 * 32);
  std::vector<int> tdd((size_t) 2 * H, -T);                                  // clamped tDD by node index k-1, pads = -T
  for (int k = 1; k <= M; k++) tdd[k - 1] = std::max(tw_at(7, k), -T);
  for (int lane = 0; lane < 32; lane++)
    for (int j = 0; j < C2; j++) {
      const int klo = 1 + lane * C2 + j, khi = klo + H;
      for (int x = 0; x < NTABROWS; x++) {
        const int lo = (x < Kp && klo <= M) ? rw_at(x, klo) : -32768, hi = (x < Kp && khi <= M) ? rw_at(x, khi) : -32768;
        rwp[((size_t) x * C2 + j) * 32 + lane] = cl(lo) | (cl(hi) << 16);
      }
      for (int t = 0; t < 8; t++) {
        const int lo = klo <= M ? tw_at(t, klo) : -32768, hi = khi <= M ? tw_at(t, khi) : -32768;
        twp[((size_t) t * C2 + j) * 32 + lane] = cl(lo) | (cl(hi) << 16);
      }
    }
  // lane totals TL(h,l) = sum of the lane's C2 tDD; W_d(l) = sum of TL over lanes l-d+1..l (or -T when l < d);
  // WH(l) = sum of the high half's TL over lanes 0..l.  All clamped at -T (a clamped sum always reaches the floor).
  auto TL = [&](int h, int l) { long long a = 0; for (int j = 0; j < C2; j++) a += tdd[(size_t) h * H + l * C2 + j]; return a; };
  for (int l = 0; l < 32; l++) {
    for (int di = 0; di < 5; di++) {
      const int d = 1 << di;
      long long lo = -T, hi = -T;
      if (l >= d) { lo = 0; hi = 0; for (int m = l - d + 1; m <= l; m++) { lo += TL(0, m); hi += TL(1, m); } }
      scanw[(size_t) di * 32 + l] = cl((int) std::max<long long>(lo, -T)) | (cl((int) std::max<long long>(hi, -T)) << 16);
    }
    long long wh = 0; for (int m = 0; m <= l; m++) wh += TL(1, m);
    scanw[(size_t) 5 * 32 + l] = 0u | (cl((int) std::max<long long>(wh, -T)) << 16);
  }
  p7gpu_ctx *ctx = prof->ctx;
  P7G_CHECK(pool_upload(ctx, &prof->d_rwp, rwp.data(), rwp.size() * 4));
  P7G_CHECK(pool_upload(ctx, &prof->d_twp, twp.data(), twp.size() * 4));
  P7G_CHECK(pool_upload(ctx, &prof->d_scanw, scanw.data(), scanw.size() * 4));
  vp.enabled = 1;
  return P7GPU_OK;
}

}  // namespace p7g

using namespace p7g;

// =================================================================================================
extern "C" {

const char *p7gpu_last_error(void) { return g_err; }

}  // extern "C"
// Live contexts: a pipeline result hands its big host buffers back to its context's cache when it is destroyed, and a result
// may outlive its context (garbage collection order in a host language): p7gpu_result_destroy asks here first.
namespace p7g {
static std::mutex g_ctx_mu;
static std::unordered_set<const p7gpu_ctx *> g_live_ctx;
std::mutex &ctx_registry_mutex() { return g_ctx_mu; }
bool ctx_registered_locked(const p7gpu_ctx *ctx) { return g_live_ctx.count(ctx) != 0; }
}
extern "C" {

int p7gpu_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int p7gpu_ctx_create(int device, p7gpu_ctx **ret_ctx)
{
  if (!ret_ctx) return P7GPU_EINVAL;
  *ret_ctx = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); libp7gpu has no CPU fallback", e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
    return P7GPU_ESYS;
  }
  if (device < 0 || device >= ndev) { set_error("device %d out of range (0..%d)", device, ndev - 1); return P7GPU_EINVAL; }
  P7G_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  P7G_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) { set_error("device %d is sm_%d%d; libp7gpu is built for sm_100a only", device, prop.major, prop.minor); return P7GPU_ESYS; }
  p7gpu_ctx *ctx = new p7gpu_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = (int) prop.sharedMemPerBlockOptin;
  for (int l = 0; l < NLANES; l++) {
    P7G_CUDA(cudaStreamCreateWithFlags(&ctx->lanes[l].stream, cudaStreamNonBlocking));
    P7G_CUDA(cudaMalloc(&ctx->lanes[l].counters, LANE_COUNTER_WORDS * sizeof(unsigned int)));
    P7G_CUDA(cudaMemset(ctx->lanes[l].counters, 0, LANE_COUNTER_WORDS * sizeof(unsigned int)));
  }
  P7G_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  {  // keep stream-ordered allocations (pipeline scratch) cached in the pool instead of returning them at every sync
    cudaMemPool_t pool; uint64_t thr = UINT64_MAX;
    P7G_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    P7G_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  }
  ctx->stream = ctx->lanes[0].stream;
  ctx->own_stream = true;
  { std::lock_guard<std::mutex> lk(g_ctx_mu); g_live_ctx.insert(ctx); }
  *ret_ctx = ctx;
  return P7GPU_OK;
}

void p7gpu_ctx_destroy(p7gpu_ctx *ctx)
{
  if (!ctx) return;
  { std::lock_guard<std::mutex> lk(g_ctx_mu); g_live_ctx.erase(ctx); }
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int l = 0; l < NLANES; l++) {
    if (ctx->lanes[l].counters) cudaFree(ctx->lanes[l].counters);
    if (ctx->lanes[l].stream && (l > 0 || ctx->own_stream)) cudaStreamDestroy(ctx->lanes[l].stream);
  }
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (auto &t : ctx->timed) { cudaEventDestroy(t.start); cudaEventDestroy(t.stop); }
  if (ctx->d_scratch) cudaFree(ctx->d_scratch);
  if (ctx->d_scratch2) cudaFree(ctx->d_scratch2);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  delete ctx;
}

int p7gpu_ctx_set_stream(p7gpu_ctx *ctx, void *cuda_stream)
{
  if (!ctx) return P7GPU_EINVAL;
  P7G_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t) cuda_stream;
  ctx->lanes[0].stream = ctx->stream;
  ctx->own_stream = false;
  return P7GPU_OK;
}

int64_t p7gpu_ctx_launch_count(const p7gpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

int p7gpu_ctx_set_option(p7gpu_ctx *ctx, const char *name, int value)
{
  if (!ctx || !name) return P7GPU_EINVAL;
  if (!strcmp(name, "vit_exact_only")) { ctx->vit_exact_only = value; return P7GPU_OK; }
  if (!strcmp(name, "fb_one_warp"))    { ctx->fb_one_warp = value; return P7GPU_OK; }
  if (!strcmp(name, "fb_exact"))       { ctx->fb_exact = value; return P7GPU_OK; }
  if (!strcmp(name, "fb_fast"))        { ctx->fb_fast = value; return P7GPU_OK; }
  if (!strcmp(name, "fbx_groups"))     { ctx->fbx_groups = value; return P7GPU_OK; }
  if (!strcmp(name, "bias_window_log2")) { ctx->bias_window = value > 0 ? ldexpf(1.0f, -value) : 0.f; return P7GPU_OK; }
  if (!strcmp(name, "fbx_warps"))      { ctx->fbx_warps = value; return P7GPU_OK; }
  if (!strcmp(name, "ssv_tmem"))       { ctx->ssv_tmem = value; return P7GPU_OK; }
  if (!strcmp(name, "bias_warp"))      { ctx->bias_warp = value; return P7GPU_OK; }
  set_error("set_option: unknown option '%s'", name);
  return P7GPU_EINVAL;
}

int64_t p7gpu_ctx_bias_recomputed(const p7gpu_ctx *ctx) { return ctx ? ctx->bias_fixed : -1; }

int64_t p7gpu_ctx_last_vit_redo(p7gpu_ctx *ctx)
{
  if (!ctx) return -1;
  unsigned int v = 0;
  if (cudaStreamSynchronize(ctx->lanes[0].stream) != cudaSuccess) return -1;
  if (cudaMemcpy(&v, ctx->lanes[0].counters + CTR_VIT_REDO, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int64_t) v;
}

int p7gpu_ctx_synchronize(p7gpu_ctx *ctx)
{
  if (!ctx) return P7GPU_EINVAL;
  P7G_CUDA(cudaStreamSynchronize(ctx->stream));
  return P7GPU_OK;
}

// ---------------------------------------------------------------------------------------------------
// Profile upload: de-stripe the P7_OPROFILE arrays (node k -> vector q=(k-1)%Q, lane z=(k-1)/Q,
// p7_oprofile.c:789-792,845-848,938-941) and build the GPU layouts.
// ---------------------------------------------------------------------------------------------------
int p7gpu_profile_upload(p7gpu_ctx *ctx, const p7gpu_oprofile_view *om, p7gpu_profile **ret_prof)
{
  if (!ctx || !om || !ret_prof) return P7GPU_EINVAL;
  *ret_prof = nullptr;
  if (om->M < 1 || om->Kp != P7GPU_KP || !om->rbv) { set_error("profile_upload: need an amino profile (Kp=29) with M>=1"); return P7GPU_EINVAL; }
  if (om->M > 64 * MAX_Q - 1) { set_error("profile_upload: M=%d exceeds the supported maximum %d", om->M, 64 * MAX_Q - 1); return P7GPU_EINVAL; }
  P7G_CUDA(cudaSetDevice(ctx->device));
  const int M = om->M, Kp = om->Kp;
  const int Q16 = nq_of(M, 16), Q8 = nq_of(M, 8), Q4 = nq_of(M, 4);
  p7gpu_profile *prof = new p7gpu_profile();
  static thread_local UploadArena arena;                // kept between calls: its buffer is reused profile after profile
  arena.bytes.clear(); arena.fix.clear();
  struct ArenaScope { ArenaScope(UploadArena *a) { g_arena = a; } ~ArenaScope() { g_arena = nullptr; } } arena_scope(&arena);
  prof->ctx = ctx; prof->M = M; prof->Kp = Kp; prof->nj = om->nj;
  memcpy(prof->evparam, om->evparam, sizeof(prof->evparam));
  prof->msv.tbm_b = om->tbm_b; prof->msv.tec_b = om->tec_b; prof->msv.base_b = om->base_b; prof->msv.bias_b = om->bias_b;
  prof->msv.scale_b = om->scale_b; prof->scale_b = om->scale_b;
  prof->geom = diag_geom(M);
  const DiagGeom &g = prof->geom;

  // ---- SSV / MSV gain tables in the diagonal layout -------------------------------------------------
  {
    std::vector<int16_t> gs((size_t) NTABROWS * g.P), gm((size_t) NTABROWS * g.P);
    for (int x = 0; x < NTABROWS; x++)
      for (int col = 0; col < g.P; col++) {
        int16_t vs = (int16_t) SSV_PADGAIN, vm = -255;        // pad column / pad residue row
        if (x < Kp && col < M) {
          int k = col + 1, q = (k - 1) % Q16, z = (k - 1) / Q16;
          uint8_t rb = om->rbv[((size_t) x * Q16 + q) * 16 + z];
          int8_t  sb;
          if (om->sbv) sb = (int8_t) om->sbv[((size_t) x * (Q16 + 17) + q) * 16 + z];
          else { int a = (127 + om->bias_b) - rb; if (a < 0) a = 0; sb = (int8_t) ((uint8_t) a ^ 127); }   // p7_oprofile.c:711-752
          vs = (int16_t) (-(int) sb);
          vm = (int16_t) (-(int) rb);
        }
        gs[(size_t) x * g.P + col] = vs;
        gm[(size_t) x * g.P + col] = vm;
      }
    std::vector<uint32_t> ts((size_t) NTABROWS * g.rowWords), tm((size_t) NTABROWS * g.rowWords);
    const int H = 32 * g.Q;                                   // column distance between the two halves
    for (int x = 0; x < NTABROWS; x++)
      for (int c = 0; c < g.rowWords; c++) {
        int lo = c % g.P, hi = (c + H) % g.P;
        ts[(size_t) x * g.rowWords + c] = (uint32_t) (uint16_t) gs[(size_t) x * g.P + lo] | ((uint32_t) (uint16_t) gs[(size_t) x * g.P + hi] << 16);
        tm[(size_t) x * g.rowWords + c] = (uint32_t) (uint16_t) gm[(size_t) x * g.P + lo] | ((uint32_t) (uint16_t) gm[(size_t) x * g.P + hi] << 16);
      }
    // column-stationary layout for the Tensor Memory kernel: lane l, register q, half h <-> node k = 1 + h*32*Q + l*Q + q
    const int Qt = (M + 63) / 64;
    if (Qt <= 16) {
      const int NX = Qt <= 1 ? 1 : Qt <= 2 ? 2 : Qt <= 4 ? 4 : Qt <= 8 ? 8 : 16;
      std::vector<uint32_t> tt((size_t) NTABROWS * NX * 32);
      auto gain = [&](int x, int col) { return (x < Kp && col < M) ? gs[(size_t) x * g.P + col] : (int16_t) SSV_PADGAIN; };   // col = k-1
      for (int x = 0; x < NTABROWS; x++)
        for (int q = 0; q < NX; q++)
          for (int l = 0; l < 32; l++) {
            const int16_t lo = q < Qt ? gain(x, l * Qt + q) : (int16_t) SSV_PADGAIN, hi = q < Qt ? gain(x, 32 * Qt + l * Qt + q) : (int16_t) SSV_PADGAIN;
            tt[((size_t) x * NX + q) * 32 + l] = (uint32_t) (uint16_t) lo | ((uint32_t) (uint16_t) hi << 16);
          }
      P7G_CHECK(pool_upload(ctx, &prof->d_ssv_tm, tt.data(), tt.size() * 4));
      prof->ssv_tm_Q = Qt;
    }
    P7G_CHECK(pool_upload(ctx, &prof->d_ssv_tab, ts.data(), ts.size() * 4));
    P7G_CHECK(pool_upload(ctx, &prof->d_msv_tab, tm.data(), tm.size() * 4));
  }

  // ---- Viterbi / Forward tables, lane-transposed [row][j][lane], node k = lane*C + j + 1 ----------------
  if (om->rwv && om->twv && om->rfv && om->tfv) {
    const int C = dp_cells_per_lane(M);
    prof->dpC = C;
    std::vector<int16_t> rw((size_t) NTABROWS * C * 32, (int16_t) -32768), tw((size_t) 8 * C * 32, (int16_t) -32768);
    std::vector<float>   rf((size_t) NTABROWS * C * 32, 0.0f),            tf((size_t) 8 * C * 32, 0.0f);
    for (int k = 1; k <= M; k++) {
      const int lane = (k - 1) / C, j = (k - 1) % C;
      const int q8 = (k - 1) % Q8, z8 = (k - 1) / Q8, q4 = (k - 1) % Q4, z4 = (k - 1) / Q4;
      for (int x = 0; x < Kp; x++) {
        rw[((size_t) x * C + j) * 32 + lane] = om->rwv[((size_t) x * Q8 + q8) * 8 + z8];
        rf[((size_t) x * C + j) * 32 + lane] = om->rfv[((size_t) x * Q4 + q4) * 4 + z4];
      }
      for (int t = 0; t < 7; t++) {
        tw[((size_t) t * C + j) * 32 + lane] = om->twv[((size_t) 7 * q8 + t) * 8 + z8];
        tf[((size_t) t * C + j) * 32 + lane] = om->tfv[((size_t) 7 * q4 + t) * 4 + z4];
      }
      tw[((size_t) 7 * C + j) * 32 + lane] = om->twv[((size_t) 7 * Q8 + q8) * 8 + z8];
      tf[((size_t) 7 * C + j) * 32 + lane] = om->tfv[((size_t) 7 * Q4 + q4) * 4 + z4];
    }
    P7G_CHECK(pool_upload(ctx, &prof->d_rw, rw.data(), rw.size() * 2));
    P7G_CHECK(pool_upload(ctx, &prof->d_tw, tw.data(), tw.size() * 2));
    P7G_CHECK(pool_upload(ctx, &prof->d_rf, rf.data(), rf.size() * 4));
    P7G_CHECK(pool_upload(ctx, &prof->d_tf, tf.data(), tf.size() * 4));
    {                                                                 // exact-order kernels (fbx.cu): striped values, wider loads
      const int Q = Q4, Qp = (Q4 + 3) / 4 * 4;
      std::vector<float> tfx((size_t) Q * 32, 0.0f), rfx((size_t) NTABROWS * 4 * Qp, 0.0f), tddx((size_t) 4 * Qp, 0.0f), tmdx((size_t) 4 * Qp, 0.0f);
      for (int q = 0; q < Q; q++)
        for (int z = 0; z < 4; z++) {
          for (int t = 0; t < 7; t++) tfx[((size_t) q * 4 + z) * 8 + t] = om->tfv[((size_t) 7 * q + t) * 4 + z];
          tfx[((size_t) q * 4 + z) * 8 + 7] = om->tfv[((size_t) 7 * Q + q) * 4 + z];
          tddx[(size_t) z * Qp + q] = tfx[((size_t) q * 4 + z) * 8 + 7]; tmdx[(size_t) z * Qp + q] = tfx[((size_t) q * 4 + z) * 8 + 4];
          for (int x = 0; x < Kp; x++) rfx[((size_t) x * 4 + z) * Qp + q] = om->rfv[((size_t) x * Q + q) * 4 + z];
        }
      P7G_CHECK(pool_upload(ctx, &prof->d_tfx, tfx.data(), tfx.size() * 4));
      P7G_CHECK(pool_upload(ctx, &prof->d_rfx, rfx.data(), rfx.size() * 4));
      P7G_CHECK(pool_upload(ctx, &prof->d_tddx, tddx.data(), tddx.size() * 4));
      P7G_CHECK(pool_upload(ctx, &prof->d_tmdx, tmdx.data(), tmdx.size() * 4));
      prof->fbxQ = Q; prof->fbxQp = Qp;
    }
    int mc = 0, mw = 0;
    if (mw_geometry(M, &mc, &mw)) {                                   // long models: [row][j][g], node k = g*mc + j + 1
      const int G = 32 * mw;
      std::vector<float> rfm((size_t) NTABROWS * mc * G, 0.0f), tfm((size_t) 8 * mc * G, 0.0f);
      for (int k = 1; k <= M; k++) {
        const int gl = (k - 1) / mc, j = (k - 1) % mc, q4 = (k - 1) % Q4, z4 = (k - 1) / Q4;
        for (int x = 0; x < Kp; x++) rfm[((size_t) x * mc + j) * G + gl] = om->rfv[((size_t) x * Q4 + q4) * 4 + z4];
        for (int t = 0; t < 7; t++)  tfm[((size_t) t * mc + j) * G + gl] = om->tfv[((size_t) 7 * q4 + t) * 4 + z4];
        tfm[((size_t) 7 * mc + j) * G + gl] = om->tfv[((size_t) 7 * Q4 + q4) * 4 + z4];
      }
      P7G_CHECK(pool_upload(ctx, &prof->d_rf_mw, rfm.data(), rfm.size() * 4));
      P7G_CHECK(pool_upload(ctx, &prof->d_tf_mw, tfm.data(), tfm.size() * 4));
      prof->mwC = mc; prof->mwW = mw;
    }
    prof->vit.xw_E_move = om->xw[0][0]; prof->vit.xw_E_loop = om->xw[0][1];
    prof->vit.base_w = om->base_w; prof->vit.ddbound_w = om->ddbound_w; prof->vit.scale_w = om->scale_w;
    P7G_CHECK(build_vit_packed(prof, om, Q8));
    prof->scale_w = om->scale_w;
    prof->fwd.xf_E_move = om->xf[0][0]; prof->fwd.xf_E_loop = om->xf[0][1];
  }
  P7G_CHECK(arena_commit(ctx, arena, &prof->d_block));
  P7G_CUDA(cudaStreamSynchronize(ctx->stream));      // the tables are complete before any stream of the context reads them
  *ret_prof = prof;
  return P7GPU_OK;
}

void p7gpu_profile_destroy(p7gpu_profile *prof)
{
  if (!prof) return;
  p7gpu_ctx *ctx = prof->ctx;
  cudaSetDevice(ctx->device);
  sync_all_lanes(ctx);                         // nothing that reads the tables is still running, on any of the context's streams
  if (prof->d_block) cudaFreeAsync(prof->d_block, ctx->stream);      // every table lives inside this one block (arena_commit)
  if (prof->d_len)   cudaFreeAsync(prof->d_len, ctx->stream);
  delete prof;
}

int p7gpu_profile_M(const p7gpu_profile *prof) { return prof ? prof->M : 0; }

// ---------------------------------------------------------------------------------------------------
// Target database
// ---------------------------------------------------------------------------------------------------
__global__ void iota_kernel(int32_t *v, int64_t n)
{
  int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (int32_t) i;
}

}  // extern "C"
// Work order of a resident database: database order for all but the last targets, which are taken longest first.
// Only the end of a launch needs the longest-first rule (the kernels draw targets from one device counter, so a long
// target drawn last is the launch's tail); everywhere else database order keeps the per-target accesses to offsets,
// lengths, scores and statuses sequential — a fully length-sorted order scatters them and doubled the SSV kernel's DRAM
// traffic (ncu r1f: 8.0 GB per launch for 3.7 GB of algorithmic bytes).
int p7g::build_order(p7gpu_ctx *ctx, p7gpu_seqdb *db)
{
  const int64_t n = db->nseq;
  if (n == 0) return P7GPU_OK;
  if (n > 0x7fffffffLL) { set_error("seqdb: more than 2^31-1 targets per device"); return P7GPU_EINVAL; }
  const int64_t tail = std::min<int64_t>(n, std::max<int64_t>(65536, n / 32)), n0 = n - tail;
  int32_t *d_keys_out = nullptr, *d_iota = nullptr; void *d_tmp = nullptr; size_t tmp_bytes = 0;
  P7G_CUDA(cudaMalloc(&db->d_order, sizeof(int32_t) * n));
  P7G_CUDA(cudaMalloc(&d_keys_out, sizeof(int32_t) * tail));
  P7G_CUDA(cudaMalloc(&d_iota, sizeof(int32_t) * n));
  iota_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>(d_iota, n);
  P7G_CUDA(cudaGetLastError());
  if (n0) P7G_CUDA(cudaMemcpyAsync(db->d_order, d_iota, sizeof(int32_t) * n0, cudaMemcpyDeviceToDevice, ctx->stream));
  P7G_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, db->d_len + n0, d_keys_out, d_iota + n0, db->d_order + n0, (int) tail, 0, 32, ctx->stream));
  P7G_CUDA(cudaMalloc(&d_tmp, tmp_bytes));
  P7G_CUDA(cub::DeviceRadixSort::SortPairsDescending(d_tmp, tmp_bytes, db->d_len + n0, d_keys_out, d_iota + n0, db->d_order + n0, (int) tail, 0, 32, ctx->stream));
  P7G_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_tmp); cudaFree(d_keys_out); cudaFree(d_iota);
  return P7GPU_OK;
}
extern "C" {

int p7gpu_seqdb_upload(p7gpu_ctx *ctx, const uint8_t *residues, int64_t nbytes, const int64_t *offsets,
                       const int32_t *lengths, int64_t nseq, p7gpu_seqdb **ret_db)
{
  if (!ctx || !ret_db || nseq < 0 || (nseq > 0 && (!residues || !offsets || !lengths))) return P7GPU_EINVAL;
  *ret_db = nullptr;
  P7G_CUDA(cudaSetDevice(ctx->device));
  int32_t Lmax = 0; int64_t nres = 0;
  for (int64_t s = 0; s < nseq; s++) {
    if (lengths[s] < 0 || offsets[s] < 1 || offsets[s] + lengths[s] >= nbytes) {
      set_error("seqdb_upload: target %lld (offset %lld, length %d) does not fit the packed array of %lld bytes",
                (long long) s, (long long) offsets[s], lengths[s], (long long) nbytes);
      return P7GPU_EINVAL;
    }
    Lmax = std::max(Lmax, lengths[s]); nres += lengths[s];
  }
  p7gpu_seqdb *db = new p7gpu_seqdb();
  db->ctx = ctx; db->nseq = nseq; db->nbytes = nbytes; db->nres = nres; db->max_length = Lmax; db->owns = true;
  const size_t padded = ((size_t) nbytes + 15) / 16 * 16 + 32;
  uint8_t *d_res = nullptr; int64_t *d_off = nullptr; int32_t *d_len = nullptr;
  // A handful of targets (the one-target drop-in calls, the reference's calibration loops through them: evalues.c:264,337,442)
  // lives in the stream-ordered pool and is searched in database order: no cudaMalloc/cudaFree round trips, no sort.
  const bool small = nseq <= 4096 && nbytes <= ((int64_t) 8 << 20);
  db->pooled = small;
  if (small) {
    P7G_CUDA(cudaMallocAsync((void **) &d_res, padded, ctx->stream));
    P7G_CUDA(cudaMallocAsync((void **) &d_off, sizeof(int64_t) * std::max<int64_t>(nseq, 1), ctx->stream));
    P7G_CUDA(cudaMallocAsync((void **) &d_len, sizeof(int32_t) * std::max<int64_t>(nseq, 1), ctx->stream));
  } else {
    P7G_CUDA(cudaMalloc(&d_res, padded));
    P7G_CUDA(cudaMalloc(&d_off, sizeof(int64_t) * std::max<int64_t>(nseq, 1)));
    P7G_CUDA(cudaMalloc(&d_len, sizeof(int32_t) * std::max<int64_t>(nseq, 1)));
  }
  P7G_CUDA(cudaMemsetAsync(d_res, 0xff, padded, ctx->stream));
  if (nbytes) P7G_CUDA(cudaMemcpyAsync(d_res, residues, (size_t) nbytes, cudaMemcpyHostToDevice, ctx->stream));
  if (nseq) {
    P7G_CUDA(cudaMemcpyAsync(d_off, offsets, sizeof(int64_t) * nseq, cudaMemcpyHostToDevice, ctx->stream));
    P7G_CUDA(cudaMemcpyAsync(d_len, lengths, sizeof(int32_t) * nseq, cudaMemcpyHostToDevice, ctx->stream));
  }
  db->d_res = d_res; db->d_off = d_off; db->d_len = d_len;
  P7G_CUDA(cudaStreamSynchronize(ctx->stream));
  if (!small) {
    int st = build_order(ctx, db);
    if (st != P7GPU_OK) { p7gpu_seqdb_destroy(db); return st; }
  }
  *ret_db = db;
  return P7GPU_OK;
}

struct WidenIt {                                                       // int32 lengths read as int64 (input iterator for cub::DeviceReduce)
  using iterator_category = std::random_access_iterator_tag; using value_type = int64_t; using difference_type = ptrdiff_t;
  using pointer = const int64_t *; using reference = int64_t;
  const int32_t *p;
  __host__ __device__ int64_t operator*() const { return (int64_t) *p; }
  __host__ __device__ int64_t operator[](ptrdiff_t i) const { return (int64_t) p[i]; }
  __host__ __device__ WidenIt operator+(ptrdiff_t i) const { return WidenIt{p + i}; }
  __host__ __device__ WidenIt &operator+=(ptrdiff_t i) { p += i; return *this; }
  __host__ __device__ WidenIt &operator++() { ++p; return *this; }
};

int p7gpu_seqdb_adopt_device(p7gpu_ctx *ctx, const void *d_residues, int64_t nbytes, const void *d_offsets,
                             const void *d_lengths, int64_t nseq, int32_t max_length, p7gpu_seqdb **ret_db)
{
  if (!ctx || !ret_db || !d_residues || !d_offsets || !d_lengths || nseq < 0) return P7GPU_EINVAL;
  if (((uintptr_t) d_residues & 15) != 0) { set_error("seqdb_adopt_device: residue array must be 16-byte aligned"); return P7GPU_EINVAL; }
  *ret_db = nullptr;
  P7G_CUDA(cudaSetDevice(ctx->device));
  p7gpu_seqdb *db = new p7gpu_seqdb();
  db->ctx = ctx; db->nseq = nseq; db->nbytes = nbytes; db->nres = 0; db->max_length = max_length; db->owns = false;
  db->d_res = (const uint8_t *) d_residues; db->d_off = (const int64_t *) d_offsets; db->d_len = (const int32_t *) d_lengths;
  if (nseq) {                                                          // pli->nres = sum of the lengths (p7_pli_NewSeq); gaps between targets are layout
    size_t tmp_bytes = 0; int64_t *d_sum = nullptr; void *d_tmp = nullptr;
    WidenIt it{db->d_len};
    P7G_CUDA(cub::DeviceReduce::Sum(nullptr, tmp_bytes, it, (int64_t *) nullptr, (int) nseq, ctx->stream));
    P7G_CUDA(cudaMalloc(&d_tmp, tmp_bytes + 16));
    d_sum = (int64_t *) ((char *) d_tmp + (tmp_bytes + 7) / 8 * 8);
    cudaError_t e = cub::DeviceReduce::Sum(d_tmp, tmp_bytes, it, d_sum, (int) nseq, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&db->nres, d_sum, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_tmp);
    if (e != cudaSuccess) { set_error("seqdb_adopt_device: length sum failed: %s", cudaGetErrorString(e)); delete db; return P7GPU_ESYS; }
    ctx->launches += 1;
  }
  int st = build_order(ctx, db);
  if (st != P7GPU_OK) { p7gpu_seqdb_destroy(db); return st; }
  *ret_db = db;
  return P7GPU_OK;
}

void p7gpu_seqdb_destroy(p7gpu_seqdb *db)
{
  if (!db) return;
  cudaSetDevice(db->ctx->device);
  cudaStreamSynchronize(db->ctx->stream);
  if (db->pooled) {
    cudaStream_t s = db->ctx->stream;
    if (db->d_res) cudaFreeAsync((void *) db->d_res, s);
    if (db->d_off) cudaFreeAsync((void *) db->d_off, s);
    if (db->d_len) cudaFreeAsync((void *) db->d_len, s);
    if (db->d_order) cudaFreeAsync(db->d_order, s);
  } else {
    if (db->owns) { cudaFree((void *) db->d_res); cudaFree((void *) db->d_off); cudaFree((void *) db->d_len); }
    cudaFree(db->d_order);
  }
  delete db;
}
int64_t p7gpu_seqdb_nseq(const p7gpu_seqdb *db) { return db ? db->nseq : 0; }
int64_t p7gpu_seqdb_nres(const p7gpu_seqdb *db) { return db ? db->nres : 0; }

// ---------------------------------------------------------------------------------------------------
// Batched stages
// ---------------------------------------------------------------------------------------------------
struct BatchBufs { int64_t *d_idx; float *d_sc; int32_t *d_st; int64_t *d_list; int64_t n; };

// carve the device scratch: [idx n][sc n][status n][noresult list 2n]
static int setup_batch(p7gpu_ctx *ctx, const p7gpu_seqdb *db, const int64_t *idx, int64_t n_in, BatchBufs *b)
{
  const int64_t n = idx ? n_in : db->nseq;
  b->n = n;
  const size_t a = ((size_t) std::max<int64_t>(n, 1) * 8 + 255) / 256 * 256;
  P7G_CHECK(ensure_scratch(ctx, a * 5));
  char *base = (char *) ctx->d_scratch;
  b->d_idx = idx ? (int64_t *) base : nullptr;
  b->d_sc  = (float *)   (base + a);
  b->d_st  = (int32_t *) (base + 2 * a);
  b->d_list = (int64_t *) (base + 3 * a);
  if (idx && n) {
    for (int64_t j = 0; j < n; j++)
      if (idx[j] < 0 || idx[j] >= db->nseq) { set_error("batch: idx[%lld]=%lld out of range", (long long) j, (long long) idx[j]); return P7GPU_EINVAL; }
    P7G_CUDA(cudaMemcpyAsync(base, idx, sizeof(int64_t) * n, cudaMemcpyHostToDevice, ctx->stream));
  }
  return P7GPU_OK;
}

static int fetch_batch(p7gpu_ctx *ctx, const BatchBufs &b, float *sc, int32_t *status)
{
  if (b.n) {
    if (sc)     P7G_CUDA(cudaMemcpyAsync(sc, b.d_sc, sizeof(float) * b.n, cudaMemcpyDeviceToHost, ctx->stream));
    if (status) P7G_CUDA(cudaMemcpyAsync(status, b.d_st, sizeof(int32_t) * b.n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  P7G_CUDA(cudaStreamSynchronize(ctx->stream));
  return P7GPU_OK;
}

int p7gpu_ssvfilter_batch(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db,
                          const int64_t *idx, int64_t n, float *sc, int32_t *status)
{
  if (!ctx || !prof || !db) return P7GPU_EINVAL;
  P7G_CUDA(cudaSetDevice(ctx->device));
  BatchBufs b;
  P7G_CHECK(setup_batch(ctx, db, idx, n, &b));
  P7G_CHECK(launch_ssv(ctx, 0, prof, db->view(), b.d_idx, b.n, b.d_sc, b.d_st, nullptr, nullptr, nullptr, nullptr));
  return fetch_batch(ctx, b, sc, status);
}

int p7gpu_msvfilter_batch(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db,
                          const int64_t *idx, int64_t n, float *sc, int32_t *status)
{
  if (!ctx || !prof || !db) return P7GPU_EINVAL;
  P7G_CUDA(cudaSetDevice(ctx->device));
  BatchBufs b;
  P7G_CHECK(setup_batch(ctx, db, idx, n, &b));
  unsigned int *d_nores = ctx->lanes[0].counters + CTR_NORES_COUNT;
  P7G_CHECK(launch_ssv(ctx, 0, prof, db->view(), b.d_idx, b.n, b.d_sc, b.d_st, b.d_list, d_nores, nullptr, nullptr));
  P7G_CHECK(launch_msv_j(ctx, 0, prof, db->view(), b.d_list, d_nores, b.n, b.d_sc, b.d_st, nullptr, nullptr));
  return fetch_batch(ctx, b, sc, status);
}

int p7gpu_msvfilter_device(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db, double F1,
                           void *d_sc, void *d_status, void *d_pass_list, void *d_pass_count)
{
  if (!ctx || !prof || !db || !d_sc || !d_status || ((d_pass_list == nullptr) != (d_pass_count == nullptr))) return P7GPU_EINVAL;
  P7G_CUDA(cudaSetDevice(ctx->device));
  const int64_t n = db->nseq;
  P7G_CHECK(ensure_len_table(prof, db->max_length, F1));
  P7G_CHECK(ensure_scratch2(ctx, (size_t) std::max<int64_t>(n, 1) * 16));
  unsigned int *d_nores = ctx->lanes[0].counters + CTR_NORES_COUNT;
  if (d_pass_count) P7G_CUDA(cudaMemsetAsync(d_pass_count, 0, sizeof(unsigned int), ctx->stream));
  P7G_CHECK(launch_ssv(ctx, 0, prof, db->view(), nullptr, n, (float *) d_sc, (int32_t *) d_status, (int64_t *) ctx->d_scratch2, d_nores,
                       (int32_t *) d_pass_list, (unsigned int *) d_pass_count));
  P7G_CHECK(launch_msv_j(ctx, 0, prof, db->view(), (int64_t *) ctx->d_scratch2, d_nores, n, (float *) d_sc, (int32_t *) d_status,
                         (int32_t *) d_pass_list, (unsigned int *) d_pass_count));
  return P7GPU_OK;
}

int p7gpu_profile_set_biasfilter(p7gpu_profile *prof, const float *t6, const float *pi3, const float *eo)
{
  if (!prof || !t6 || !pi3 || !eo) return P7GPU_EINVAL;
  memcpy(prof->bias.t, t6, sizeof(float) * 6);
  memcpy(prof->bias.pi, pi3, sizeof(float) * 3);
  memcpy(prof->bias.eo, eo, sizeof(float) * P7GPU_KP * 2);
  prof->have_bias = true;
  return P7GPU_OK;
}

int p7gpu_biasfilter_batch(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db,
                           const int64_t *idx, int64_t n, float *filtersc)
{
  if (!ctx || !prof || !db) return P7GPU_EINVAL;
  P7G_CUDA(cudaSetDevice(ctx->device));
  BatchBufs b;
  P7G_CHECK(setup_batch(ctx, db, idx, n, &b));
  int32_t *d_unsure = (int32_t *) b.d_list;                                   // [n] work items + one counter behind them
  unsigned int *d_ucount = (unsigned int *) (d_unsure + std::max<int64_t>(b.n, 1));
  P7G_CHECK(launch_bias(ctx, 0, prof, db->view(), b.d_idx, b.n, b.d_sc, d_unsure, d_ucount));
  P7G_CHECK(fix_unsure_filtersc(ctx, prof, db->view(), b.d_idx, b.d_sc, d_unsure, d_ucount, nullptr, nullptr));
  return fetch_batch(ctx, b, filtersc, nullptr);
}

int p7gpu_vitfilter_batch(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db,
                          const int64_t *idx, int64_t n, float *sc, int32_t *status)
{
  if (!ctx || !prof || !db) return P7GPU_EINVAL;
  if (!prof->d_rw) { set_error("vitfilter: profile was uploaded without the Viterbi part"); return P7GPU_EINVAL; }
  P7G_CUDA(cudaSetDevice(ctx->device));
  BatchBufs b;
  P7G_CHECK(setup_batch(ctx, db, idx, n, &b));
  P7G_CHECK(launch_vit(ctx, 0, prof, db->view(), b.d_idx, b.n, b.d_sc, b.d_st));
  return fetch_batch(ctx, b, sc, status);
}

// xmx blocks: target j gets (L_j+1)*6 floats at xmx + xmx_offsets[j]
static int64_t xmx_extent(const p7gpu_seqdb *, const int64_t *xmx_offsets, int64_t n, const int32_t *h_len_of_j)
{
  int64_t ext = 0;
  for (int64_t j = 0; j < n; j++) ext = std::max(ext, xmx_offsets[j] + (int64_t) (h_len_of_j[j] + 1) * P7GPU_NXCELLS);
  return ext;
}

static int lengths_of(p7gpu_ctx *ctx, const p7gpu_seqdb *db, const int64_t *idx, int64_t n, std::vector<int32_t> &out)
{
  // host needs the lengths of the selected targets to size the specials; fetch them from the device copy
  std::vector<int32_t> all;
  out.resize(n);
  if (!idx) { P7G_CUDA(cudaMemcpy(out.data(), db->d_len, sizeof(int32_t) * n, cudaMemcpyDeviceToHost)); return P7GPU_OK; }
  for (int64_t j = 0; j < n; j++)
    P7G_CUDA(cudaMemcpyAsync(&out[j], db->d_len + idx[j], sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  P7G_CUDA(cudaStreamSynchronize(ctx->stream));
  return P7GPU_OK;
}

// dense offsets for callers that pass no specials buffer but run the exact-order kernels (which always write specials)
static void dense_xmx_offsets(const std::vector<int32_t> &lens, std::vector<int64_t> &off, int64_t *ext)
{
  off.resize(lens.size()); int64_t e = 0;
  for (size_t j = 0; j < lens.size(); j++) { off[j] = e; e += (int64_t) (lens[j] + 1) * P7GPU_NXCELLS; }
  *ext = e;
}

int p7gpu_forward_batch(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db,
                        const int64_t *idx, int64_t n, float *sc, int32_t *status,
                        float *xmx, const int64_t *xmx_offsets, float *totscale)
{
  if (!ctx || !prof || !db) return P7GPU_EINVAL;
  if (!prof->d_rf) { set_error("forward: profile was uploaded without the Forward part"); return P7GPU_EINVAL; }
  if (xmx && !xmx_offsets) return P7GPU_EINVAL;
  P7G_CUDA(cudaSetDevice(ctx->device));
  BatchBufs b;
  P7G_CHECK(setup_batch(ctx, db, idx, n, &b));
  const bool exact = ctx->fb_exact != 0;
  float *d_xmx = nullptr; int64_t *d_xoff = nullptr; float *d_tot = nullptr; int64_t ext = 0;
  std::vector<int32_t> lens; std::vector<int64_t> own_off;
  const size_t nb = (size_t) std::max<int64_t>(b.n, 1);
  if (xmx || exact) P7G_CHECK(lengths_of(ctx, db, idx, b.n, lens));
  if (xmx) ext = xmx_extent(db, xmx_offsets, b.n, lens.data());
  else if (exact) { dense_xmx_offsets(lens, own_off, &ext); xmx_offsets = own_off.data(); }
  const bool dev_x = xmx || exact;
  const size_t off_bytes = (nb * 8 + 255) / 256 * 256, tot_bytes = (nb * 4 + 255) / 256 * 256;
  P7G_CHECK(ensure_scratch2(ctx, off_bytes + tot_bytes + (size_t) ext * 4 + 256));
  d_xoff = (int64_t *) ctx->d_scratch2; d_tot = (float *) ((char *) ctx->d_scratch2 + off_bytes);
  if (dev_x) {
    d_xmx = (float *) ((char *) ctx->d_scratch2 + off_bytes + tot_bytes);
    P7G_CUDA(cudaMemcpyAsync(d_xoff, xmx_offsets, sizeof(int64_t) * b.n, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (exact) P7G_CHECK(launch_fbx_fwd(ctx, 0, prof, db->view(), b.d_idx, b.n, b.d_sc, b.d_st, d_xmx, d_xoff, d_tot));
  else       P7G_CHECK(launch_fwd(ctx, 0, prof, db->view(), b.d_idx, b.n, b.d_sc, b.d_st, d_xmx, xmx ? d_xoff : nullptr, d_tot));
  if (xmx && ext) P7G_CUDA(cudaMemcpyAsync(xmx, d_xmx, sizeof(float) * ext, cudaMemcpyDeviceToHost, ctx->stream));
  if (totscale && b.n) P7G_CUDA(cudaMemcpyAsync(totscale, d_tot, sizeof(float) * b.n, cudaMemcpyDeviceToHost, ctx->stream));
  P7G_CHECK(fetch_batch(ctx, b, sc, status));
  if (exact && xmx)                                                   // scores from the specials with the host libm (fwdback.c:434,:463)
    for (int64_t j = 0; j < b.n; j++) {
      float t, s; int st;
      fbx_host_fwd_score(xmx + xmx_offsets[j], lens[j], prof->h_len[lens[j]].pmove, &t, &s, &st);
      if (sc) sc[j] = s; if (status) status[j] = st; if (totscale) totscale[j] = t;
    }
  return P7GPU_OK;
}

static int backward_batch_impl(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db,
                               const int64_t *idx, int64_t n, const float *fwd_xmx, const int64_t *xmx_offsets,
                               float *sc, int32_t *status, float *bck_xmx, float *bck_totscale, int32_t *own_scales)
{
  if (!ctx || !prof || !db || !fwd_xmx || !xmx_offsets) return P7GPU_EINVAL;
  if (!prof->d_rf) { set_error("backward: profile was uploaded without the Forward part"); return P7GPU_EINVAL; }
  P7G_CUDA(cudaSetDevice(ctx->device));
  BatchBufs b;
  P7G_CHECK(setup_batch(ctx, db, idx, n, &b));
  std::vector<int32_t> lens;
  P7G_CHECK(lengths_of(ctx, db, idx, b.n, lens));
  const int64_t ext = xmx_extent(db, xmx_offsets, b.n, lens.data());
  const size_t nb = (size_t) std::max<int64_t>(b.n, 1);
  const size_t off_bytes = (nb * 8 + 255) / 256 * 256, tot_bytes = (nb * 4 + 255) / 256 * 256;
  const size_t x_bytes = ((size_t) ext * 4 + 255) / 256 * 256;
  P7G_CHECK(ensure_scratch2(ctx, off_bytes + 2 * tot_bytes + 2 * x_bytes + 256));
  int64_t *d_xoff = (int64_t *) ctx->d_scratch2;
  float *d_tot = (float *) ((char *) ctx->d_scratch2 + off_bytes);
  int32_t *d_own = (int32_t *) ((char *) ctx->d_scratch2 + off_bytes + tot_bytes);
  float *d_fx  = (float *) ((char *) ctx->d_scratch2 + off_bytes + 2 * tot_bytes);
  float *d_bx  = (float *) ((char *) ctx->d_scratch2 + off_bytes + 2 * tot_bytes + x_bytes);
  P7G_CUDA(cudaMemcpyAsync(d_xoff, xmx_offsets, sizeof(int64_t) * b.n, cudaMemcpyHostToDevice, ctx->stream));
  if (ext) P7G_CUDA(cudaMemcpyAsync(d_fx, fwd_xmx, sizeof(float) * ext, cudaMemcpyHostToDevice, ctx->stream));
  if (ctx->fb_exact) P7G_CHECK(launch_fbx_bwd(ctx, 0, prof, db->view(), b.d_idx, b.n, d_fx, d_xoff, b.d_sc, b.d_st, d_bx, d_tot, d_own));
  else               P7G_CHECK(launch_bwd(ctx, 0, prof, db->view(), b.d_idx, b.n, d_fx, d_xoff, b.d_sc, b.d_st, d_bx, d_tot, d_own));
  if (bck_xmx && ext) P7G_CUDA(cudaMemcpyAsync(bck_xmx, d_bx, sizeof(float) * ext, cudaMemcpyDeviceToHost, ctx->stream));
  if (bck_totscale && b.n) P7G_CUDA(cudaMemcpyAsync(bck_totscale, d_tot, sizeof(float) * b.n, cudaMemcpyDeviceToHost, ctx->stream));
  if (own_scales && b.n) P7G_CUDA(cudaMemcpyAsync(own_scales, d_own, sizeof(int32_t) * b.n, cudaMemcpyDeviceToHost, ctx->stream));
  P7G_CHECK(fetch_batch(ctx, b, sc, status));
  if (ctx->fb_exact && bck_xmx)                                        // fwdback.c:553,:689,:735 with the host libm
    for (int64_t j = 0; j < b.n; j++) {
      float t, s; int st;
      fbx_host_bwd_score(bck_xmx + xmx_offsets[j], lens[j], &t, &s, &st);
      if (sc) sc[j] = s; if (status) status[j] = st; if (bck_totscale) bck_totscale[j] = t;
    }
  return P7GPU_OK;
}

int p7gpu_backward_batch(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db,
                         const int64_t *idx, int64_t n, const float *fwd_xmx, const int64_t *xmx_offsets,
                         float *sc, int32_t *status, float *bck_xmx, float *bck_totscale)
{
  return backward_batch_impl(ctx, prof, db, idx, n, fwd_xmx, xmx_offsets, sc, status, bck_xmx, bck_totscale, nullptr);
}

// The region trigger of p7_domaindef_ByPosteriorHeuristics for targets whose specials are on the host (decode.cu).
int p7gpu_region_trigger_batch(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db,
                               const int64_t *idx, int64_t n, const float *fwd_xmx, const float *bck_xmx,
                               const int64_t *xmx_offsets, const int32_t *bck_own_scales, float *mocc_max, int32_t *flag)
{
  if (!ctx || !prof || !db || !fwd_xmx || !bck_xmx || !xmx_offsets || !mocc_max || !flag) return P7GPU_EINVAL;
  P7G_CUDA(cudaSetDevice(ctx->device));
  BatchBufs b;
  P7G_CHECK(setup_batch(ctx, db, idx, n, &b));
  if (b.n == 0) return P7GPU_OK;
  std::vector<int32_t> lens;
  P7G_CHECK(lengths_of(ctx, db, idx, b.n, lens));
  const int64_t ext = xmx_extent(db, xmx_offsets, b.n, lens.data());
  const size_t nb = (size_t) b.n;
  const size_t off_bytes = (nb * 8 + 255) / 256 * 256, w_bytes = (nb * 4 + 255) / 256 * 256;
  const size_t x_bytes = ((size_t) ext * 4 + 255) / 256 * 256;
  P7G_CHECK(ensure_scratch2(ctx, off_bytes + 3 * w_bytes + 2 * x_bytes + 256));
  char *base = (char *) ctx->d_scratch2;
  int64_t *d_xoff = (int64_t *) base;
  int32_t *d_own = (int32_t *) (base + off_bytes), *d_flag = (int32_t *) (base + off_bytes + w_bytes);
  float *d_mocc = (float *) (base + off_bytes + 2 * w_bytes);
  float *d_fx = (float *) (base + off_bytes + 3 * w_bytes), *d_bx = (float *) (base + off_bytes + 3 * w_bytes + x_bytes);
  cudaStream_t st = ctx->stream;
  P7G_CUDA(cudaMemcpyAsync(d_xoff, xmx_offsets, 8 * nb, cudaMemcpyHostToDevice, st));
  if (bck_own_scales) P7G_CUDA(cudaMemcpyAsync(d_own, bck_own_scales, 4 * nb, cudaMemcpyHostToDevice, st));
  P7G_CUDA(cudaMemcpyAsync(d_fx, fwd_xmx, 4 * (size_t) ext, cudaMemcpyHostToDevice, st));
  P7G_CUDA(cudaMemcpyAsync(d_bx, bck_xmx, 4 * (size_t) ext, cudaMemcpyHostToDevice, st));
  P7G_CHECK(launch_region_trigger(ctx, 0, prof, db->view(), b.d_idx, b.n, d_fx, d_bx, d_xoff, bck_own_scales ? d_own : nullptr,
                                  d_mocc, d_flag, nullptr));
  P7G_CUDA(cudaMemcpyAsync(mocc_max, d_mocc, 4 * nb, cudaMemcpyDeviceToHost, st));
  P7G_CUDA(cudaMemcpyAsync(flag, d_flag, 4 * nb, cudaMemcpyDeviceToHost, st));
  P7G_CUDA(cudaStreamSynchronize(st));
  return P7GPU_OK;
}

// ---------------------------------------------------------------------------------------------------
// Signature-level wrappers: one target per call (impl_sse.h:309-344)
// ---------------------------------------------------------------------------------------------------
struct OneSeq { p7gpu_seqdb *db = nullptr; ~OneSeq() { p7gpu_seqdb_destroy(db); } };

static int one_seq_db(p7gpu_ctx *ctx, const uint8_t *dsq, int L, OneSeq *one)
{
  if (!dsq || L < 0) return P7GPU_EINVAL;
  int64_t off = 1; int32_t len = L;
  return p7gpu_seqdb_upload(ctx, dsq, (int64_t) L + 2, &off, &len, 1, &one->db);
}

int p7gpu_SSVFilter(p7gpu_ctx *ctx, const uint8_t *dsq, int L, const p7gpu_profile *om, float *ret_sc)
{
  OneSeq one; int32_t st = 0; float sc = 0;
  P7G_CHECK(one_seq_db(ctx, dsq, L, &one));
  P7G_CHECK(p7gpu_ssvfilter_batch(ctx, om, one.db, nullptr, 1, &sc, &st));
  if (st != P7GPU_ENORESULT || std::isinf(sc)) *ret_sc = sc;
  return st;
}
int p7gpu_MSVFilter(p7gpu_ctx *ctx, const uint8_t *dsq, int L, const p7gpu_profile *om, float *ret_sc)
{
  OneSeq one; int32_t st = 0;
  P7G_CHECK(one_seq_db(ctx, dsq, L, &one));
  P7G_CHECK(p7gpu_msvfilter_batch(ctx, om, one.db, nullptr, 1, ret_sc, &st));
  return st;
}
int p7gpu_ViterbiFilter(p7gpu_ctx *ctx, const uint8_t *dsq, int L, const p7gpu_profile *om, float *ret_sc)
{
  OneSeq one; int32_t st = 0;
  P7G_CHECK(one_seq_db(ctx, dsq, L, &one));
  P7G_CHECK(p7gpu_vitfilter_batch(ctx, om, one.db, nullptr, 1, ret_sc, &st));
  return st;
}
// the one-target parsers answer in the reference's own order of operations (bit-identical) unless option "fb_fast" is set
struct ExactScope { p7gpu_ctx *c; int saved; explicit ExactScope(p7gpu_ctx *ctx) : c(ctx), saved(ctx->fb_exact) { if (!ctx->fb_fast) ctx->fb_exact = 1; } ~ExactScope() { c->fb_exact = saved; } };

int p7gpu_ForwardParser(p7gpu_ctx *ctx, const uint8_t *dsq, int L, const p7gpu_profile *om,
                        float *fwd_xmx, float *fwd_totscale, float *opt_sc)
{
  OneSeq one; int32_t st = 0; float sc = 0; int64_t xoff = 0;
  if (!ctx) return P7GPU_EINVAL;
  P7G_CHECK(one_seq_db(ctx, dsq, L, &one));
  ExactScope ex(ctx);
  std::vector<float> tmp;
  if (!fwd_xmx && ctx->fb_exact) { tmp.resize((size_t) (L + 1) * P7GPU_NXCELLS); fwd_xmx = tmp.data(); }   // scores come from the specials
  P7G_CHECK(p7gpu_forward_batch(ctx, om, one.db, nullptr, 1, &sc, &st, fwd_xmx, fwd_xmx ? &xoff : nullptr, fwd_totscale));
  if (opt_sc) *opt_sc = sc;
  return st;
}
int p7gpu_BackwardParser2(p7gpu_ctx *ctx, const uint8_t *dsq, int L, const p7gpu_profile *om,
                          const float *fwd_xmx, float *bck_xmx, float *bck_totscale, float *opt_sc, int *opt_own_scales)
{
  OneSeq one; int32_t st = 0, own = 0; float sc = 0; int64_t xoff = 0;
  if (!ctx) return P7GPU_EINVAL;
  P7G_CHECK(one_seq_db(ctx, dsq, L, &one));
  ExactScope ex(ctx);
  std::vector<float> tmp;
  if (!bck_xmx && ctx->fb_exact) { tmp.resize((size_t) (L + 1) * P7GPU_NXCELLS); bck_xmx = tmp.data(); }
  P7G_CHECK(backward_batch_impl(ctx, om, one.db, nullptr, 1, fwd_xmx, &xoff, &sc, &st, bck_xmx, bck_totscale, &own));
  if (opt_sc) *opt_sc = sc;
  if (opt_own_scales) *opt_own_scales = own;
  return st;
}
int p7gpu_BackwardParser(p7gpu_ctx *ctx, const uint8_t *dsq, int L, const p7gpu_profile *om,
                         const float *fwd_xmx, float *bck_xmx, float *bck_totscale, float *opt_sc)
{
  return p7gpu_BackwardParser2(ctx, dsq, L, om, fwd_xmx, bck_xmx, bck_totscale, opt_sc, nullptr);
}

float p7gpu_nullsc(int L) { return host_nullsc(L); }

// esl_gumbel_surv of (usc - nullsc)/ln2 (p7_pipeline.c:691-692; easel/esl_gumbel.c:129-138)
double p7gpu_msv_pvalue(float usc, float nullsc, float mmu, float mlambda)
{
  return host_filter_pvalue_gumbel(usc, nullsc, mmu, mlambda);
}

}  // extern "C"
