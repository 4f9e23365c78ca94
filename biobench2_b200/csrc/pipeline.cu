// pipeline.cu — the batched p7_Pipeline() driver (include/p7gpu.h §7) and the streamed-database MSV pass.
//
// Stage-major restatement of p7_Pipeline() (HMMER/src/p7_pipeline.c:665-735) for a whole database:
//   A. MSV filter on every target (SSV kernel + J-state kernel); the P-value test of :691-693 is decided inside the
//      kernels from a host-built per-length table (smallest passing xJ byte), survivors are compacted on the device.
//   B. bias filter score for the MSV survivors (kernel), P-value test on the host with the host libm (:697-703);
//   C. Viterbi filter for those whose P still exceeds F2 (:716-722);   D. Forward parser (:727-730);
//   E. Backward parser for the Forward survivors (:734-735), whose specials go back to the host.
// Every score->P conversion after the MSV stage runs on the host, in the reference's own float/double expressions,
// on the <= 2% of targets that get that far: pass/fail decisions therefore never depend on device libm.
#include "p7gpu_internal.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <math_constants.h>
#include <iterator>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <sys/mman.h>

// Host array that grows without zero-filling, on 2 MB-aligned storage advised for transparent huge pages: a --max result
// holds GBs of specials, and first-touch page faults on 4 KB pages (not PCIe) bounded the copy at 2 GB/s.  A result hands
// its arrays back to the context when it is destroyed, so a second search faults nothing.
struct RawFloats {
  float *p = nullptr; size_t n = 0, cap = 0;
  bool resize(size_t m) {
    if (m > cap) {
      const size_t c = (std::max(m, cap + cap / 2) * sizeof(float) + ((size_t) 2 << 20) - 1) & ~(((size_t) 2 << 20) - 1);
      void *q = nullptr;
      if (posix_memalign(&q, (size_t) 2 << 20, c) != 0) return false;
      madvise(q, c, MADV_HUGEPAGE);
      if (n) memcpy(q, p, n * sizeof(float));
      free(p); p = (float *) q; cap = c / sizeof(float);
    }
    n = m; return true;
  }
  float *data() const { return p; }
  void swap(RawFloats &o) { std::swap(p, o.p); std::swap(n, o.n); std::swap(cap, o.cap); }
  RawFloats() = default;
  RawFloats(const RawFloats &) = delete; RawFloats &operator=(const RawFloats &) = delete;
  ~RawFloats() { free(p); }
};

struct p7gpu_pipeline_result {
  p7gpu_ctx *ctx = nullptr;                       // the MSV-survivor buffer goes back to this context's cache on destroy
  p7gpu_pipeline_counters counters;
  std::vector<p7gpu_survivor> survivors;
  RawFloats fwd_xmx, bck_xmx, dec;                // dec: opts.decode blocks (btot, etot, mocc, fwd SCALE per row)
  std::vector<p7gpu_msv_survivor> msv_pass;       // only when opts->stop_after_msv
  uint64_t stage_nres[6] = {0, 0, 0, 0, 0, 0};    // residues seen by the MSV, bias, Viterbi, Forward, Backward stages and the exact-order Forward recomputation
  uint64_t n_region_flagged = 0;                  // survivors whose specials were returned
};

namespace p7g {

// P7GPU_TRACE=1: wall-clock checkpoints of the batched entry points on stderr
struct Trace {
  bool on; std::chrono::steady_clock::time_point t0, last;
  Trace() : on(getenv("P7GPU_TRACE") != nullptr) { t0 = last = std::chrono::steady_clock::now(); }
  void mark(const char *what) {
    if (!on) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[p7gpu] %-28s +%8.3f ms  (total %8.3f ms)\n", what,
            std::chrono::duration<double, std::milli>(now - last).count(), std::chrono::duration<double, std::milli>(now - t0).count());
    last = now;
  }
};

// the per-survivor host arithmetic (P-values in the reference's float/double expressions) is embarrassingly parallel:
// the reference evaluates it inside its worker threads (hmmsearch.c:1372-1424), so do we
template <class F> static void parallel_for(size_t n, F f)
{
  unsigned hw = std::thread::hardware_concurrency();
  size_t nt = std::min<size_t>(hw ? hw : 1, 16);
  if (n < 16384 || nt < 2) { f(0, n); return; }
  std::vector<std::thread> th;
  for (size_t t = 0; t < nt; t++) th.emplace_back([=]() { f(n * t / nt, n * (t + 1) / nt); });
  for (auto &x : th) x.join();
}

// stream-ordered device buffer
struct DevBuf {
  void *p = nullptr; cudaStream_t s = nullptr;
  int alloc(size_t bytes, cudaStream_t stream) {
    s = stream;
    P7G_CUDA(cudaMallocAsync(&p, std::max<size_t>(bytes, 256), stream));
    return P7GPU_OK;
  }
  ~DevBuf() { if (p) cudaFreeAsync(p, s); }
  template <class T> T *as() const { return (T *) p; }
};

__global__ void gather_pass_kernel(const int32_t *pass, unsigned n, const float *usc, const int32_t *st, const int32_t *len,
                                   int64_t *idx64, float *o_usc, int32_t *o_st, int32_t *o_len)
{
  unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int32_t s = pass[j];
  idx64[j] = s; o_usc[j] = usc[s]; o_st[j] = st[s]; o_len[j] = len[s];
}

struct Cand {              // a target still alive after the bias stage
  int64_t idx; int32_t L, msv_st, vit_st; float usc, nullsc, filtersc, vfsc, fwdsc, bcksc, ftot, btot; double P;
  int32_t pos;             // position in the MSV survivor list (opts->stage_records), else -1
};

// Host-side working arrays of the batched pipeline, kept with the context: a database-sized search allocates tens of MB of
// them per call, and fresh allocations of that size are mmap'ed and page-faulted every time (16 ms per call at 3e5 MSV
// survivors before they were made persistent).
struct HostScratch {
  std::vector<int64_t> idx, vidx; std::vector<float> usc, filtersc, vsc; std::vector<int32_t> st, len, vst;
  std::vector<double> P; std::vector<size_t> where; std::vector<Cand> cand;
  std::vector<p7gpu_msv_survivor> msv_cache;
  RawFloats xmx_cache[2];
};
static HostScratch &host_scratch(p7gpu_ctx *ctx)
{
  if (!ctx->host_scratch) ctx->host_scratch = std::shared_ptr<void>(new HostScratch(), [](void *p) { delete (HostScratch *) p; });
  return *(HostScratch *) ctx->host_scratch.get();
}

// sorted survivors of the MSV stage on the host (arrays live in the context's HostScratch)
struct MsvPass { std::vector<int64_t> &idx; std::vector<float> &usc; std::vector<int32_t> &st, &len; DevBuf d_idx64, g_usc, g_st, g_len;
                 explicit MsvPass(HostScratch &h) : idx(h.idx), usc(h.usc), st(h.st), len(h.len) {} };

static int fetch_msv_pass(p7gpu_ctx *ctx, const p7gpu_seqdb *db, const float *d_usc, const int32_t *d_st,
                          int32_t *d_pass, unsigned npass, MsvPass *out, bool to_host)
{
  cudaStream_t st = ctx->stream;
  if (to_host) { out->idx.resize(npass); out->usc.resize(npass); out->st.resize(npass); out->len.resize(npass); }
  P7G_CHECK(out->d_idx64.alloc(sizeof(int64_t) * npass, st));
  if (npass == 0) return P7GPU_OK;
  DevBuf sorted, tmp; DevBuf &g_usc = out->g_usc, &g_st = out->g_st, &g_len = out->g_len;
  P7G_CHECK(sorted.alloc(sizeof(int32_t) * npass, st));
  size_t tmp_bytes = 0;
  P7G_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_pass, sorted.as<int32_t>(), (int) npass, 0, 32, st));
  P7G_CHECK(tmp.alloc(tmp_bytes, st));
  P7G_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, d_pass, sorted.as<int32_t>(), (int) npass, 0, 32, st));
  P7G_CHECK(g_usc.alloc(4 * (size_t) npass, st)); P7G_CHECK(g_st.alloc(4 * (size_t) npass, st)); P7G_CHECK(g_len.alloc(4 * (size_t) npass, st));
  gather_pass_kernel<<<(npass + 255) / 256, 256, 0, st>>>(sorted.as<int32_t>(), npass, d_usc, d_st, db->d_len,
                                                           out->d_idx64.as<int64_t>(), g_usc.as<float>(), g_st.as<int32_t>(), g_len.as<int32_t>());
  P7G_CUDA(cudaGetLastError());
  ctx->launches += 3;
  if (!to_host) return P7GPU_OK;
  P7G_CUDA(cudaMemcpyAsync(out->idx.data(), out->d_idx64.p, 8 * (size_t) npass, cudaMemcpyDeviceToHost, st));
  P7G_CUDA(cudaMemcpyAsync(out->usc.data(), g_usc.p, 4 * (size_t) npass, cudaMemcpyDeviceToHost, st));
  P7G_CUDA(cudaMemcpyAsync(out->st.data(), g_st.p, 4 * (size_t) npass, cudaMemcpyDeviceToHost, st));
  P7G_CUDA(cudaMemcpyAsync(out->len.data(), g_len.p, 4 * (size_t) npass, cudaMemcpyDeviceToHost, st));
  P7G_CUDA(cudaStreamSynchronize(st));
  return P7GPU_OK;
}

// Large device -> pageable host copy: DMA into two page-locked bounce buffers, host threads move each piece to its
// destination while the next piece is in flight (a plain cudaMemcpy into pageable memory ran at 2.5 GB/s).
static int d2h_large(p7gpu_ctx *ctx, void *dst, const void *d_src, size_t bytes)
{
  cudaStream_t st = ctx->stream;
  const size_t piece = (size_t) 32 << 20;
  if (bytes <= piece) {
    if (bytes) { P7G_CUDA(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, st)); P7G_CUDA(cudaStreamSynchronize(st)); }
    return P7GPU_OK;
  }
  P7G_CHECK(ensure_pinned(ctx, 2 * piece));
  char *buf[2] = {(char *) ctx->h_pinned, (char *) ctx->h_pinned + piece};
  cudaEvent_t ev[2];
  P7G_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming)); P7G_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
  const size_t np = (bytes + piece - 1) / piece;
  auto len_of = [&](size_t k) { return std::min(piece, bytes - k * piece); };
  int rc = P7GPU_OK;
  cudaError_t e = cudaMemcpyAsync(buf[0], d_src, len_of(0), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaEventRecord(ev[0], st);
  for (size_t k = 0; k < np && e == cudaSuccess; k++) {
    if (k + 1 < np) {
      e = cudaMemcpyAsync(buf[(k + 1) & 1], (const char *) d_src + (k + 1) * piece, len_of(k + 1), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaEventRecord(ev[(k + 1) & 1], st);
      if (e != cudaSuccess) break;
    }
    e = cudaEventSynchronize(ev[k & 1]);
    if (e != cudaSuccess) break;
    char *d = (char *) dst + k * piece; const char *sbuf = buf[k & 1]; const size_t n = len_of(k);
    parallel_for(n >> 12, [&](size_t lo, size_t hi) { memcpy(d + (lo << 12), sbuf + (lo << 12), (hi - lo) << 12); });
    if (n & 4095) memcpy(d + (n & ~(size_t) 4095), sbuf + (n & ~(size_t) 4095), n & 4095);
    // the buffer just drained is refilled two pieces later: that copy is queued only after this memcpy returned
  }
  if (e != cudaSuccess) { set_error("d2h_large: %s", cudaGetErrorString(e)); rc = P7GPU_ESYS; cudaStreamSynchronize(st); }
  cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
  return rc;
}

// upload a host idx list, run `fn`, fetch float scores (+ statuses)
static int upload_idx(p7gpu_ctx *ctx, const std::vector<int64_t> &idx, DevBuf *d_idx)
{
  P7G_CHECK(d_idx->alloc(8 * idx.size(), ctx->stream));
  if (!idx.empty()) P7G_CUDA(cudaMemcpyAsync(d_idx->p, idx.data(), 8 * idx.size(), cudaMemcpyHostToDevice, ctx->stream));
  return P7GPU_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Stage decisions on the device.  Every P-value test of p7_Pipeline has the form "P(seq_score) > F" with
// seq_score = (float)((sc - filtersc) / ln 2) and P a non-increasing function of seq_score (esl_gumbel_surv, esl_exp_surv).
// The host finds, with its own libm and the reference's expressions, the smallest float T whose P passes, and checks that
// the 64 floats on either side of T are ordered the way monotonicity says; the device then only compares floats: its
// seq_score is formed with the same IEEE operations (float subtract, double divide, round to float) as the host's, so the
// decisions are the host's, bit for bit, with no transcendental evaluated on the device.
// ------------------------------------------------------------------------------------------------------------------
static inline float float_step(float x, int n)                     // n-th representable float after (before) x
{
  int32_t i; memcpy(&i, &x, 4);
  i = (i < 0) ? (int32_t) 0x80000000 - i : i;                      // to a monotone integer line
  i += n;
  i = (i < 0) ? (int32_t) 0x80000000 - i : i;
  memcpy(&x, &i, 4);
  return x;
}
template <class F> static bool pvalue_threshold(F passes, float guess, float *T)
{
  float lo = guess - 4.0f, hi = guess + 4.0f;
  if (passes(lo) || !passes(hi)) return false;
  while (float_step(lo, 1) < hi) {                                  // invariant: lo fails, hi passes
    const float mid = lo + 0.5f * (hi - lo);
    if (mid <= lo || mid >= hi) break;
    if (passes(mid)) hi = mid; else lo = mid;
  }
  while (float_step(lo, 1) < hi) { const float nx = float_step(lo, 1); if (passes(nx)) { hi = nx; break; } lo = nx; }
  for (int k = 1; k <= 64; k++) if (passes(float_step(hi, -k)) || !passes(float_step(hi, k))) return false;
  *T = hi;
  return true;
}

struct CandRec { int64_t idx; int32_t L, msv_st, vit_st; float usc, filtersc, vfsc; };

__device__ __forceinline__ float seq_score_of(float sc, float filtersc)
{
  return (float) ((double) __fsub_rn(sc, filtersc) / 0.69314718055994529);
}

__global__ void bias_decide_kernel(unsigned n, const float *usc, const float *fsc, const int32_t *len, float T1, float T2,
                                   uint8_t *keep, uint8_t *need, unsigned long long *sums)
{
  const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long l = 0;
  if (j < n) {
    const float x = seq_score_of(usc[j], fsc[j]);
    const bool k = (x >= T1);                                        // !(P > F1), p7_pipeline.c:700-703
    keep[j] = k; need[j] = k && !(x >= T2);                          // P > F2: the Viterbi filter runs, :716
    l = (unsigned long long) len[j];
  }
  for (int d = 16; d >= 1; d >>= 1) l += __shfl_xor_sync(0xffffffffu, l, d);
  if ((threadIdx.x & 31) == 0 && l) atomicAdd(sums, l);
}

__global__ void gather_vit_kernel(unsigned n, const int32_t *pos, const int64_t *idx64, const int32_t *len, int64_t *out_idx, unsigned long long *sums)
{
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long l = 0;
  if (k < n) { const int32_t j = pos[k]; out_idx[k] = idx64[j]; l = (unsigned long long) len[j]; }
  for (int d = 16; d >= 1; d >>= 1) l += __shfl_xor_sync(0xffffffffu, l, d);
  if ((threadIdx.x & 31) == 0 && l) atomicAdd(sums + 1, l);
}

__global__ void vit_decide_kernel(unsigned n, const int32_t *pos, const float *vsc, const int32_t *vst, const float *fsc, float Tv,
                                  uint8_t *keep, float *vf_full, int32_t *vs_full)
{
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int32_t j = pos[k];
  vf_full[j] = vsc[k]; vs_full[j] = vst[k];
  if (!(seq_score_of(vsc[k], fsc[j]) >= Tv)) keep[j] = 0;            // P > F2 after the Viterbi filter, :720-722
}

__global__ void gather_cand_kernel(unsigned n, const int32_t *pos, const int64_t *idx64, const int32_t *len, const int32_t *mst, const float *usc,
                                   const float *fsc, const float *vf_full, const int32_t *vs_full, const uint8_t *need, CandRec *out)
{
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int32_t j = pos[k];
  CandRec r; r.idx = idx64[j]; r.L = len[j]; r.msv_st = mst[j]; r.usc = usc[j]; r.filtersc = fsc[j];
  r.vit_st = need[j] ? vs_full[j] : -1; r.vfsc = need[j] ? vf_full[j] : __int_as_float(0x7fc00000);   // the host path's NAN
  out[k] = r;
}

struct IotaIt {                                                      // minimal counting iterator for cub::DeviceSelect
  using iterator_category = std::random_access_iterator_tag; using value_type = int32_t; using difference_type = ptrdiff_t;
  using pointer = const int32_t *; using reference = int32_t;
  int32_t v;
  __host__ __device__ int32_t operator*() const { return v; }
  __host__ __device__ int32_t operator[](ptrdiff_t i) const { return v + (int32_t) i; }
  __host__ __device__ IotaIt operator+(ptrdiff_t i) const { return IotaIt{v + (int32_t) i}; }
  __host__ __device__ IotaIt &operator+=(ptrdiff_t i) { v += (int32_t) i; return *this; }
  __host__ __device__ IotaIt &operator++() { ++v; return *this; }
};

static int select_positions(p7gpu_ctx *ctx, const uint8_t *d_flags, unsigned n, int32_t *d_out, unsigned *d_count)
{
  size_t bytes = 0;
  P7G_CUDA(cub::DeviceSelect::Flagged(nullptr, bytes, IotaIt{0}, d_flags, d_out, d_count, (int) n, ctx->stream));
  DevBuf tmp; P7G_CHECK(tmp.alloc(bytes, ctx->stream));
  P7G_CUDA(cub::DeviceSelect::Flagged(tmp.p, bytes, IotaIt{0}, d_flags, d_out, d_count, (int) n, ctx->stream));
  ctx->launches += 2;
  return P7GPU_OK;
}

// Stages B + C with the decisions on the device: fills `cand` (the targets alive after the Viterbi stage) and the counters.
static int bias_vit_on_device(p7gpu_ctx *ctx, const p7gpu_profile *prof, const DbView &dbv, MsvPass &mp, unsigned npass,
                              float T1, float T2, float Tv, std::vector<Cand> &cand, p7gpu_pipeline_result *res)
{
  cudaStream_t st = ctx->stream;
  cand.clear();
  if (npass == 0) return P7GPU_OK;
  DevBuf d_f, d_keep, d_need, d_sums, d_pos, d_vpos, d_cnt, d_vf, d_vs;
  P7G_CHECK(d_f.alloc(4 * (size_t) npass, st)); P7G_CHECK(d_keep.alloc(npass, st)); P7G_CHECK(d_need.alloc(npass, st));
  P7G_CHECK(d_sums.alloc(16, st)); P7G_CHECK(d_pos.alloc(4 * (size_t) npass, st)); P7G_CHECK(d_vpos.alloc(4 * (size_t) npass, st));
  P7G_CHECK(d_cnt.alloc(16, st)); P7G_CHECK(d_vf.alloc(4 * (size_t) npass, st)); P7G_CHECK(d_vs.alloc(4 * (size_t) npass, st));
  P7G_CUDA(cudaMemsetAsync(d_sums.p, 0, 16, st)); P7G_CUDA(cudaMemsetAsync(d_cnt.p, 0, 16, st));
  DevBuf d_unsure, d_ucount;
  P7G_CHECK(d_unsure.alloc(4 * (size_t) npass, st)); P7G_CHECK(d_ucount.alloc(16, st));
  P7G_CHECK(launch_bias(ctx, 0, prof, dbv, mp.d_idx64.as<int64_t>(), npass, d_f.as<float>(), d_unsure.as<int32_t>(), d_ucount.as<unsigned int>()));
  const unsigned nb = (npass + 255) / 256;
  bool refixed = false;
redo_decisions:
  P7G_CUDA(cudaMemsetAsync(d_sums.p, 0, 16, st));
  bias_decide_kernel<<<nb, 256, 0, st>>>(npass, mp.g_usc.as<float>(), d_f.as<float>(), mp.g_len.as<int32_t>(), T1, T2,
                                         d_keep.as<uint8_t>(), d_need.as<uint8_t>(), d_sums.as<unsigned long long>());
  P7G_CUDA(cudaGetLastError());
  unsigned *cnt = d_cnt.as<unsigned>();
  P7G_CHECK(select_positions(ctx, d_keep.as<uint8_t>(), npass, d_pos.as<int32_t>(), cnt + 0));
  P7G_CHECK(select_positions(ctx, d_need.as<uint8_t>(), npass, d_vpos.as<int32_t>(), cnt + 1));
  unsigned h_cnt[4] = {0, 0, 0, 0}, h_unsure = 0;
  P7G_CUDA(cudaMemcpyAsync(h_cnt, cnt, 8, cudaMemcpyDeviceToHost, st));
  P7G_CUDA(cudaMemcpyAsync(&h_unsure, d_ucount.p, 4, cudaMemcpyDeviceToHost, st));
  P7G_CUDA(cudaStreamSynchronize(st));
  if (h_unsure && !refixed) {                                        // ~3 targets per 10^7: filtersc from the host libm, decisions again
    P7G_CHECK(fix_unsure_filtersc(ctx, prof, dbv, mp.d_idx64.as<int64_t>(), d_f.as<float>(), d_unsure.as<int32_t>(), d_ucount.as<unsigned int>(), nullptr, nullptr));
    refixed = true;
    goto redo_decisions;
  }
  const unsigned nc = h_cnt[0], nv = h_cnt[1];
  res->counters.n_past_bias = nc;
  ctx->launches += 1;
  if (nv) {                                                          // Viterbi filter on the candidates that still need it
    DevBuf d_vidx, d_vsc, d_vst;
    P7G_CHECK(d_vidx.alloc(8 * (size_t) nv, st)); P7G_CHECK(d_vsc.alloc(4 * (size_t) nv, st)); P7G_CHECK(d_vst.alloc(4 * (size_t) nv, st));
    gather_vit_kernel<<<(nv + 255) / 256, 256, 0, st>>>(nv, d_vpos.as<int32_t>(), mp.d_idx64.as<int64_t>(), mp.g_len.as<int32_t>(),
                                                        d_vidx.as<int64_t>(), d_sums.as<unsigned long long>());
    P7G_CUDA(cudaGetLastError());
    P7G_CHECK(launch_vit(ctx, 0, prof, dbv, d_vidx.as<int64_t>(), (int64_t) nv, d_vsc.as<float>(), d_vst.as<int32_t>()));
    vit_decide_kernel<<<(nv + 255) / 256, 256, 0, st>>>(nv, d_vpos.as<int32_t>(), d_vsc.as<float>(), d_vst.as<int32_t>(), d_f.as<float>(), Tv,
                                                        d_keep.as<uint8_t>(), d_vf.as<float>(), d_vs.as<int32_t>());
    P7G_CUDA(cudaGetLastError());
    P7G_CHECK(select_positions(ctx, d_keep.as<uint8_t>(), npass, d_pos.as<int32_t>(), cnt + 2));
    ctx->launches += 2;
  } else {
    P7G_CUDA(cudaMemcpyAsync(cnt + 2, cnt + 0, 4, cudaMemcpyDeviceToDevice, st));
  }
  unsigned long long h_sums[2] = {0, 0};
  P7G_CUDA(cudaMemcpyAsync(h_cnt + 2, cnt + 2, 4, cudaMemcpyDeviceToHost, st));
  P7G_CUDA(cudaMemcpyAsync(h_sums, d_sums.p, 16, cudaMemcpyDeviceToHost, st));
  P7G_CUDA(cudaStreamSynchronize(st));
  const unsigned nf = h_cnt[2];
  res->stage_nres[1] += h_sums[0]; res->stage_nres[2] += h_sums[1];
  if (nf) {
    DevBuf d_rec; P7G_CHECK(d_rec.alloc(sizeof(CandRec) * (size_t) nf, st));
    gather_cand_kernel<<<(nf + 255) / 256, 256, 0, st>>>(nf, d_pos.as<int32_t>(), mp.d_idx64.as<int64_t>(), mp.g_len.as<int32_t>(), mp.g_st.as<int32_t>(),
                                                         mp.g_usc.as<float>(), d_f.as<float>(), d_vf.as<float>(), d_vs.as<int32_t>(), d_need.as<uint8_t>(),
                                                         d_rec.as<CandRec>());
    P7G_CUDA(cudaGetLastError());
    std::vector<CandRec> rec(nf);
    P7G_CUDA(cudaMemcpyAsync(rec.data(), d_rec.p, sizeof(CandRec) * (size_t) nf, cudaMemcpyDeviceToHost, st));
    P7G_CUDA(cudaStreamSynchronize(st));
    ctx->launches += 1;
    cand.resize(nf);
    const LenConst *hl = prof->h_len.data();
    for (unsigned k = 0; k < nf; k++) {
      Cand &c = cand[k]; const CandRec &r = rec[k];
      c.idx = r.idx; c.L = r.L; c.msv_st = r.msv_st; c.vit_st = r.vit_st; c.usc = r.usc; c.nullsc = hl[r.L].nullsc; c.filtersc = r.filtersc;
      c.vfsc = r.vfsc; c.fwdsc = NAN; c.bcksc = NAN; c.ftot = 0.f; c.btot = 0.f; c.P = 0.0; c.pos = -1;
    }
  }
  return P7GPU_OK;
}

// Stages B..E on a resident database, given the MSV stage's outputs.
static int pipeline_after_msv(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db, const p7gpu_pipeline_opts *o,
                              const float *d_usc, const int32_t *d_st, int32_t *d_pass, unsigned npass, p7gpu_pipeline_result *res)
{
  cudaStream_t st = ctx->stream;
  const DbView dbv = db->view();
  const double F1 = o->do_max ? 1.0 : o->F1, F2 = o->do_max ? 1.0 : o->F2, F3 = o->do_max ? 1.0 : o->F3;
  const bool do_bias = o->do_biasfilter && !o->do_max;
  const float *ev = prof->evparam;
  res->counters.nseqs = (uint64_t) db->nseq; res->counters.nres = (uint64_t) db->nres;
  res->counters.n_past_msv = npass;
  res->stage_nres[0] = (uint64_t) db->nres;

  Trace tr;
  HostScratch &hs = host_scratch(ctx);
  MsvPass mp(hs);
  // default thresholds: the bias and Viterbi stages decide on the device (float thresholds found and verified on the host)
  float T1 = 0.f, T2 = 0.f, Tv = 0.f;
  bool dev_decide = do_bias && !o->stop_after_msv && !o->stage_records && F1 > 0 && F1 < 1 && F2 > 0 && F2 < 1 && !getenv("P7GPU_HOST_DECISIONS");
  if (dev_decide) {
    auto gum = [&](double F, float mu, float lam) { return (float) (mu - log(-log1p(-F)) / lam); };
    dev_decide = pvalue_threshold([&](float x) { return !(host_gumbel_surv((double) x, ev[0], ev[1]) > F1); }, gum(F1, ev[0], ev[1]), &T1)
              && pvalue_threshold([&](float x) { return !(host_gumbel_surv((double) x, ev[0], ev[1]) > F2); }, gum(F2, ev[0], ev[1]), &T2)
              && pvalue_threshold([&](float x) { return !(host_gumbel_surv((double) x, ev[2], ev[3]) > F2); }, gum(F2, ev[2], ev[3]), &Tv);
  }
  P7G_CHECK(fetch_msv_pass(ctx, db, d_usc, d_st, d_pass, npass, &mp, !dev_decide));
  tr.mark("after_msv: fetch msv pass");
  if (o->stop_after_msv) {
    res->ctx = ctx;
    res->msv_pass.swap(hs.msv_cache);             // a buffer a previous result gave back: already mapped, no page faults
    res->msv_pass.resize(npass);
    parallel_for(npass, [&](size_t lo, size_t hi) {
      for (size_t j = lo; j < hi; j++) res->msv_pass[j] = {mp.idx[j], mp.len[j], mp.st[j], mp.usc[j], 1};
    });
    return P7GPU_OK;
  }

  std::vector<Cand> &cand = hs.cand; cand.clear();
  std::vector<uint8_t> stage;                                      // opts->stage_records: how far each MSV survivor got (1 MSV .. 4 Forward)
  if (o->stage_records) stage.assign(npass, 1);
  if (dev_decide) {
    P7G_CHECK(bias_vit_on_device(ctx, prof, dbv, mp, npass, T1, T2, Tv, cand, res));
    tr.mark("after_msv: bias + viterbi (device decisions)");
  } else {
  // ---- B. bias filter (p7_pipeline.c:697-705) ----
  std::vector<float> &filtersc = hs.filtersc; filtersc.resize(npass);
  if (do_bias && npass) {
    DevBuf d_f; P7G_CHECK(d_f.alloc(4 * (size_t) npass, st));
    for (unsigned j = 0; j < npass; j++) res->stage_nres[1] += (uint64_t) mp.len[j];
    DevBuf d_unsure, d_ucount;
    P7G_CHECK(d_unsure.alloc(4 * (size_t) npass, st)); P7G_CHECK(d_ucount.alloc(16, st));
    P7G_CHECK(launch_bias(ctx, 0, prof, dbv, mp.d_idx64.as<int64_t>(), npass, d_f.as<float>(), d_unsure.as<int32_t>(), d_ucount.as<unsigned int>()));
    P7G_CHECK(fix_unsure_filtersc(ctx, prof, dbv, mp.d_idx64.as<int64_t>(), d_f.as<float>(), d_unsure.as<int32_t>(), d_ucount.as<unsigned int>(), nullptr, nullptr));
    P7G_CUDA(cudaMemcpyAsync(filtersc.data(), d_f.p, 4 * (size_t) npass, cudaMemcpyDeviceToHost, st));
    P7G_CUDA(cudaStreamSynchronize(st));
  }
  tr.mark("after_msv: bias kernel");
  std::vector<double> &Pv = hs.P; Pv.resize(npass);
  const LenConst *hl = prof->h_len.data();
  parallel_for(npass, [&](size_t lo, size_t hi) {
    for (size_t j = lo; j < hi; j++)
      Pv[j] = host_filter_pvalue_gumbel(mp.usc[j], do_bias ? filtersc[j] : hl[mp.len[j]].nullsc, ev[0], ev[1]);
  });
  auto make_cand = [&](unsigned j) {
    Cand c;
    c.idx = mp.idx[j]; c.L = mp.len[j]; c.msv_st = mp.st[j]; c.vit_st = -1; c.usc = mp.usc[j]; c.nullsc = hl[c.L].nullsc;
    c.filtersc = do_bias ? filtersc[j] : c.nullsc; c.vfsc = NAN; c.fwdsc = NAN; c.bcksc = NAN; c.ftot = 0.f; c.btot = 0.f; c.P = Pv[j];
    c.pos = (int32_t) j;
    return c;
  };
  if (!do_bias) {                                                    // --max / --nobias: every MSV survivor goes on (10^7 records under --max)
    cand.resize(npass);
    parallel_for(npass, [&](size_t lo, size_t hi) { for (size_t j = lo; j < hi; j++) cand[j] = make_cand((unsigned) j); });
  } else {
    for (unsigned j = 0; j < npass; j++) if (!(Pv[j] > F1)) cand.push_back(make_cand(j));
  }
  res->counters.n_past_bias = cand.size();
  tr.mark("after_msv: bias P-values");

  // ---- C. Viterbi filter for the candidates whose P still exceeds F2 (:716-723) ----
  {
    std::vector<int64_t> &vidx = hs.vidx; std::vector<size_t> &where = hs.where; vidx.clear(); where.clear();
    for (size_t j = 0; j < cand.size(); j++) if (cand[j].P > F2) { vidx.push_back(cand[j].idx); where.push_back(j); res->stage_nres[2] += (uint64_t) cand[j].L; }
    if (!vidx.empty()) {
      DevBuf d_idx, d_sc, d_stt;
      P7G_CHECK(upload_idx(ctx, vidx, &d_idx));
      P7G_CHECK(d_sc.alloc(4 * vidx.size(), st)); P7G_CHECK(d_stt.alloc(4 * vidx.size(), st));
      P7G_CHECK(launch_vit(ctx, 0, prof, dbv, d_idx.as<int64_t>(), (int64_t) vidx.size(), d_sc.as<float>(), d_stt.as<int32_t>()));
      std::vector<float> &vsc = hs.vsc; std::vector<int32_t> &vst = hs.vst; vsc.resize(vidx.size()); vst.resize(vidx.size());
      P7G_CUDA(cudaMemcpyAsync(vsc.data(), d_sc.p, 4 * vidx.size(), cudaMemcpyDeviceToHost, st));
      P7G_CUDA(cudaMemcpyAsync(vst.data(), d_stt.p, 4 * vidx.size(), cudaMemcpyDeviceToHost, st));
      P7G_CUDA(cudaStreamSynchronize(st));
      parallel_for(vidx.size(), [&](size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; k++) {
          Cand &c = cand[where[k]];
          c.vfsc = vsc[k]; c.vit_st = vst[k];
          c.P = host_filter_pvalue_gumbel(c.vfsc, c.filtersc, ev[2], ev[3]);
        }
      });
      size_t w = 0;
      for (size_t j = 0; j < cand.size(); j++) { if (cand[j].vit_st >= 0 && cand[j].P > F2) continue; cand[w++] = cand[j]; }
      cand.resize(w);
    }
  }
  }   // host decisions
  res->counters.n_past_vit = cand.size();
  if (o->stage_records) {
    for (unsigned j = 0; j < npass; j++) if (!(do_bias && hs.P[j] > F1)) stage[j] = 2;     // passed the bias filter (or it is off)
    for (const Cand &c : cand) if (c.pos >= 0) stage[(size_t) c.pos] = 3;
  }
  tr.mark("after_msv: viterbi stage");

  // ---- D + E. Forward parser, F3 test, Backward parser for the survivors (:727-735), in slices bounded by the
  //      size of the specials ((L+1)*6 floats per target and direction) ----
  // opts->fb_exact: whatever reaches the host is computed by the exact-order kernels (fbx.cu), bit-identical to the reference.
  // Under the default thresholds the fast Forward kernel first discards every candidate whose P-value cannot reach F3 even if
  // its score were 2e-4 relative higher (the fast kernels are tested to 1e-4, measured 4e-6); what is left — the survivors and
  // the few candidates next to the threshold — is recomputed exactly and decided on the exact score.
  const bool exact = o->fb_exact != 0;
  if (exact && F3 < 1.0 && !cand.empty()) {
    const size_t n = cand.size();
    std::vector<int64_t> idx(n);
    for (size_t k = 0; k < n; k++) { idx[k] = cand[k].idx; res->stage_nres[3] += (uint64_t) cand[k].L; }
    DevBuf d_idx, d_sc, d_stt;
    P7G_CHECK(upload_idx(ctx, idx, &d_idx));
    P7G_CHECK(d_sc.alloc(4 * n, st)); P7G_CHECK(d_stt.alloc(4 * n, st));
    P7G_CHECK(launch_fwd(ctx, 0, prof, dbv, d_idx.as<int64_t>(), (int64_t) n, d_sc.as<float>(), d_stt.as<int32_t>(), nullptr, nullptr, nullptr));
    std::vector<float> fsc(n);
    P7G_CUDA(cudaMemcpyAsync(fsc.data(), d_sc.p, 4 * n, cudaMemcpyDeviceToHost, st));
    P7G_CUDA(cudaStreamSynchronize(st));
    size_t w = 0;
    for (size_t k = 0; k < n; k++) {
      const float f = fsc[k];
      const float hi = std::isfinite(f) ? f + 2e-4f * std::max(1.0f, std::fabs(f)) : f;
      if (std::isfinite(f) && host_filter_pvalue_exp(hi, cand[k].filtersc, ev[4], ev[5]) > F3) continue;   // cannot pass
      cand[w++] = cand[k];
    }
    cand.resize(w);
    tr.mark("after_msv: forward prefilter (fast kernel)");
  }
  const size_t slice_floats = (size_t) 1 << 29;                 // 2 GiB of forward specials per slice
  size_t host_floats = 0;
  RawFloats exact_fx;                                           // exact mode, F3 < 1: the slice's forward specials on the host (scores)
  for (size_t lo = 0; lo < cand.size();) {
    size_t hi = lo, ext = 0;
    std::vector<int64_t> idx, xoff;
    idx.reserve(std::min<size_t>(cand.size() - lo, (size_t) 1 << 20)); xoff.reserve(idx.capacity());
    while (hi < cand.size() && (hi == lo || ext + (size_t) (cand[hi].L + 1) * 6 <= slice_floats)) {
      idx.push_back(cand[hi].idx); xoff.push_back((int64_t) ext); ext += (size_t) (cand[hi].L + 1) * 6;
      res->stage_nres[exact && F3 < 1.0 ? 5 : 3] += (uint64_t) cand[hi].L; hi++;
    }
    const size_t n = hi - lo;
    DevBuf d_idx, d_xoff, d_sc, d_stt, d_tot, d_fx;
    P7G_CHECK(upload_idx(ctx, idx, &d_idx)); P7G_CHECK(upload_idx(ctx, xoff, &d_xoff));
    P7G_CHECK(d_sc.alloc(4 * n, st)); P7G_CHECK(d_stt.alloc(4 * n, st)); P7G_CHECK(d_tot.alloc(4 * n, st));
    P7G_CHECK(d_fx.alloc(4 * ext, st));
    if (exact) P7G_CHECK(launch_fbx_fwd(ctx, 0, prof, dbv, d_idx.as<int64_t>(), (int64_t) n, d_sc.as<float>(), d_stt.as<int32_t>(),
                                        d_fx.as<float>(), d_xoff.as<int64_t>(), d_tot.as<float>()));
    else       P7G_CHECK(launch_fwd(ctx, 0, prof, dbv, d_idx.as<int64_t>(), (int64_t) n, d_sc.as<float>(), d_stt.as<int32_t>(),
                                    d_fx.as<float>(), d_xoff.as<int64_t>(), d_tot.as<float>()));
    std::vector<float> fsc(n), ftot(n);
    P7G_CUDA(cudaMemcpyAsync(fsc.data(), d_sc.p, 4 * n, cudaMemcpyDeviceToHost, st));
    P7G_CUDA(cudaMemcpyAsync(ftot.data(), d_tot.p, 4 * n, cudaMemcpyDeviceToHost, st));
    const bool allpass = F3 >= 1.0;                                  // --max: no decision between the two parsers, so no host round trip either
    if (!allpass) { P7G_CUDA(cudaStreamSynchronize(st)); tr.mark("  slice: forward kernel"); }
    if (exact && F3 < 1.0) {                                     // the F3 test needs the exact score: host libm on the specials
      DevBuf d_cs;                                               // SCALE column + xC(L): a sixth of the specials
      P7G_CHECK(d_cs.alloc(4 * (ext / 6), st));
      P7G_CHECK(launch_gather_scale(ctx, 0, dbv, d_idx.as<int64_t>(), (int64_t) n, d_fx.as<float>(), d_xoff.as<int64_t>(), d_cs.as<float>()));
      if (!exact_fx.resize(ext / 6)) { set_error("pipeline: out of host memory for %zu floats of scale factors", ext / 6); return P7GPU_EMEM; }
      P7G_CHECK(d2h_large(ctx, exact_fx.data(), d_cs.p, 4 * (ext / 6)));
      P7G_CUDA(cudaStreamSynchronize(st));
      const LenConst *hl = prof->h_len.data();
      parallel_for(n, [&](size_t a, size_t b) {
        for (size_t k = a; k < b; k++) { int stt; fbx_host_fwd_score_compact(exact_fx.data() + xoff[k] / 6, cand[lo + k].L, hl[cand[lo + k].L].pmove, &ftot[k], &fsc[k], &stt); }
      });
    }
    tr.mark("  slice: forward done");
    std::vector<int64_t> bidx, boff; std::vector<size_t> bwhere;
    if (allpass) {                                                   // --max: P <= 1 always passes (p7_pipeline.c:730); no P-value, no compaction
      bidx = idx; boff = xoff; bwhere.resize(n);
      parallel_for(n, [&](size_t a, size_t b) { for (size_t k = a; k < b; k++) bwhere[k] = lo + k; });   // scores are read after the Backward sync
      res->stage_nres[4] += (uint64_t) ext / 6 - n;                  // sum of L over the slice: ext = sum (L+1)*6
      if (o->stage_records) for (size_t k = 0; k < n; k++) if (cand[lo + k].pos >= 0) stage[(size_t) cand[lo + k].pos] = 4;
    } else
    for (size_t k = 0; k < n; k++) {
      Cand &c = cand[lo + k];
      c.fwdsc = fsc[k]; c.ftot = ftot[k];
      c.P = host_filter_pvalue_exp(c.fwdsc, c.filtersc, ev[4], ev[5]);
      if (c.P > F3) continue;
      bidx.push_back(c.idx); boff.push_back(xoff[k]); bwhere.push_back(lo + k); res->stage_nres[4] += (uint64_t) c.L;
      if (o->stage_records && c.pos >= 0) stage[(size_t) c.pos] = 4;
    }
    res->counters.n_past_fwd += bidx.size();
    if (!bidx.empty()) {
      const size_t nb = bidx.size();
      DevBuf d_bidx, d_boff, d_bsc, d_bst, d_btot, d_bx, d_own, d_mocc, d_flag;
      if (allpass) { d_bidx.p = d_idx.p; d_bidx.s = st; d_idx.p = nullptr; d_boff.p = d_xoff.p; d_boff.s = st; d_xoff.p = nullptr; }   // same lists: hand the buffers over
      else { P7G_CHECK(upload_idx(ctx, bidx, &d_bidx)); P7G_CHECK(upload_idx(ctx, boff, &d_boff)); }
      P7G_CHECK(d_bsc.alloc(4 * nb, st)); P7G_CHECK(d_bst.alloc(4 * nb, st)); P7G_CHECK(d_btot.alloc(4 * nb, st));
      P7G_CHECK(d_own.alloc(4 * nb, st));
      const bool keep_bx = o->want_specials || o->region_trigger;
      P7G_CHECK(d_bx.alloc(keep_bx ? 4 * ext : 256, st));
      if (exact) P7G_CHECK(launch_fbx_bwd(ctx, 0, prof, dbv, d_bidx.as<int64_t>(), (int64_t) nb, d_fx.as<float>(), d_boff.as<int64_t>(),
                                          d_bsc.as<float>(), d_bst.as<int32_t>(), keep_bx ? d_bx.as<float>() : nullptr, d_btot.as<float>(),
                                          d_own.as<int32_t>()));
      else       P7G_CHECK(launch_bwd(ctx, 0, prof, dbv, d_bidx.as<int64_t>(), (int64_t) nb, d_fx.as<float>(), d_boff.as<int64_t>(),
                                      d_bsc.as<float>(), d_bst.as<int32_t>(), keep_bx ? d_bx.as<float>() : nullptr, d_btot.as<float>(),
                                      d_own.as<int32_t>()));
      std::vector<float> bsc(nb), btot(nb), mocc(nb, NAN); std::vector<int32_t> own(nb), flag(nb, -1);
      P7G_CUDA(cudaMemcpyAsync(bsc.data(), d_bsc.p, 4 * nb, cudaMemcpyDeviceToHost, st));
      P7G_CUDA(cudaMemcpyAsync(btot.data(), d_btot.p, 4 * nb, cudaMemcpyDeviceToHost, st));
      P7G_CUDA(cudaMemcpyAsync(own.data(), d_own.p, 4 * nb, cudaMemcpyDeviceToHost, st));
      if (o->region_trigger) {                                       // p7_domaindef.c:419 on the device (decode.cu)
        P7G_CHECK(d_mocc.alloc(4 * nb, st)); P7G_CHECK(d_flag.alloc(4 * nb, st));
        P7G_CHECK(launch_region_trigger(ctx, 0, prof, dbv, d_bidx.as<int64_t>(), (int64_t) nb, d_fx.as<float>(), d_bx.as<float>(),
                                        d_boff.as<int64_t>(), d_own.as<int32_t>(), d_mocc.as<float>(), d_flag.as<int32_t>(), nullptr,
                                        exact ? 0.0f : 1.0e-4f));
        P7G_CUDA(cudaMemcpyAsync(mocc.data(), d_mocc.p, 4 * nb, cudaMemcpyDeviceToHost, st));
        P7G_CUDA(cudaMemcpyAsync(flag.data(), d_flag.p, 4 * nb, cudaMemcpyDeviceToHost, st));
      }
      P7G_CUDA(cudaStreamSynchronize(st));
      if (allpass) parallel_for(n, [&](size_t a, size_t b) { for (size_t k = a; k < b; k++) { Cand &c = cand[lo + k]; c.fwdsc = fsc[k]; c.ftot = ftot[k]; c.P = 0.0; } });
      tr.mark("  slice: backward + trigger done");
      // which survivors' specials go to the host: all of them, or those whose region trigger fired; with opts->decode the flagged
      // ones get their p7_DomainDecoding arrays instead (16 bytes per row for 48), except those the host must decide itself
      std::vector<int64_t> src, dst; std::vector<int32_t> nfl; std::vector<int64_t> host_off(nb, -1), dec_off(nb, -1);
      std::vector<int64_t> didx, dxoff, ddoff; std::vector<size_t> dwhere;
      const bool decode = o->decode && o->want_specials && o->region_trigger;
      size_t need = 0, dneed = 0;
      if (o->want_specials)
        for (size_t k = 0; k < nb; k++) {
          if (o->region_trigger && flag[k] == 0) continue;
          const int Lk = cand[bwhere[k]].L;
          if (decode && !own[k] && std::isfinite(mocc[k])) {          // mocc_max is NaN/inf only when the scale product is not finite
            didx.push_back(bidx[k]); dxoff.push_back(boff[k]); ddoff.push_back((int64_t) dneed); dwhere.push_back(k);
            dec_off[k] = (int64_t) (res->dec.n + dneed);
            dneed += (size_t) 4 * (Lk + 1) + 2;
            continue;
          }
          const size_t len = (size_t) (Lk + 1) * 6;
          src.push_back(boff[k]); dst.push_back((int64_t) need); nfl.push_back((int32_t) len);
          host_off[k] = (int64_t) (host_floats + need);
          need += len;
        }
      if (dneed) {
        DevBuf d_didx, d_dxoff, d_ddoff, d_dec;
        P7G_CHECK(upload_idx(ctx, didx, &d_didx)); P7G_CHECK(upload_idx(ctx, dxoff, &d_dxoff)); P7G_CHECK(upload_idx(ctx, ddoff, &d_ddoff));
        P7G_CHECK(d_dec.alloc(4 * dneed, st));
        P7G_CHECK(launch_domain_decoding(ctx, 0, prof, dbv, d_didx.as<int64_t>(), (int64_t) didx.size(), d_fx.as<float>(), d_bx.as<float>(),
                                         d_dxoff.as<int64_t>(), d_ddoff.as<int64_t>(), d_dec.as<float>()));
        const size_t base = res->dec.n;
        if (!res->ctx) { res->ctx = ctx; res->fwd_xmx.swap(hs.xmx_cache[0]); res->bck_xmx.swap(hs.xmx_cache[1]); res->fwd_xmx.n = res->bck_xmx.n = 0; }
        if (!res->dec.resize(base + dneed)) { set_error("pipeline: out of host memory for %zu floats of decoding arrays", base + dneed); return P7GPU_EMEM; }
        P7G_CHECK(d2h_large(ctx, res->dec.data() + base, d_dec.p, 4 * dneed));
        P7G_CUDA(cudaStreamSynchronize(st));
        const LenConst *hl = prof->h_len.data();
        parallel_for(dwhere.size(), [&](size_t a, size_t b) {          // Forward totscale and score from the SCALE column, host libm
          for (size_t t = a; t < b; t++) {
            const size_t k = dwhere[t]; Cand &c = cand[bwhere[k]];
            const float *blk = res->dec.data() + dec_off[k], *fs = blk + (size_t) 3 * (c.L + 1);
            float tot = 0.0f;
            for (int i = 1; i <= c.L; i++) if (fs[i] != 1.0f) tot = (float) ((double) tot + log((double) fs[i]));
            const float xC = fs[c.L + 1];
            c.ftot = tot;
            if (exact && !(std::isnan(xC) || (c.L > 0 && xC == 0.0f) || std::isinf(xC))) c.fwdsc = (float) ((double) tot + log((double) (xC * hl[c.L].pmove)));
          }
        });
      }
      if (need) {
        if ((host_floats + need) * 8 > ((size_t) 48 << 30)) {
          set_error("pipeline: Forward/Backward specials of the survivors exceed 48 GiB of host memory; run with region_trigger=1 "
                    "(specials only for targets that can have a domain) or want_specials=0");
          return P7GPU_EMEM;
        }
        if (!res->ctx) { res->ctx = ctx; res->fwd_xmx.swap(hs.xmx_cache[0]); res->bck_xmx.swap(hs.xmx_cache[1]); res->fwd_xmx.n = res->bck_xmx.n = 0; }
        if (!res->fwd_xmx.resize(host_floats + need) || !res->bck_xmx.resize(host_floats + need)) {
          set_error("pipeline: out of host memory for %zu floats of specials", 2 * (host_floats + need)); return P7GPU_EMEM;
        }
        if (src.size() == n) {                                       // every target of the slice: one bulk copy each
          P7G_CHECK(d2h_large(ctx, res->fwd_xmx.data() + host_floats, d_fx.p, 4 * need));
          P7G_CHECK(d2h_large(ctx, res->bck_xmx.data() + host_floats, d_bx.p, 4 * need));
        } else {                                                     // dense copy on the device first, then one bulk copy each
          DevBuf d_src, d_dst, d_nfl, d_ofx, d_obx;
          P7G_CHECK(upload_idx(ctx, src, &d_src)); P7G_CHECK(upload_idx(ctx, dst, &d_dst));
          P7G_CHECK(d_nfl.alloc(4 * nfl.size(), st));
          P7G_CUDA(cudaMemcpyAsync(d_nfl.p, nfl.data(), 4 * nfl.size(), cudaMemcpyHostToDevice, st));
          P7G_CHECK(d_ofx.alloc(4 * need, st)); P7G_CHECK(d_obx.alloc(4 * need, st));
          P7G_CHECK(launch_gather_xmx(ctx, 0, d_fx.as<float>(), d_bx.as<float>(), d_src.as<int64_t>(), d_dst.as<int64_t>(),
                                      d_nfl.as<int32_t>(), (int64_t) src.size(), d_ofx.as<float>(), d_obx.as<float>()));
          P7G_CHECK(d2h_large(ctx, res->fwd_xmx.data() + host_floats, d_ofx.p, 4 * need));
          P7G_CHECK(d2h_large(ctx, res->bck_xmx.data() + host_floats, d_obx.p, 4 * need));
          P7G_CUDA(cudaStreamSynchronize(st));                       // the staging vectors and buffers die with this scope
        }
        P7G_CUDA(cudaStreamSynchronize(st));
        if (exact) {                                                 // scores of what the host sees: its own libm on the specials
          const LenConst *hl = prof->h_len.data();
          parallel_for(nb, [&](size_t a, size_t b) {
            for (size_t k = a; k < b; k++) {
              if (host_off[k] < 0) continue;
              Cand &c = cand[bwhere[k]]; int stt;
              fbx_host_fwd_score(res->fwd_xmx.data() + host_off[k], c.L, hl[c.L].pmove, &c.ftot, &c.fwdsc, &stt);
              fbx_host_bwd_score(res->bck_xmx.data() + host_off[k], c.L, &btot[k], &bsc[k], &stt);
            }
          });
        }
      }
      tr.mark("  slice: specials on the host");
      {
        const size_t base = res->survivors.size();
        if (F3 >= 1.0 && res->survivors.capacity() < cand.size()) res->survivors.reserve(cand.size());   // --max: one allocation, not a doubling series
        res->survivors.resize(base + nb);
        p7gpu_survivor *out = res->survivors.data() + base;
        parallel_for(nb, [&](size_t a, size_t b) {
          for (size_t k = a; k < b; k++) {
            const Cand &c = cand[bwhere[k]];
            p7gpu_survivor s;
            s.seqidx = c.idx; s.L = c.L; s.msv_status = c.msv_st; s.vit_status = c.vit_st;
            s.usc = c.usc; s.vfsc = c.vfsc; s.fwdsc = c.fwdsc; s.bcksc = bsc[k]; s.nullsc = c.nullsc; s.filtersc = c.filtersc;
            s.fwd_totscale = c.ftot; s.bck_totscale = btot[k];
            s.bck_own_scales = own[k];
            s.xmx_offset = host_off[k];
            s.region_flag = flag[k]; s.mocc_max = mocc[k];
            s.dec_offset = dec_off[k];
            out[k] = s;
          }
        });
      }
      host_floats += need;
      res->n_region_flagged += (uint64_t) (src.size() + didx.size());
      tr.mark("  slice: survivor records");
    }
    lo = hi;
  }
  if (o->stage_records) {
    res->msv_pass.resize(npass);
    for (unsigned j = 0; j < npass; j++) res->msv_pass[j] = {mp.idx[j], mp.len[j], mp.st[j], mp.usc[j], (int32_t) stage[j]};
  }
  tr.mark("after_msv: forward/backward");
  return P7GPU_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Streamed MSV pass over a HOST database: the packed residues are copied in chunks on a copy stream while the SSV/MSV
// kernels of earlier chunks run on alternating lanes.  The database stays resident in HBM (returned in *ret_db).
// ------------------------------------------------------------------------------------------------------------------
// 5-bit transfer format of the residues (p7gpu_pack5): eight codes (0..28, 31 = the 255 sentinel) in five bytes, little-endian
// bit order.  One thread expands 40 packed bytes into 64 residue bytes (both sides 4-byte / 16-byte aligned).
__global__ void __launch_bounds__(128) unpack5_kernel(const uint32_t *__restrict__ in, uint4 *__restrict__ out, long long nblocks64)
{
  const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nblocks64) return;
  uint32_t w[10];
#pragma unroll
  for (int k = 0; k < 10; k++) w[k] = __ldg(in + t * 10 + k);
  uint32_t o[16];
#pragma unroll
  for (int g = 0; g < 8; g++) {                                        // group g: packed bytes 5g..5g+4 = bits 40g..40g+39
    const int bit = 40 * g, wi = bit >> 5, sh = bit & 31;
    const unsigned long long v = ((unsigned long long) w[wi] >> sh) | ((unsigned long long) w[wi + 1] << (32 - sh));   // sh is 0, 8, 16 or 24: 40 bits fit
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      uint32_t a = (uint32_t) (v >> (5 * r)) & 31u, b = (uint32_t) (v >> (5 * (r + 4))) & 31u;
      a = (a == 31u) ? 255u : a; b = (b == 31u) ? 255u : b;
      lo |= a << (8 * r); hi |= b << (8 * r);
    }
    o[2 * g] = lo; o[2 * g + 1] = hi;
  }
#pragma unroll
  for (int k = 0; k < 4; k++) out[t * 4 + k] = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
}

__global__ void iota_from_kernel(int32_t *v, int32_t first, int32_t n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = first + i;
}

static int stream_msv(p7gpu_ctx *ctx, const p7gpu_profile *prof, const uint8_t *residues, const uint8_t *packed5, int64_t nbytes,
                      const int64_t *offsets, const int32_t *lengths, int64_t nseq, double F1,
                      float *h_sc, int32_t *h_status, float *d_sc, int32_t *d_st, int32_t *d_pass, unsigned int *d_pass_count,
                      p7gpu_seqdb **ret_db)
{
  if (nseq > 0x7fffffffLL) { set_error("stream: more than 2^31-1 targets per device"); return P7GPU_EINVAL; }
  Trace tr;
  // ---- chunk plan: 16 MB first (0.3 ms of exposed copy), then x4/3 up to 256 MB.  A chunk can only be computed once it
  //      is copied completely, so while chunk c is in flight the GPU has chunk c-1 to work on: growth must stay below the
  //      ratio of copy rate (55 GB/s measured) to SSV consumption rate (38 GB/s of residues at M=400), or the kernels
  //      starve during the ramp (doubling lost 5 ms of a 96 ms pass) ----
  struct Chunk { int64_t s_lo, s_hi; };
  std::vector<Chunk> plan;
  int64_t max_chunk_seqs = 1;
  {
    const char *e0 = getenv("P7GPU_STREAM_CHUNK0_MB"), *e1 = getenv("P7GPU_STREAM_CHUNKMAX_MB"), *e2 = getenv("P7GPU_STREAM_GROWTH_PCT");
    // (the 5-bit transfer form arrives 1.6x faster per residue, 88 against 40 GB/s of consumption: chunks may double, and fewer,
    //  larger chunks mean fewer kernel boundaries — 0.12 ms each; measured 96.5 -> 95.0 ms per pass, scripts/e2e_chunk_sweep.sh)
    const int64_t chunk0 = (int64_t) (e0 ? atoi(e0) : 16) << 20, chunkmax = (int64_t) (e1 ? atoi(e1) : (packed5 ? 1024 : 256)) << 20;
    const int growth = e2 ? atoi(e2) : (packed5 ? 100 : 33);
    int64_t s = 0, bytes = chunk0;
    while (s < nseq) {
      const int64_t lim = offsets[s] + bytes;
      int64_t lo = s + 1, hi = nseq;
      while (lo < hi) { int64_t mid = (lo + hi) / 2; if (offsets[mid] <= lim) lo = mid + 1; else hi = mid; }
      plan.push_back({s, lo});
      max_chunk_seqs = std::max(max_chunk_seqs, lo - s);
      s = lo;
      bytes = std::min<int64_t>(bytes + bytes * growth / 100, chunkmax);
    }
  }
  cudaStream_t ms = ctx->stream, cs = ctx->copy_stream;
  p7gpu_seqdb *db = new p7gpu_seqdb();
  db->ctx = ctx; db->nseq = nseq; db->nbytes = nbytes; db->owns = true; db->pooled = true;
  const size_t padded = ((size_t) nbytes + 63) / 64 * 64 + 64;
  uint8_t *d_res = nullptr, *d_pack = nullptr; int64_t *d_off = nullptr; int32_t *d_len = nullptr;
  int64_t *d_nores[3] = {nullptr, nullptr, nullptr}; void *d_sorttmp[3] = {nullptr, nullptr, nullptr};
  int32_t *d_keys_out[3] = {nullptr, nullptr, nullptr}, *d_iota[3] = {nullptr, nullptr, nullptr};
  size_t sort_tmp_bytes = 0;
  std::vector<cudaEvent_t> events;
  cudaEvent_t ev_main = nullptr;
  auto cleanup = [&]() {
    for (auto e : events) cudaEventDestroy(e);
    if (ev_main) cudaEventDestroy(ev_main);
    if (d_pack) cudaFreeAsync(d_pack, ms);
    for (int l = 1; l <= 2; l++) {
      if (d_nores[l]) cudaFreeAsync(d_nores[l], ms);
      if (d_sorttmp[l]) cudaFreeAsync(d_sorttmp[l], ms);
      if (d_keys_out[l]) cudaFreeAsync(d_keys_out[l], ms);
      if (d_iota[l]) cudaFreeAsync(d_iota[l], ms);
    }
  };
  auto fail = [&](int code) { cudaDeviceSynchronize(); cleanup(); p7gpu_seqdb_destroy(db); return code; };
#define S_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); return fail(P7GPU_ESYS); } } while (0)
#define S_CHECK(call) do { int s__ = (call); if (s__ != P7GPU_OK) return fail(s__); } while (0)
  // stream-ordered allocations: after the first call they come back from the pool without touching the driver
  S_CUDA(cudaMallocAsync((void **) &d_res, padded, ms));                                       db->d_res = d_res;
  if (packed5) S_CUDA(cudaMallocAsync((void **) &d_pack, padded / 64 * 40, ms));
  S_CUDA(cudaMallocAsync((void **) &d_off, 8 * (size_t) std::max<int64_t>(nseq, 1), ms));      db->d_off = d_off;
  S_CUDA(cudaMallocAsync((void **) &d_len, 4 * (size_t) std::max<int64_t>(nseq, 1), ms));      db->d_len = d_len;
  S_CUDA(cudaMallocAsync((void **) &db->d_order, 4 * (size_t) std::max<int64_t>(nseq, 1), ms));
  S_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_tmp_bytes, (const int32_t *) nullptr, (int32_t *) nullptr,
                                                   (const int32_t *) nullptr, (int32_t *) nullptr, (int) max_chunk_seqs, 0, 32, ms));
  for (int l = 1; l <= 2; l++) {
    S_CUDA(cudaMallocAsync((void **) &d_nores[l], 16 * (size_t) max_chunk_seqs, ms));
    S_CUDA(cudaMallocAsync(&d_sorttmp[l], std::max<size_t>(sort_tmp_bytes, 256), ms));
    S_CUDA(cudaMallocAsync((void **) &d_keys_out[l], 4 * (size_t) max_chunk_seqs, ms));
    S_CUDA(cudaMallocAsync((void **) &d_iota[l], 4 * (size_t) max_chunk_seqs, ms));
  }
  S_CUDA(cudaMemsetAsync(d_res + (padded - 128), 0xff, 128, ms));
  S_CUDA(cudaMemsetAsync(d_pass_count, 0, sizeof(unsigned int), ms));
  // everything queued on the copy stream and the lanes comes after what is already queued on the main stream
  S_CUDA(cudaEventCreateWithFlags(&ev_main, cudaEventDisableTiming));
  S_CUDA(cudaEventRecord(ev_main, ms));
  S_CUDA(cudaStreamWaitEvent(cs, ev_main, 0));
  for (int l = 1; l <= 3; l++) S_CUDA(cudaStreamWaitEvent(ctx->lanes[l].stream, ev_main, 0));
  cudaStream_t us = ctx->lanes[3].stream;               // the 5-bit chunks are expanded here: see below
  tr.mark("stream: setup");

  // the per-length constant table is sized once, before anything is queued: growing it in the middle of the stream would
  // synchronise the device under the copy/compute overlap (the per-chunk call below is then a no-op)
  {
    std::vector<int32_t> part(64, 0);
    parallel_for((size_t) nseq, [&](size_t a, size_t b) {
      int32_t m = 0;
      for (size_t k = a; k < b; k++) m = std::max(m, lengths[k]);
      int32_t &slot = part[(a * 64) / (size_t) std::max<int64_t>(nseq, 1) % 64];
      slot = std::max(slot, m);
    });
    int32_t Lall = 0;
    for (int32_t m : part) Lall = std::max(Lall, m);
    S_CHECK(ensure_len_table(prof, Lall, F1));
  }
  int64_t nres = 0; int32_t Lmax = 0;
  for (size_t c = 0; c < plan.size(); c++) {
    const int64_t s_lo = plan[c].s_lo, s_hi = plan[c].s_hi, cn = s_hi - s_lo;
    // validate this chunk's targets (same checks as p7gpu_seqdb_upload) while earlier chunks run
    int32_t cLmax = 0; int64_t cres = 0; bool bad = false;
    int64_t prev_end = s_lo > 0 ? offsets[s_lo - 1] + lengths[s_lo - 1] : 0;
    for (int64_t s = s_lo; s < s_hi; s++) {
      const int64_t o = offsets[s]; const int32_t L = lengths[s];
      bad |= (L < 0) | (o <= prev_end) | (o + L >= nbytes);
      prev_end = o + (L > 0 ? L : 0);
      cLmax = std::max(cLmax, L); cres += L;
    }
    if (bad) {
      set_error("stream: targets %lld..%lld: an (offset, length) pair does not fit the packed array of %lld bytes or overlaps its predecessor",
                (long long) s_lo, (long long) s_hi - 1, (long long) nbytes);
      return fail(P7GPU_EINVAL);
    }
    nres += cres; Lmax = std::max(Lmax, cLmax);
    S_CHECK(ensure_len_table(prof, Lmax, F1));
    if (packed5) {                                                    // 5/8 of the bytes over PCIe, expanded on the device (below)
      const int64_t g_lo = (offsets[s_lo] - 1) / 64, g_hi = (offsets[s_hi - 1] + lengths[s_hi - 1] + 1 + 63) / 64;
      S_CUDA(cudaMemcpyAsync(d_pack + g_lo * 40, packed5 + g_lo * 40, (size_t) (g_hi - g_lo) * 40, cudaMemcpyHostToDevice, cs));
    } else {
      const int64_t b_lo = (offsets[s_lo] - 1) & ~15LL;
      const int64_t b_hi = std::min<int64_t>(nbytes, (offsets[s_hi - 1] + lengths[s_hi - 1] + 1 + 15) & ~15LL);
      S_CUDA(cudaMemcpyAsync(d_res + b_lo, residues + b_lo, (size_t) (b_hi - b_lo), cudaMemcpyHostToDevice, cs));
    }
    S_CUDA(cudaMemcpyAsync(d_off + s_lo, offsets + s_lo, 8 * (size_t) cn, cudaMemcpyHostToDevice, cs));
    S_CUDA(cudaMemcpyAsync(d_len + s_lo, lengths + s_lo, 4 * (size_t) cn, cudaMemcpyHostToDevice, cs));
    cudaEvent_t ev; S_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); events.push_back(ev);
    S_CUDA(cudaEventRecord(ev, cs));
    if (packed5) {
      // The expansion kernel needs SM resources, which the persistent SSV grid of the previous chunk may hold until it ends.  On the
      // copy stream it would hold back every later copy with it (and the copies lead the kernels by little once the chunks
      // double); on a stream of its own only the chunk it expands waits for it.
      const int64_t g_lo = (offsets[s_lo] - 1) / 64, g_hi = (offsets[s_hi - 1] + lengths[s_hi - 1] + 1 + 63) / 64;
      S_CUDA(cudaStreamWaitEvent(us, ev, 0));
      unpack5_kernel<<<(unsigned) ((g_hi - g_lo + 127) / 128), 128, 0, us>>>((const uint32_t *) (d_pack + g_lo * 40), (uint4 *) (d_res + g_lo * 64), g_hi - g_lo);
      ctx->launches++;
      S_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); events.push_back(ev);
      S_CUDA(cudaEventRecord(ev, us));
    }
    const int lane = 1 + (int) (c & 1);
    const Lane &ln = ctx->lanes[lane];
    S_CUDA(cudaStreamWaitEvent(ln.stream, ev, 0));
    // Work order.  The persistent SSV grid of the previous chunk leaves a sort no SM until it ends, so a per-chunk sort is
    // exposed at every chunk boundary (0.2 ms each); in database order the next chunk's CTAs fill the SMs as the previous
    // kernel's CTAs retire, and the accesses to offsets/lengths/outputs are sequential.  Only the last chunk, whose tail
    // nothing overlaps, is taken longest target first.
    const bool sorted = (c + 1 == plan.size());
    if (sorted) {
      iota_from_kernel<<<(unsigned) ((cn + 255) / 256), 256, 0, ln.stream>>>(d_iota[lane], (int32_t) s_lo, (int32_t) cn);
      S_CUDA(cub::DeviceRadixSort::SortPairsDescending(d_sorttmp[lane], sort_tmp_bytes, d_len + s_lo, d_keys_out[lane],
                                                       d_iota[lane], db->d_order + s_lo, (int) cn, 0, 32, ln.stream));
      ctx->launches += 4;
    }
    DbView v; v.res = d_res; v.off = d_off; v.len = d_len; v.order = sorted ? db->d_order + s_lo : nullptr; v.first = s_lo; v.max_length = Lmax;
    unsigned int *d_nores_count = ln.counters + CTR_NORES_COUNT;
    S_CHECK(launch_ssv(ctx, lane, prof, v, nullptr, cn, d_sc, d_st, d_nores[lane], d_nores_count, d_pass, d_pass_count));
    S_CHECK(launch_msv_j(ctx, lane, prof, v, d_nores[lane], d_nores_count, cn, d_sc, d_st, d_pass, d_pass_count));
    if (h_sc)     S_CUDA(cudaMemcpyAsync(h_sc + s_lo, d_sc + s_lo, 4 * (size_t) cn, cudaMemcpyDeviceToHost, ln.stream));
    if (h_status) S_CUDA(cudaMemcpyAsync(h_status + s_lo, d_st + s_lo, 4 * (size_t) cn, cudaMemcpyDeviceToHost, ln.stream));
  }
  // the main stream continues after both lanes and the copy stream
  for (int l = 0; l <= 3; l++) {
    cudaEvent_t ev; S_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); events.push_back(ev);
    S_CUDA(cudaEventRecord(ev, l == 0 ? cs : ctx->lanes[l].stream));
    S_CUDA(cudaStreamWaitEvent(ms, ev, 0));
  }
  tr.mark("stream: all chunks queued");
  S_CUDA(cudaStreamSynchronize(ms));
  tr.mark("stream: drained");
  cleanup();
#undef S_CUDA
#undef S_CHECK
  // the per-chunk orders together are not a global longest-first order; later stages use explicit index lists
  cudaFreeAsync(db->d_order, ms); db->d_order = nullptr;
  db->nres = nres; db->max_length = Lmax;
  *ret_db = db;
  return P7GPU_OK;
}

static int read_count(p7gpu_ctx *ctx, const unsigned int *d_count, unsigned *out)
{
  P7G_CUDA(cudaMemcpyAsync(out, d_count, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
  P7G_CUDA(cudaStreamSynchronize(ctx->stream));
  return P7GPU_OK;
}

static int check_opts(const p7gpu_profile *prof, const p7gpu_pipeline_opts *o)
{
  if (o->stop_after_msv) return P7GPU_OK;
  if (!prof->d_rw || !prof->d_rf) { set_error("pipeline: the profile was uploaded without its Viterbi/Forward parts"); return P7GPU_EINVAL; }
  if (o->do_biasfilter && !o->do_max && !prof->have_bias) { set_error("pipeline: do_biasfilter needs p7gpu_profile_set_biasfilter()"); return P7GPU_EINVAL; }
  return P7GPU_OK;
}

}  // namespace p7g
using namespace p7g;

extern "C" {

int p7gpu_pipeline_run(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db, const p7gpu_pipeline_opts *opts,
                       p7gpu_pipeline_result **ret_res)
{
  if (!ctx || !prof || !db || !opts || !ret_res) return P7GPU_EINVAL;
  *ret_res = nullptr;
  P7G_CHECK(check_opts(prof, opts));
  P7G_CUDA(cudaSetDevice(ctx->device));
  const int64_t n = db->nseq;
  const double F1 = opts->do_max ? 1.0 : opts->F1;
  P7G_CHECK(ensure_len_table(prof, db->max_length, F1));
  p7gpu_pipeline_result *res = new p7gpu_pipeline_result();
  memset(&res->counters, 0, sizeof(res->counters));
  int status = P7GPU_OK;
  {
    DevBuf d_sc, d_st, d_pass, d_nores;
    cudaStream_t st = ctx->stream;
    unsigned npass = 0;
    unsigned int *d_pass_count = ctx->lanes[0].counters + CTR_PASS_COUNT;
    unsigned int *d_nores_count = ctx->lanes[0].counters + CTR_NORES_COUNT;
    Trace tr;
    auto body = [&]() -> int {
      const size_t nn = (size_t) std::max<int64_t>(n, 1);
      P7G_CHECK(d_sc.alloc(4 * nn, st)); P7G_CHECK(d_st.alloc(4 * nn, st)); P7G_CHECK(d_pass.alloc(4 * nn, st)); P7G_CHECK(d_nores.alloc(16 * nn, st));
      P7G_CUDA(cudaMemsetAsync(d_pass_count, 0, sizeof(unsigned int), st));
      P7G_CHECK(launch_ssv(ctx, 0, prof, db->view(), nullptr, n, d_sc.as<float>(), d_st.as<int32_t>(), d_nores.as<int64_t>(), d_nores_count,
                           d_pass.as<int32_t>(), d_pass_count));
      P7G_CHECK(launch_msv_j(ctx, 0, prof, db->view(), d_nores.as<int64_t>(), d_nores_count, n, d_sc.as<float>(), d_st.as<int32_t>(),
                             d_pass.as<int32_t>(), d_pass_count));
      tr.mark("run: msv queued");
      P7G_CHECK(read_count(ctx, d_pass_count, &npass));
      tr.mark("run: msv done");
      return pipeline_after_msv(ctx, prof, db, opts, d_sc.as<float>(), d_st.as<int32_t>(), d_pass.as<int32_t>(), npass, res);
    };
    status = body();
  }
  if (status != P7GPU_OK) { cudaStreamSynchronize(ctx->stream); delete res; return status; }
  *ret_res = res;
  return P7GPU_OK;
}

// Many profiles against one resident database: the outer loop over queries of hmmsearch.c:471-560 (for hmmscan, the roles of
// query and target database swapped, the loop over profiles of hmmscan.c) with the per-query rewind and re-read of the
// sequence file (hmmsearch.c:480-487) gone — the database stays in HBM, each profile costs its tables (<= 350 KB) and one
// per-length constant table.
int p7gpu_pipeline_run_many(p7gpu_ctx *ctx, const p7gpu_profile *const *profs, int64_t nprof, const p7gpu_seqdb *db,
                            const p7gpu_pipeline_opts *opts, p7gpu_pipeline_result **ret_res)
{
  if (!ctx || !profs || nprof < 0 || !db || !opts || !ret_res) return P7GPU_EINVAL;
  for (int64_t i = 0; i < nprof; i++) ret_res[i] = nullptr;
  for (int64_t i = 0; i < nprof; i++) {
    int st = profs[i] ? p7gpu_pipeline_run(ctx, profs[i], db, opts, &ret_res[i]) : P7GPU_EINVAL;
    if (st != P7GPU_OK) {
      for (int64_t j = 0; j < i; j++) { p7gpu_result_destroy(ret_res[j]); ret_res[j] = nullptr; }
      return st;
    }
  }
  return P7GPU_OK;
}

static int run_host_impl(p7gpu_ctx *ctx, const p7gpu_profile *prof, const uint8_t *residues, const uint8_t *packed5, int64_t nbytes,
                         const int64_t *offsets, const int32_t *lengths, int64_t nseq,
                         const p7gpu_pipeline_opts *opts, p7gpu_pipeline_result **ret_res)
{
  if (!ctx || !prof || !opts || !ret_res || nseq < 0 || (nseq > 0 && ((!residues && !packed5) || !offsets || !lengths))) return P7GPU_EINVAL;
  *ret_res = nullptr;
  P7G_CHECK(check_opts(prof, opts));
  P7G_CUDA(cudaSetDevice(ctx->device));
  p7gpu_pipeline_result *res = new p7gpu_pipeline_result();
  memset(&res->counters, 0, sizeof(res->counters));
  if (nseq == 0) { *ret_res = res; return P7GPU_OK; }
  const double F1 = opts->do_max ? 1.0 : opts->F1;
  int status = P7GPU_OK;
  p7gpu_seqdb *db = nullptr;
  {
    DevBuf d_sc, d_st, d_pass;
    cudaStream_t st = ctx->stream;
    unsigned npass = 0;
    unsigned int *d_pass_count = ctx->lanes[0].counters + CTR_PASS_COUNT;
    Trace tr;
    auto body = [&]() -> int {
      P7G_CHECK(d_sc.alloc(4 * (size_t) nseq, st)); P7G_CHECK(d_st.alloc(4 * (size_t) nseq, st)); P7G_CHECK(d_pass.alloc(4 * (size_t) nseq, st));
      tr.mark("alloc");
      P7G_CHECK(stream_msv(ctx, prof, residues, packed5, nbytes, offsets, lengths, nseq, F1, nullptr, nullptr,
                           d_sc.as<float>(), d_st.as<int32_t>(), d_pass.as<int32_t>(), d_pass_count, &db));
      tr.mark("stream_msv");
      P7G_CHECK(read_count(ctx, d_pass_count, &npass));
      int rc = pipeline_after_msv(ctx, prof, db, opts, d_sc.as<float>(), d_st.as<int32_t>(), d_pass.as<int32_t>(), npass, res);
      tr.mark("after_msv");
      return rc;
    };
    status = body();
    cudaStreamSynchronize(st);
    tr.mark("sync");
    p7gpu_seqdb_destroy(db); db = nullptr;
    tr.mark("db destroy");
  }
  p7gpu_seqdb_destroy(db);
  if (status != P7GPU_OK) { delete res; return status; }
  *ret_res = res;
  return P7GPU_OK;
}

int p7gpu_pipeline_run_host(p7gpu_ctx *ctx, const p7gpu_profile *prof, const uint8_t *residues, int64_t nbytes,
                            const int64_t *offsets, const int32_t *lengths, int64_t nseq,
                            const p7gpu_pipeline_opts *opts, p7gpu_pipeline_result **ret_res)
{
  if (nseq > 0 && !residues) return P7GPU_EINVAL;
  return run_host_impl(ctx, prof, residues, nullptr, nbytes, offsets, lengths, nseq, opts, ret_res);
}

int p7gpu_pipeline_run_host5(p7gpu_ctx *ctx, const p7gpu_profile *prof, const uint8_t *packed5, int64_t nbytes,
                             const int64_t *offsets, const int32_t *lengths, int64_t nseq,
                             const p7gpu_pipeline_opts *opts, p7gpu_pipeline_result **ret_res)
{
  if (nseq > 0 && !packed5) return P7GPU_EINVAL;
  return run_host_impl(ctx, prof, nullptr, packed5, nbytes, offsets, lengths, nseq, opts, ret_res);
}

// ------------------------------------------------------------------------------------------------------------------
// Several GPUs from one C host (include/p7gpu.h section 9): contiguous shards balanced by residue count, one host thread and
// one context per shard, no data-path exchange; the caller merges counters (p7gpu_counters_merge) and hit lists.
// ------------------------------------------------------------------------------------------------------------------
int p7gpu_shard_plan(const int64_t *offsets, const int32_t *lengths, int64_t nseq, int nshards, int64_t *first)
{
  if (nshards < 1 || nseq < 0 || !first || (nseq > 0 && (!offsets || !lengths))) return P7GPU_EINVAL;
  first[0] = 0;
  if (nseq == 0) { for (int k = 1; k <= nshards; k++) first[k] = 0; return P7GPU_OK; }
  // offsets grow with the residues before a target (plus the same padding per target everywhere): cut them evenly
  const int64_t lo0 = offsets[0], span = offsets[nseq - 1] + lengths[nseq - 1] - lo0;
  for (int k = 1; k < nshards; k++) {
    const int64_t want = lo0 + (int64_t) ((double) span * k / nshards);
    int64_t lo = first[k - 1], hi = nseq;
    while (lo < hi) { const int64_t mid = (lo + hi) / 2; if (offsets[mid] < want) lo = mid + 1; else hi = mid; }
    first[k] = lo;
  }
  first[nshards] = nseq;
  return P7GPU_OK;
}

void p7gpu_counters_merge(p7gpu_pipeline_counters *into, const p7gpu_pipeline_counters *from)
{
  if (!into || !from) return;
  into->nseqs += from->nseqs; into->nres += from->nres; into->n_past_msv += from->n_past_msv;
  into->n_past_bias += from->n_past_bias; into->n_past_vit += from->n_past_vit; into->n_past_fwd += from->n_past_fwd;
}

int p7gpu_pipeline_run_sharded(p7gpu_ctx *const *ctxs, const p7gpu_profile *const *profs, int nshards,
                               const uint8_t *residues, const uint8_t *packed5, int64_t nbytes,
                               const int64_t *offsets, const int32_t *lengths, int64_t nseq,
                               const p7gpu_pipeline_opts *opts, int64_t *first, p7gpu_pipeline_result **ret_res)
{
  if (!ctxs || !profs || nshards < 1 || !opts || !first || !ret_res || nseq < 0 || (nseq > 0 && ((!residues && !packed5) || !offsets || !lengths))) return P7GPU_EINVAL;
  for (int k = 0; k < nshards; k++) { ret_res[k] = nullptr; if (!ctxs[k] || !profs[k] || profs[k]->ctx != ctxs[k]) { set_error("run_sharded: shard %d: profile and context do not belong together", k); return P7GPU_EINVAL; } }
  P7G_CHECK(p7gpu_shard_plan(offsets, lengths, nseq, nshards, first));
  std::vector<int> rc((size_t) nshards, P7GPU_OK);
  std::vector<std::string> err((size_t) nshards);
  auto work = [&](int k) {
    const int64_t s_lo = first[k], s_hi = first[k + 1], cn = s_hi - s_lo;
    if (cn == 0) { rc[k] = run_host_impl(ctxs[k], profs[k], residues, packed5, 0, nullptr, nullptr, 0, opts, &ret_res[k]); return; }
    // the shard's slice of the packed array starts on a 64-byte boundary at or before its first target's leading sentinel
    // (64 bytes = one group of the 5-bit transfer format; the kernels' 16-byte residue vectors keep their alignment)
    const int64_t base = (offsets[s_lo] - 1) & ~(int64_t) 63;
    const int64_t end  = std::min<int64_t>(nbytes, offsets[s_hi - 1] + lengths[s_hi - 1] + 1);
    std::vector<int64_t> off((size_t) cn);
    for (int64_t j = 0; j < cn; j++) off[(size_t) j] = offsets[s_lo + j] - base;
    if (off[0] < 1) { err[k] = "run_sharded: no sentinel byte before the shard's first target"; rc[k] = P7GPU_EINVAL; return; }
    rc[k] = run_host_impl(ctxs[k], profs[k], residues ? residues + base : nullptr, packed5 ? packed5 + base / 64 * 40 : nullptr, end - base,
                          off.data(), lengths + s_lo, cn, opts, &ret_res[k]);
    if (rc[k] != P7GPU_OK) { err[k] = p7gpu_last_error(); return; }
    p7gpu_pipeline_result *r = ret_res[k];                           // shard numbering -> database numbering
    for (auto &sv : r->survivors) sv.seqidx += s_lo;
    for (auto &mv : r->msv_pass) mv.seqidx += s_lo;
  };
  if (nshards == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int k = 0; k < nshards; k++) th.emplace_back(work, k);
    for (auto &t : th) t.join();
  }
  for (int k = 0; k < nshards; k++)
    if (rc[k] != P7GPU_OK) {
      set_error("%s", err[k].c_str());
      for (int j = 0; j < nshards; j++) { p7gpu_result_destroy(ret_res[j]); ret_res[j] = nullptr; }
      return rc[k];
    }
  return P7GPU_OK;
}

// Diagnostic: the float threshold the device-side stage decisions use for "P(seq_score) > F" (kind 0: esl_gumbel_surv with
// (mu, lambda); kind 1: esl_exp_surv).  Returns P7GPU_OK and *T = the smallest seq_score that passes, or P7GPU_ERANGE when no
// verified threshold exists (the pipeline then evaluates that test on the host).  Host code; needs no device.
int p7gpu_stage_threshold(int kind, double F, float mu, float lambda, float *T)
{
  if (!T || !(F > 0.0 && F < 1.0) || !(lambda > 0.0f)) return P7GPU_EINVAL;
  bool ok;
  if (kind == 0) ok = pvalue_threshold([&](float x) { return !(host_gumbel_surv((double) x, mu, lambda) > F); }, (float) (mu - log(-log1p(-F)) / lambda), T);
  else           ok = pvalue_threshold([&](float x) { return !(host_exp_surv((double) x, mu, lambda) > F); }, (float) (mu - log(F) / lambda), T);
  return ok ? P7GPU_OK : P7GPU_ERANGE;
}

int64_t p7gpu_pack5_bytes(int64_t nbytes) { return (nbytes + 63) / 64 * 40 + 40; }

// residues (codes 0..28 and the 255 sentinel) -> the 5-bit transfer format; out must hold p7gpu_pack5_bytes(nbytes) bytes
int p7gpu_pack5(const uint8_t *residues, int64_t nbytes, uint8_t *out, int nthreads)
{
  if (nbytes < 0 || (nbytes > 0 && (!residues || !out))) return P7GPU_EINVAL;
  const int64_t ngroups = p7gpu_pack5_bytes(nbytes) / 5;
  unsigned hw = std::thread::hardware_concurrency();
  int64_t T = nthreads > 0 ? nthreads : (int64_t) std::min<unsigned>(hw ? hw : 1, 32);
  if (T > ngroups / 65536 + 1) T = ngroups / 65536 + 1;
  std::atomic<int> bad(0);
  auto work = [&](int64_t g_lo, int64_t g_hi) {
    for (int64_t g = g_lo; g < g_hi; g++) {
      unsigned long long v = 0;
      for (int r = 0; r < 8; r++) {
        const int64_t i = g * 8 + r;
        unsigned c = i < nbytes ? residues[i] : 255u;
        if (c == 255u) c = 31u; else if (c > 28u) { bad = 1; c = 31u; }
        v |= (unsigned long long) c << (5 * r);
      }
      for (int b = 0; b < 5; b++) out[g * 5 + b] = (uint8_t) (v >> (8 * b));
    }
  };
  std::vector<std::thread> th;
  for (int64_t k = 0; k < T; k++) th.emplace_back(work, ngroups * k / T, ngroups * (k + 1) / T);
  for (auto &x : th) x.join();
  if (bad) { set_error("pack5: a residue code outside 0..28 / 255"); return P7GPU_EINVAL; }
  return P7GPU_OK;
}

int p7gpu_msvfilter_stream(p7gpu_ctx *ctx, const p7gpu_profile *prof, const uint8_t *residues, int64_t nbytes,
                           const int64_t *offsets, const int32_t *lengths, int64_t nseq, double F1,
                           float *sc, int32_t *status, int32_t *pass_list, int64_t *npass_out, p7gpu_seqdb **opt_ret_db)
{
  if (!ctx || !prof || nseq < 0 || (nseq > 0 && (!residues || !offsets || !lengths))) return P7GPU_EINVAL;
  if (opt_ret_db) *opt_ret_db = nullptr;
  if (npass_out) *npass_out = 0;
  if (nseq == 0) return P7GPU_OK;
  P7G_CUDA(cudaSetDevice(ctx->device));
  p7gpu_seqdb *db = nullptr;
  int rc = P7GPU_OK;
  {
    DevBuf d_sc, d_st, d_pass;
    cudaStream_t st = ctx->stream;
    unsigned int *d_pass_count = ctx->lanes[0].counters + CTR_PASS_COUNT;
    auto body = [&]() -> int {
      P7G_CHECK(d_sc.alloc(4 * (size_t) nseq, st)); P7G_CHECK(d_st.alloc(4 * (size_t) nseq, st)); P7G_CHECK(d_pass.alloc(4 * (size_t) nseq, st));
      P7G_CHECK(stream_msv(ctx, prof, residues, nullptr, nbytes, offsets, lengths, nseq, F1, sc, status,
                           d_sc.as<float>(), d_st.as<int32_t>(), d_pass.as<int32_t>(), d_pass_count, &db));
      unsigned npass = 0;
      P7G_CHECK(read_count(ctx, d_pass_count, &npass));
      if (npass_out) *npass_out = npass;
      if (pass_list && npass) {
        P7G_CUDA(cudaMemcpyAsync(pass_list, d_pass.p, 4 * (size_t) npass, cudaMemcpyDeviceToHost, st));
        P7G_CUDA(cudaStreamSynchronize(st));
        std::sort(pass_list, pass_list + npass);
      }
      return P7GPU_OK;
    };
    rc = body();
    cudaStreamSynchronize(st);
  }
  if (rc != P7GPU_OK || !opt_ret_db) p7gpu_seqdb_destroy(db);
  else *opt_ret_db = db;
  return rc;
}

const p7gpu_pipeline_counters *p7gpu_result_counters(const p7gpu_pipeline_result *res) { return res ? &res->counters : nullptr; }
int64_t p7gpu_result_nsurvivors(const p7gpu_pipeline_result *res) { return res ? (int64_t) res->survivors.size() : 0; }
const p7gpu_survivor *p7gpu_result_survivors(const p7gpu_pipeline_result *res) { return res ? res->survivors.data() : nullptr; }
const float *p7gpu_result_fwd_xmx(const p7gpu_pipeline_result *res) { return res ? res->fwd_xmx.data() : nullptr; }
const float *p7gpu_result_bck_xmx(const p7gpu_pipeline_result *res) { return res ? res->bck_xmx.data() : nullptr; }
const float *p7gpu_result_decoding(const p7gpu_pipeline_result *res) { return res ? res->dec.data() : nullptr; }
const uint64_t *p7gpu_result_stage_residues(const p7gpu_pipeline_result *res) { return res ? res->stage_nres : nullptr; }
int64_t p7gpu_result_nmsvpass(const p7gpu_pipeline_result *res) { return res ? (int64_t) res->msv_pass.size() : 0; }
const p7gpu_msv_survivor *p7gpu_result_msvpass(const p7gpu_pipeline_result *res) { return res ? res->msv_pass.data() : nullptr; }
void p7gpu_result_destroy(p7gpu_pipeline_result *res)
{
  if (!res) return;
  std::lock_guard<std::mutex> lk(ctx_registry_mutex());            // the context may be gone already (host-language GC order)
  if (res->ctx && ctx_registered_locked(res->ctx)) {
    HostScratch &hs = host_scratch(res->ctx);
    if (res->msv_pass.capacity() > hs.msv_cache.capacity()) hs.msv_cache.swap(res->msv_pass);
    if (res->fwd_xmx.cap > hs.xmx_cache[0].cap) hs.xmx_cache[0].swap(res->fwd_xmx);
    if (res->bck_xmx.cap > hs.xmx_cache[1].cap) hs.xmx_cache[1].swap(res->bck_xmx);
  }
  delete res;
}

}
