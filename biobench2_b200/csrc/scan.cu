// scan.cu — host side of hmmscan's block prefilter (p7gpu_scan_msv_prefilter, include/p7gpu.h; kernel: ssv_msv.cu ssv_scan_kernel).
//
// The reference's hmmscan reads one model's MSV part, runs p7_MSVFilter against the query, and reads the rest of the model
// only for a survivor (HMMER/src/hmmscan.c:1176-1198, p7_pipeline.c:688-713, impl_sse/io.c:233,493).  Here the MSV parts of a
// whole block of models are turned into scan tables side by side in one device buffer and ONE launch scores every
// (model, query) pair; the caller uploads and runs the full pipeline only for the models some query may get past MSV of.
#include "p7gpu_internal.h"
#include <algorithm>
#include <thread>

using namespace p7g;

namespace {

// SSV gain of residue x at node k (1..M): -(signed byte of om->sbv), as p7gpu_profile_upload derives it (p7_oprofile.c:711-752)
inline int16_t ssv_gain(const p7gpu_oprofile_view *om, int Q16, int x, int k)
{
  const int q = (k - 1) % Q16, z = (k - 1) / Q16;
  int8_t sb;
  if (om->sbv) sb = (int8_t) om->sbv[((size_t) x * (Q16 + 17) + q) * 16 + z];
  else {
    const uint8_t rb = om->rbv[((size_t) x * Q16 + q) * 16 + z];
    int a = (127 + om->bias_b) - rb; if (a < 0) a = 0;
    sb = (int8_t) ((uint8_t) a ^ 127);
  }
  return (int16_t) (-(int) sb);
}

struct ScanGeom { int npass, P, rowWords; size_t words; };
inline ScanGeom scan_geom(int M)
{
  ScanGeom g;
  g.npass = (M + 1 + 255) / 256;                                       // P >= M + 1: at least one pad column resets a wrapping diagonal
  g.P = 256 * g.npass;
  g.rowWords = g.P + 96;
  g.words = (size_t) P7GPU_KP * g.rowWords;
  return g;
}

void build_scan_table(const p7gpu_oprofile_view *om, const ScanGeom &g, uint32_t *out)
{
  const int M = om->M, Q16 = std::max(2, (M - 1) / 16 + 1);            // impl_sse.h:25
  std::vector<int16_t> col((size_t) g.P);
  for (int x = 0; x < P7GPU_KP; x++) {
    for (int c = 0; c < g.P; c++) col[c] = c < M ? ssv_gain(om, Q16, x, c + 1) : (int16_t) SSV_PADGAIN;
    uint32_t *row = out + (size_t) x * g.rowWords;
    int lo = 0, hi = 128 % g.P;                                          // w mod P and (w + 128) mod P, kept by increment
    for (int w = 0; w < g.rowWords; w++) {
      row[w] = (uint32_t) (uint16_t) col[lo] | ((uint32_t) (uint16_t) col[hi] << 16);
      if (++lo == g.P) lo = 0;
      if (++hi == g.P) hi = 0;
    }
  }
}

}  // namespace

extern "C" int p7gpu_scan_table_build(const p7gpu_oprofile_view *om, int32_t geom[3], uint32_t *table, int32_t L, double F1, uint32_t *len_word)
{
  if (!om || !geom || om->M < 1 || om->Kp != P7GPU_KP || !om->rbv || om->M > 64 * MAX_Q - 1 || L < 0) return P7GPU_EINVAL;
  const ScanGeom g = scan_geom(om->M);
  geom[0] = g.P; geom[1] = g.rowWords; geom[2] = g.npass;
  if (table) build_scan_table(om, g, table);
  if (len_word) scan_len_words(om, &L, 1, F1 < 0 ? DEFAULT_F1 : F1, len_word);
  return P7GPU_OK;
}

extern "C" int p7gpu_scan_msv_prefilter(p7gpu_ctx *ctx, const p7gpu_oprofile_view *oms, int64_t nmodels, const p7gpu_seqdb *queries,
                                        const int32_t *lengths, double F1, int32_t *ret_npossible)
{
  if (!ctx || nmodels < 0 || (nmodels > 0 && (!oms || !ret_npossible)) || !queries || (queries->nseq > 0 && !lengths)) return P7GPU_EINVAL;
  if (nmodels == 0) return P7GPU_OK;
  const int64_t nq = queries->nseq;
  for (int64_t m = 0; m < nmodels; m++) {
    ret_npossible[m] = 0;
    if (oms[m].M < 1 || oms[m].Kp != P7GPU_KP || !oms[m].rbv) { set_error("scan_msv_prefilter: model %lld is not an amino profile with M>=1", (long long) m); return P7GPU_EINVAL; }
    if (oms[m].M > 64 * MAX_Q - 1) { set_error("scan_msv_prefilter: M=%d exceeds the supported maximum %d", oms[m].M, 64 * MAX_Q - 1); return P7GPU_EINVAL; }
  }
  if (nq == 0) return P7GPU_OK;
  if (F1 < 0) F1 = DEFAULT_F1;
  P7G_CUDA(cudaSetDevice(ctx->device));

  // the distinct query lengths and each query's index among them
  std::vector<int32_t> lens(lengths, lengths + nq);
  std::sort(lens.begin(), lens.end());
  lens.erase(std::unique(lens.begin(), lens.end()), lens.end());
  const int nlens = (int) lens.size();
  std::vector<int32_t> lenidx((size_t) nq);
  for (int64_t s = 0; s < nq; s++) {
    if (lengths[s] < 0) return P7GPU_EINVAL;
    lenidx[s] = (int32_t) (std::lower_bound(lens.begin(), lens.end(), lengths[s]) - lens.begin());
  }

  // tables, scalars and per-length words of every model, built on all host threads
  std::vector<ScanModel> models((size_t) nmodels);
  std::vector<ScanGeom> geoms((size_t) nmodels);
  size_t total = 0;
  for (int64_t m = 0; m < nmodels; m++) {
    geoms[m] = scan_geom(oms[m].M);
    if (total + geoms[m].words > 0xffffffffull) { set_error("scan_msv_prefilter: block too large (%lld models); pass fewer per call", (long long) nmodels); return P7GPU_EINVAL; }
    ScanModel &d = models[m];
    d.tab_off = (uint32_t) total; d.rowWords = geoms[m].rowWords; d.P = geoms[m].P; d.npass = geoms[m].npass;
    d.s.tbm_b = oms[m].tbm_b; d.s.tec_b = oms[m].tec_b; d.s.base_b = oms[m].base_b; d.s.bias_b = oms[m].bias_b; d.s.scale_b = oms[m].scale_b;
    total += geoms[m].words;
  }
  const size_t b_tabs = total * 4, b_models = sizeof(ScanModel) * (size_t) nmodels, b_lc = 4 * (size_t) nmodels * nlens,
               b_lenidx = 4 * (size_t) nq, b_out = 4 * (size_t) nmodels;
  auto up = [](size_t b) { return (b + 255) & ~(size_t) 255; };
  const size_t o_tabs = 0, o_models = o_tabs + up(b_tabs), o_lc = o_models + up(b_models), o_lenidx = o_lc + up(b_lc),
               o_out = o_lenidx + up(b_lenidx), o_work = o_out + up(b_out), bytes = o_work + 256;
  // Staging: an ordinary host buffer kept between calls (page-locking tens of MB costs more than the copy it would speed up),
  // grown with head room because the blocks of one scan differ in size by their models' lengths.
  static thread_local std::vector<char> stage;
  if (bytes > stage.size()) stage.resize(bytes + bytes / 2);
  if (bytes > ctx->d_scratch_bytes) P7G_CHECK(ensure_scratch(ctx, bytes + bytes / 2));
  char *h = stage.data(), *d = (char *) ctx->d_scratch;
  {
    uint32_t *h_tabs = (uint32_t *) (h + o_tabs), *h_lc = (uint32_t *) (h + o_lc);
    unsigned hw = std::thread::hardware_concurrency();
    const int T = (int) std::max<int64_t>(1, std::min<int64_t>({(int64_t) (hw ? hw : 1), (int64_t) 16, nmodels / 8}));
    auto work = [&](int t) {
      for (int64_t m = t; m < nmodels; m += T) {
        build_scan_table(&oms[m], geoms[m], h_tabs + models[m].tab_off);
        scan_len_words(&oms[m], lens.data(), nlens, F1, h_lc + (size_t) m * nlens);
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
  }
  memcpy(h + o_models, models.data(), b_models);
  memcpy(h + o_lenidx, lenidx.data(), b_lenidx);
  memset(h + o_out, 0, up(b_out) + 256);                               // per-model counts and the work counter start at zero
  P7G_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));

  ScanParams p;
  p.models = (const ScanModel *) (d + o_models); p.tabs = (const uint32_t *) (d + o_tabs);
  p.lc = (const uint32_t *) (d + o_lc); p.lenidx = (const int32_t *) (d + o_lenidx); p.nlens = nlens;
  p.res = queries->d_res; p.off = queries->d_off; p.len = queries->d_len; p.nq = nq;
  p.nitems = nmodels * nq;
  p.npossible = (int32_t *) (d + o_out);
  p.work_counter = (unsigned long long *) (d + o_work);
  P7G_CHECK(launch_ssv_scan(ctx, ctx->stream, p));
  P7G_CUDA(cudaMemcpyAsync(ret_npossible, d + o_out, b_out, cudaMemcpyDeviceToHost, ctx->stream));
  P7G_CUDA(cudaStreamSynchronize(ctx->stream));
  return P7GPU_OK;
}
