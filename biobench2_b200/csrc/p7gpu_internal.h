// p7gpu_internal.h — internal structures of libp7gpu.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <mutex>
#include <memory>
#include "../../include/p7gpu.h"

namespace p7g {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char *fmt, ...);
#define P7G_CUDA(call)                                                                           \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess) {                                                                    \
      p7g::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));     \
      return P7GPU_ESYS;                                                                         \
    }                                                                                            \
  } while (0)
#define P7G_CHECK(call)                                                                          \
  do { int s__ = (call); if (s__ != P7GPU_OK) return s__; } while (0)

// ---- geometry of the diagonal ("slot") layout used by the SSV and MSV kernels ------------------
// A warp holds 64*Q slots: lane l, register q, half h  <->  diagonal start column l*Q + q + h*32*Q.
// At DP row i the slot sits on column (start + i) mod P, P = 64*Q >= M+1; columns >= M are pad columns
// whose gain is PADGAIN, which resets the running diagonal to the floor before it re-enters at k=1.
// Q odd: lane stride Q words, so the 32 lanes of one LDS.32 hit 32 distinct banks.  Q even (<= MAX_EVEN_Q): start column
// l + 32*q instead (lane stride 1, register stride 32): conflict-free for any Q; the wrap-free row is 32*(Q-1) words longer.
constexpr int DIAG_CHUNK   = 16;      // DP rows per unrolled chunk = one 16-byte residue load
constexpr int PADRES       = 29;      // table row used for masked (out-of-sequence) rows: all PADGAIN
constexpr int NTABROWS     = 30;      // Kp + 1
constexpr int SSV_PADGAIN  = -20000;  // int16; |v| stays < 2^15 (overflow argument)
constexpr int MAX_Q        = 65;      // M <= 64*65-1 = 4159
constexpr int MAX_EVEN_Q   = 14;      // even Q (consecutive-lane layout) up to here

struct DiagGeom {
  int M, Q, P, rowWords;              // rowWords = P + DIAG_CHUNK + Q (odd Q) or P + DIAG_CHUNK + 32*(Q-1) + 1 (even Q): wrap-free reads inside one chunk
  size_t tableBytes() const { return (size_t) NTABROWS * rowWords * 4; }
};
std::mutex &ctx_registry_mutex();                 // live contexts (p7gpu_host.cu): results hand buffers back only to a live one
bool ctx_registered_locked(const p7gpu_ctx *ctx);
DiagGeom diag_geom(int M);
int dp_cells_per_lane(int M);
bool mw_geometry(int M, int *C, int *W);   // Forward/Backward CTA-per-target geometry for long models; false = use one warp
int vit_packed_c2(int M);           // registers per lane of the packed Viterbi kernel (vit.cu), -1 = unsupported       // smallest supported C with 32*C >= M (dp_common.cuh)

// ---- per-target-length constants: what p7_oprofile_ReconfigLength()/p7_bg_SetLength() derive ----
// (impl_sse/p7_oprofile.c:1085-1127, p7_bg.c:190-199,358-363), built on the host with the host libm,
// plus the MSV pass threshold of p7_pipeline.c:691-693 folded into "smallest passing xJ byte" per L.
struct LenConst {
  float    pmove;      // xf[N|C|J][MOVE]; LOOP = 1.0f - pmove
  float    nullsc;     // p7_bg_NullOne
  int16_t  xw_move;    // xw[N|C|J][MOVE]
  uint8_t  tjb_b;
  uint8_t  pad;
  uint16_t msv_xj_min; // smallest xJ (0..255) whose P-value passes F1 for this L; 256 = none; L=0 never passes
  uint16_t pad2;
  float    logf_p1;    // logf(p1), p1 = L/(L+1)          } the length term of p7_bg_FilterScore (p7_bg.c:481)
  float    logf_1mp1;  // logf(1. - p1)                   }
};

// scalar parameters handed to kernels by value
struct MsvScalars { int tbm_b, tec_b, base_b, bias_b; float scale_b; };
struct VitScalars { int xw_E_move, xw_E_loop, base_w, ddbound_w; float scale_w; };
// packed (2 x int16) Viterbi fast path: clamp window and table geometry (vit.cu)
struct VitPacked  { int enabled, C2, T, W0, W1, floff; };
struct FwdScalars { float xf_E_move, xf_E_loop; };
struct BiasHmm    { float t[2][3]; float pi[3]; float eo[P7GPU_KP][2]; };   // bg->fhmm after p7_bg_SetFilter

// A lane = one stream + its own device work counters, so that launches on different lanes may overlap.
constexpr int NLANES = 4;            // lane 0 is the ctx's main stream
constexpr int LANE_COUNTER_WORDS = 32;
struct Lane { cudaStream_t stream = nullptr; unsigned int *counters = nullptr; };

// counter slots inside a lane (unsigned int words; 64-bit work counters take two)
enum { CTR_SSV_WORK = 0, CTR_NORES_COUNT = 2, CTR_MSVJ_WORK = 4, CTR_VIT_WORK = 6, CTR_FWD_WORK = 8, CTR_BWD_WORK = 10,
       CTR_PASS_COUNT = 12, CTR_BIAS_WORK = 14, CTR_VIT_REDO = 16, CTR_VIT_WORK2 = 18 };

// what a kernel needs to know about the target set
struct DbView {
  const uint8_t *res = nullptr;      // residue base such that res[off[s]+i] is residue i of target s
  const int64_t *off = nullptr;
  const int32_t *len = nullptr;
  const int32_t *order = nullptr;    // work order (longest first) or null
  int64_t first = 0;                 // when neither idx nor order: target = first + work index
  int32_t max_length = 0;
};

}  // namespace p7g

// per-launch CUDA-event timing (p7gpu_ctx_set_timing): bench.py's live roofline measurement
struct TimedLaunch { int kind; cudaEvent_t start, stop; };

struct p7gpu_ctx {
  bool         timing = false;
  std::vector<TimedLaunch> timed;
  double       kernel_ms[P7GPU_NKERNELS] = {0};
  int64_t      kernel_launches[P7GPU_NKERNELS] = {0};
  int          device = 0;
  int          sm_count = 0;
  int          smem_optin = 0;
  cudaStream_t stream = nullptr;        // == lanes[0].stream
  bool         own_stream = false;
  p7g::Lane    lanes[p7g::NLANES];
  cudaStream_t copy_stream = nullptr;   // H2D staging for the streamed database path
  int64_t      launches = 0;
  int          vit_exact_only = 0;      // option "vit_exact_only": skip the packed Viterbi fast path (tests)
  int          fb_one_warp = 0;         // option "fb_one_warp": Forward/Backward with the one-warp kernels even for long models (tests)
  int          fb_exact = 0;            // option "fb_exact": p7gpu_forward_batch/p7gpu_backward_batch run the exact-order kernels (fbx.cu)
  int          fb_fast = 0;             // option "fb_fast": the one-target wrappers p7gpu_ForwardParser/BackwardParser run the fast kernels
  int          ssv_tmem = 1;            // option "ssv_tmem": SSV gain table in Tensor Memory when 193 <= M <= 1024 (ssv_tm_kernel; 2 = at every M <= 1024); 0 = shared memory
  int          fbx_warps = 0;           // option "fbx_warps": cap on resident one-warp CTAs per SM of the exact-order kernels (0 = occupancy)
  int          bias_warp = 1;           // option "bias_warp": one warp per target in the bias kernel (0 never, 1 when the launch is small, 2 always)
  float        bias_window = 0.f;       // option "bias_window_log2": half-width 2^-n of the bias kernel's log() certainty test (tests widen it)
  int64_t      bias_fixed = 0;          // targets whose filtersc the host recomputed so far (diagnostic)
  int          fbx_groups = 0;          // option "fbx_groups": targets per warp of the exact-order kernels, 1/2/4/8 (0 = auto)
  // growable scratch buffers (device + pinned host)
  void  *d_scratch = nullptr;  size_t d_scratch_bytes = 0;
  void  *h_pinned  = nullptr;  size_t h_pinned_bytes  = 0;
  void  *d_scratch2 = nullptr; size_t d_scratch2_bytes = 0;
  std::shared_ptr<void> host_scratch;   // pipeline.cu: persistent host working arrays of the batched pipeline
};

struct p7gpu_profile {
  p7gpu_ctx *ctx = nullptr;
  int M = 0, Kp = 0;
  float nj = 0.f;
  float evparam[6];
  p7g::MsvScalars msv;
  p7g::VitScalars vit;
  p7g::FwdScalars fwd;
  // SSV / MSV diagonal tables (device): [NTABROWS][rowWords] words = (int16 gain lo, int16 gain hi)
  p7g::DiagGeom geom;
  void     *d_block = nullptr;         // the one device allocation all tables below point into (p7gpu_profile_upload)
  uint32_t *d_ssv_tab = nullptr;
  uint32_t *d_ssv_tm = nullptr;        // SSV gains in the Tensor Memory layout [NTABROWS][NX][32] (M <= 1024), see ssv_tm_kernel
  int ssv_tm_Q = 0;
  uint32_t *d_msv_tab = nullptr;
  // Viterbi (int16) and Forward/Backward (f32) k-space tables (device)
  int       dpC = 0;                   // cells per lane of the row-synchronous kernels (dp_common.cuh)
  int16_t  *d_rw = nullptr;            // [NTABROWS][dpC][32]  lane-transposed: node k = lane*dpC + j + 1
  int16_t  *d_tw = nullptr;            // [8][dpC][32]
  p7g::VitPacked vitp = {0, 0, 0, 0, 0, 0};
  uint32_t *d_rwp = nullptr;           // [NTABROWS][C2][32] (lo = node lane*C2+j+1, hi = that + 32*C2), clamped at -T
  uint32_t *d_twp = nullptr;           // [8][C2][32]
  uint32_t *d_scanw = nullptr;         // [6][32] lane-span sums of tDD for the packed D->D scan
  float    *d_rf = nullptr;            // [NTABROWS][dpC][32]
  float    *d_tf = nullptr;            // [8][dpC][32]
  // long models (M > 512): Forward/Backward with one CTA of mwW warps per target (fwdbwd_mw.cuh); 0 = not used
  int       mwC = 0, mwW = 0;            // cells per lane, warps per CTA
  float    *d_rf_mw = nullptr;         // [NTABROWS][mwC][32*mwW]  node k = g*mwC + j + 1
  float    *d_tf_mw = nullptr;         // [8][mwC][32*mwW]
  // exact-order Forward/Backward (fbx.cu): the reference's striped values, re-laid out for load width only
  int       fbxQ = 0, fbxQp = 0;       // Q = p7O_NQF(M); Qp = Q rounded up to a multiple of 4
  float    *d_tfx = nullptr;           // [Q][4][8]: BM,MM,IM,DM,MD,MI,II,DD of vector q, element z
  float    *d_rfx = nullptr;           // [NTABROWS][4][Qp]
  float    *d_tddx = nullptr, *d_tmdx = nullptr;   // [4][Qp] each: tDD / tMD of vector q, element z
  // bias-filter HMM (p7_bg_SetFilter), optional
  bool        have_bias = false;
  p7g::BiasHmm bias;
  // per-L constants (valid for the F1 they were built with)
  mutable std::mutex  len_mutex;
  mutable int         len_cap = 0;
  mutable double      len_F1 = -1.0;
  mutable bool        len_sparse = false;       // built by p7gpu_profile_prepare_lengths: only the prepared lengths are valid
  mutable p7g::LenConst *d_len = nullptr;
  mutable std::vector<p7g::LenConst> h_len;
  float scale_b = 0, scale_w = 0;
};

struct p7gpu_seqdb {
  p7gpu_ctx *ctx = nullptr;
  const uint8_t *d_res = nullptr;      // padded to a 16-byte multiple + 16 slack bytes
  const int64_t *d_off = nullptr;
  const int32_t *d_len = nullptr;
  int32_t       *d_order = nullptr;    // targets sorted by length, longest first (work order); may be null
  int64_t nseq = 0, nbytes = 0, nres = 0;
  int32_t max_length = 0;
  bool owns = false;
  bool pooled = false;                 // buffers came from cudaMallocAsync on ctx->stream (streamed path)
  p7g::DbView view() const { p7g::DbView v; v.res = d_res; v.off = d_off; v.len = d_len; v.order = d_order; v.first = 0; v.max_length = max_length; return v; }
};

namespace p7g {
// brackets one kernel launch with events on the launching stream when ctx->timing is on
struct LaunchTimer {
  p7gpu_ctx *ctx; int kind; cudaStream_t stream; cudaEvent_t start = nullptr, stop = nullptr;
  LaunchTimer(p7gpu_ctx *c, int k, cudaStream_t s) : ctx(c), kind(k), stream(s) {
    if (ctx->timing) { cudaEventCreate(&start); cudaEventCreate(&stop); cudaEventRecord(start, stream); }
  }
  ~LaunchTimer() {
    if (start) { cudaEventRecord(stop, stream); ctx->timed.push_back({kind, start, stop}); }
  }
};

constexpr double DEFAULT_F1 = 0.02;
int ensure_len_table(const p7gpu_profile *prof, int Lmax, double F1 = -1.0);   // F1 < 0: keep the current table's F1
int ensure_scratch(p7gpu_ctx *ctx, size_t bytes);
int ensure_scratch2(p7gpu_ctx *ctx, size_t bytes);
int ensure_pinned(p7gpu_ctx *ctx, size_t bytes);
int build_order(p7gpu_ctx *ctx, p7gpu_seqdb *db);

// the reference's scalar score -> P-value arithmetic, host libm (p7_pipeline.c:691-730)
float  host_nullsc(int L);
double host_gumbel_surv(double x, double mu, double lambda);
double host_exp_surv(double x, double mu, double lambda);
double host_filter_pvalue_gumbel(float sc, float nullsc_or_filtersc, float mu, float lambda);
double host_filter_pvalue_exp(float sc, float filtersc, float tau, float lambda);

// kernel launchers (ssv_msv.cu, vit.cu, fwdbwd.cu, bias.cu).  Work item j of n maps to target
//   idx ? idx[j] : (db.order ? db.order[j] : db.first + j);  outputs are indexed by j when idx != null, else by target.
// pass_list/pass_count (optional): targets whose MSV score passes F1 (LenConst::msv_xj_min) are appended.
int launch_ssv(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
               float *d_sc, int32_t *d_status, int64_t *d_noresult_list, unsigned int *d_noresult_count,
               int32_t *d_pass_list, unsigned int *d_pass_count);
int launch_msv_j(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_list,
                 const unsigned int *d_count, int64_t max_n, float *d_sc, int32_t *d_status,
                 int32_t *d_pass_list, unsigned int *d_pass_count);
int launch_vit(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
               float *d_sc, int32_t *d_status);
int launch_fwd(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
               float *d_sc, int32_t *d_status, float *d_xmx, const int64_t *d_xmx_off, float *d_totscale);
int launch_bwd(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
               const float *d_fwd_xmx, const int64_t *d_xmx_off, float *d_sc, int32_t *d_status,
               float *d_bck_xmx, float *d_totscale, int32_t *d_own = nullptr);
// fbx.cu: the two parsers in the reference's own order of float operations (bit-identical specials); same arguments
int launch_fbx_fwd(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
                   float *d_sc, int32_t *d_status, float *d_xmx, const int64_t *d_xmx_off, float *d_totscale);
int launch_fbx_bwd(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
                   const float *d_fwd_xmx, const int64_t *d_xmx_off, float *d_sc, int32_t *d_status,
                   float *d_bck_xmx, float *d_totscale, int32_t *d_own);
// host libm restatement of the two scores from the specials (fwdback.c:434,:463 / :553,:689,:735): totscale and score exactly as the
// reference accumulates them; used whenever exact-order specials reach the host
void fbx_host_fwd_score_compact(const float *blk, int L, float pmove, float *totscale, float *sc, int *status);
void fbx_host_fwd_score(const float *xmx, int L, float pmove, float *totscale, float *sc, int *status);
void fbx_host_bwd_score(const float *xmx, int L, float *totscale, float *sc, int *status);
// hmmscan's MSV stage over a block of models (ssv_msv.cu: ssv_scan_kernel; scan.cu builds the tables): work item = (model, query)
struct ScanModel { uint32_t tab_off; int rowWords, P, npass; MsvScalars s; };      // tab_off: this model's table inside ScanParams::tabs, in words
struct ScanParams {
  const ScanModel *models; const uint32_t *tabs;
  const uint32_t *lc; const int32_t *lenidx; int nlens;      // lc[model * nlens + lenidx[query]] = tjb_b | msv_xj_min << 16
  const uint8_t *res; const int64_t *off; const int32_t *len; long long nq;
  long long nitems;                                          // models in the block x nq, model-major
  int32_t *npossible;                                        // per model: queries that pass MSV or need the J recursion
  unsigned long long *work_counter;
};
int launch_ssv_scan(p7gpu_ctx *ctx, cudaStream_t stream, const ScanParams &p);
// the two per-length MSV constants of a model that is not uploaded (p7gpu_host.cu): tjb_b | msv_xj_min << 16 for each length
void scan_len_words(const p7gpu_oprofile_view *om, const int32_t *lengths, int64_t n, double F1, uint32_t *out);
// decode.cu: the region trigger of p7_domaindef_ByPosteriorHeuristics (max_i mocc[i] >= rt1) and the gather of selected specials
int launch_region_trigger(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
                          const float *d_fx, const float *d_bx, const int64_t *d_xoff, const int32_t *d_own,
                          float *d_mocc_max, int32_t *d_flag, unsigned int *d_nflag, float slack = 0.0f);
int launch_gather_xmx(p7gpu_ctx *ctx, int lane, const float *d_fx, const float *d_bx, const int64_t *d_src, const int64_t *d_dst,
                      const int32_t *d_nfloats, int64_t n, float *d_ofx, float *d_obx);
// decode.cu: p7_DomainDecoding (btot, etot, mocc per row + Forward's SCALE column and xC(L)) for selected targets, 4*(L+1)+2 floats each
int launch_gather_scale(p7gpu_ctx *ctx, int lane, const DbView &db, const int64_t *d_idx, int64_t n, const float *d_fx,
                        const int64_t *d_xoff, float *d_out);
int launch_domain_decoding(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
                           const float *d_fx, const float *d_bx, const int64_t *d_xoff, const int64_t *d_doff, float *d_dec);
int launch_bias(p7gpu_ctx *ctx, int lane, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, int64_t n,
                float *d_filtersc, int32_t *d_unsure_list = nullptr, unsigned int *d_unsure_count = nullptr);
// p7_bg_FilterScore (p7_bg.c:472-484; esl_hmm_Forward, easel/esl_hmm.c:351-409) on the host, host libm: the bias kernel's arithmetic
// for the few targets where it reports that a 1-ulp difference between CUDA's and glibc's log() could show (bias.cu)
float host_bias_filtersc(const BiasHmm &h, const uint8_t *x, int L);
// recompute filtersc on the host for the work items the bias kernel could not be sure of and patch them into d_filtersc;
// h_filtersc (optional) is patched as well.  *nfixed (optional) receives their number.
int fix_unsure_filtersc(p7gpu_ctx *ctx, const p7gpu_profile *prof, const DbView &db, const int64_t *d_idx, float *d_filtersc,
                        const int32_t *d_unsure_list, const unsigned int *d_unsure_count, float *h_filtersc, unsigned *nfixed);
}  // namespace p7g
