/* p7gpu.h — C ABI of libp7gpu.so: the B200 (sm_100a) engine behind HMMER3's impl_* operator API.
 *
 * Everything here is `extern "C"`, plain pointers and sizes, no torch/CUDA types.  Each entry point
 * names the reference interface it replaces (paths relative to reiverjohn/biobench2, HMMER/src/...).
 * INTEGRATION.md shows the reference-side stub (an `impl_b200/` directory beside impl_sse/) that binds
 * these symbols under HMMER's own names (p7_MSVFilter, p7_ViterbiFilter, ...).
 *
 * Conventions (same as the reference):
 *   - digital sequences are ESL_DSQ: dsq[0] = dsq[L+1] = 255 sentinels, residues dsq[1..L] in 0..Kp-1
 *     (HMMER/easel/easel.h:308-313);
 *   - return codes are easel's: P7GPU_OK 0, P7GPU_EMEM 5, P7GPU_EINVAL 11, P7GPU_ERANGE 16 (= overflow,
 *     score already +inf, "certain hit"), P7GPU_ENORESULT 19 (only from the SSV stage),
 *     P7GPU_ESYS 12 for CUDA runtime failures (text via p7gpu_last_error());
 *   - the caller owns every host buffer; the library retains nothing but the handles it returns;
 *   - there is NO CPU fallback: without a CUDA device every compute entry returns P7GPU_ESYS.
 */
#ifndef P7GPU_H_INCLUDED
#define P7GPU_H_INCLUDED

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P7GPU_OK         0
#define P7GPU_EMEM       5
#define P7GPU_EINVAL    11
#define P7GPU_ESYS      12
#define P7GPU_ERANGE    16
#define P7GPU_ENORESULT 19

#define P7GPU_KP        29   /* amino Kp (HMMER/easel/esl_alphabet.c:276)  */
#define P7GPU_K         20
#define P7GPU_NXCELLS    6   /* E,N,J,B,C,SCALE per row (impl_sse/impl_sse.h:172-173) */

/* ------------------------------------------------------------------------------------------------
 * 1. The profile as the reference hands it over: a read-only *view* of a P7_OPROFILE
 *    (impl_sse/impl_sse.h:76-143).  Pointers address the reference's own striped SSE arrays; nothing
 *    is copied on the reference side.  p7gpu_profile_upload() de-stripes them into the GPU layout.
 * ---------------------------------------------------------------------------------------------- */
typedef struct p7gpu_oprofile_view {
  int32_t M;                /* om->M                                                            */
  int32_t Kp;               /* om->abc->Kp (29 for amino)                                       */
  int32_t K;                /* om->abc->K  (20)                                                 */
  int32_t mode;             /* om->mode (p7_LOCAL=1 / p7_UNILOCAL=3 ...; filters need local)    */
  float   nj;               /* om->nj: expected # of J's (1.0 multihit, 0.0 unihit)             */
  /* MSV/SSV part: 16 x u8 vectors, node k=q+1+z*Q16 in byte z of vector q (p7_oprofile.c:789-792) */
  const uint8_t *rbv;       /* om->rbv[0]: [Kp][Q16][16]                                        */
  const uint8_t *sbv;       /* om->sbv[0]: [Kp][Q16+17][16] (signed bytes); may be NULL         */
  uint8_t tbm_b, tec_b, tjb_b, base_b, bias_b;  /* tjb_b is ignored: recomputed per target L   */
  float   scale_b;
  /* ViterbiFilter part: 8 x i16 vectors (p7_oprofile.c:816-912)                                  */
  const int16_t *rwv;       /* om->rwv[0]: [Kp][Q8][8]                                          */
  const int16_t *twv;       /* om->twv   : [7*Q8 interleaved BM,MM,IM,DM,MD,MI,II][8] then [Q8][8] DD */
  int16_t xw[4][2];         /* om->xw[E,N,J,C][MOVE,LOOP]; N/C/J MOVE recomputed per target L    */
  float   scale_w;
  int16_t base_w, ddbound_w;
  /* Forward/Backward part: 4 x f32 vectors (p7_oprofile.c:919-985)                               */
  const float   *rfv;       /* om->rfv[0]: [Kp][Q4][4]                                          */
  const float   *tfv;       /* om->tfv   : same interleaving as twv                             */
  float   xf[4][2];         /* om->xf; N/C/J entries recomputed per target L                    */
  /* statistics + bias-filter composition copied from the profile                                 */
  float   evparam[6];       /* MMU, MLAMBDA, VMU, VLAMBDA, FTAU, FLAMBDA (hmmer.h:72)           */
  float   compo[20];        /* om->compo (p7_bg_SetFilter input, p7_bg.c:430-453)               */
} p7gpu_oprofile_view;

typedef struct p7gpu_ctx     p7gpu_ctx;      /* one per worker thread / GPU (cf. hmmsearch.c:36-44 WORKER_INFO) */
typedef struct p7gpu_profile p7gpu_profile;  /* GPU-resident repacked profile                                   */
typedef struct p7gpu_seqdb   p7gpu_seqdb;    /* GPU-resident packed digital target database                     */

/* ------------------------------------------------------------------------------------------------
 * 2. Context
 * ---------------------------------------------------------------------------------------------- */
/* replaces impl_Init() (impl_sse/impl_sse.h:356-376) + per-worker setup (hmmsearch.c:505-516) */
int         p7gpu_ctx_create (int device, p7gpu_ctx **ret_ctx);
void        p7gpu_ctx_destroy(p7gpu_ctx *ctx);
const char *p7gpu_last_error (void);                 /* thread-local text of the last failure */
int         p7gpu_device_count(void);
/* run later launches of this ctx on an externally owned cudaStream_t (e.g. torch's current stream) */
int         p7gpu_ctx_set_stream(p7gpu_ctx *ctx, void *cuda_stream);
/* number of kernel launches issued through this ctx so far (bench.py's gpu_launches claim) */
int64_t     p7gpu_ctx_launch_count(const p7gpu_ctx *ctx);
int         p7gpu_ctx_synchronize(p7gpu_ctx *ctx);
/* engine options (no reference counterpart).  "fb_exact"=1: p7gpu_forward_batch/p7gpu_backward_batch run the exact-order kernels
 * (csrc/fbx.cu: bit-identical to the reference); "fb_fast"=1: the one-target wrappers run the fast kernels instead of the
 * exact-order ones; "fbx_groups" (targets per warp: 1, 2, 4, 8; 0 = auto), "fbx_warps" (cap on resident warps per SM): geometry of the exact-order kernels (diagnostics).  "vit_exact_only"=1: p7gpu_vitfilter_* skips the packed 2 x int16 fast
 * path and runs the exact 32-bit kernel on every target (the two must agree bit for bit; tests use this). */
int         p7gpu_ctx_set_option(p7gpu_ctx *ctx, const char *name, int value);
/* how many targets of the last Viterbi launch on the ctx's main stream left the packed kernel's window and were
 * recomputed by the exact kernel (diagnostic) */
int64_t     p7gpu_ctx_last_vit_redo(p7gpu_ctx *ctx);
/* how many bias-filter scores this ctx has recomputed on the host so far: the bias kernel reports the targets for which a
 * 1-ulp difference between CUDA's and glibc's double log() could change the float it is rounded to (about three per 10^7-target
 * pass), and those are redone with the host libm, so p7_bg_FilterScore is the reference's bit for bit by construction
 * (option "bias_window_log2" = n widens the test to 2^-n, for tests) */
int64_t     p7gpu_ctx_bias_recomputed(const p7gpu_ctx *ctx);

/* Diagnostics for bench.py's roofline line (no reference counterpart; cf. the Mc/s drivers msvfilter.c:512-600).
 * With timing on, every kernel launch is bracketed by CUDA events on the launching stream;
 * p7gpu_ctx_kernel_time() synchronises, folds the finished events in and returns the totals since the last
 * p7gpu_ctx_reset_timing() for one kernel kind. */
#define P7GPU_KERNEL_SSV  0
#define P7GPU_KERNEL_MSVJ 1
#define P7GPU_KERNEL_VIT  2
#define P7GPU_KERNEL_FWD  3
#define P7GPU_KERNEL_BWD  4
#define P7GPU_KERNEL_BIAS 5
#define P7GPU_KERNEL_TRIG 6
#define P7GPU_KERNEL_FWDX 7   /* exact-order Forward (csrc/fbx.cu) */
#define P7GPU_KERNEL_BWDX 8   /* exact-order Backward */
#define P7GPU_NKERNELS    9
int         p7gpu_ctx_set_timing  (p7gpu_ctx *ctx, int on);
int         p7gpu_ctx_reset_timing(p7gpu_ctx *ctx);
int         p7gpu_ctx_kernel_time (p7gpu_ctx *ctx, int kind, double *ms_total, int64_t *nlaunches);
/* measured issue peaks of this device for the roofline denominators: packed int16x2 add-max (the SSV/MSV/Viterbi
 * arithmetic), non-fused fp32 mul+add (Forward/Backward) and shared-memory LDS.32 bandwidth. */
int         p7gpu_measure_peaks(p7gpu_ctx *ctx, double *int16x2_gops, double *fp32_gflops, double *lds_gbs);

/* ------------------------------------------------------------------------------------------------
 * 3. Profile: replaces the consumer side of p7_oprofile_Convert()/p7_oprofile_Clone()
 *    (impl_sse/p7_oprofile.c:1004-1041, :403-420).  Per-target p7_oprofile_ReconfigLength()
 *    (:1061-1127) is folded into per-L tables built on the host with the host libm.
 *    Ownership: a profile belongs to the context it was uploaded through and, like the reference's per-worker om clone
 *    (hmmsearch.c:505-516), to one host thread at a time — its per-L table is rebuilt in place when a call brings a longer
 *    target or another F1.  All of its tables live in one device allocation made on the context's stream.
 * ---------------------------------------------------------------------------------------------- */
int  p7gpu_profile_upload (p7gpu_ctx *ctx, const p7gpu_oprofile_view *om, p7gpu_profile **ret_prof);
void p7gpu_profile_destroy(p7gpu_profile *prof);
int  p7gpu_profile_M      (const p7gpu_profile *prof);
/* Optional, for scans of many profiles against few fixed targets (hmmscan: src/hmmscan.c:1180-1222 streams the models past each
 * query): build the profile's per-length constants (what p7_oprofile_ReconfigLength / p7_bg_SetLength / the MSV P-value threshold
 * derive from L, p7_oprofile.c:1085-1127, p7_pipeline.c:690-693) for exactly these target lengths and this F1 (1.0 for --max)
 * instead of for every length up to the longest.  Later calls with a target of another length rebuild the full table. */
int  p7gpu_profile_prepare_lengths(p7gpu_profile *prof, const int32_t *lengths, int64_t n, double F1);
/* the bias-filter HMM as p7_bg_SetFilter() built it (p7_bg.c:430-453): bg->fhmm->t[2][3] (row-major), pi[3], eo[Kp][2].
 * t[0][0], t[0][1] are overwritten per target exactly like p7_bg_SetLength (p7_bg.c:193-196). */
int  p7gpu_profile_set_biasfilter(p7gpu_profile *prof, const float *t6, const float *pi3, const float *eo);

/* ------------------------------------------------------------------------------------------------
 * 4. Target database: replaces ESL_SQ_BLOCK hand-off (easel/esl_sqio.c:1041; hmmsearch.c:1317-1361).
 *    Packed form: `residues` holds every target's dsq[1..L] one after the other, in target order, with AT LEAST ONE
 *    255 sentinel byte before the first target and after every target (any further gap bytes must be 255 too), so
 *    residues + offsets[s] - 1 is a valid sentinel-framed ESL_DSQ of length lengths[s] (offsets[s] = index of the
 *    target's first residue).  Recommended: start every target on a 16-byte boundary — the kernels read residues as
 *    aligned 16-byte vectors, and a target that starts mid-vector pays for up to 15 masked DP rows (4% of the MSV pass
 *    at mean length 350).  `nbytes` counts everything up to and including the last sentinel.
 * ---------------------------------------------------------------------------------------------- */
int  p7gpu_seqdb_upload (p7gpu_ctx *ctx, const uint8_t *residues, int64_t nbytes,
                         const int64_t *offsets, const int32_t *lengths, int64_t nseq,
                         p7gpu_seqdb **ret_db);
/* same, for arrays that already live in this device's HBM (no copy; caller keeps them alive) */
int  p7gpu_seqdb_adopt_device(p7gpu_ctx *ctx, const void *d_residues, int64_t nbytes,
                              const void *d_offsets, const void *d_lengths, int64_t nseq,
                              int32_t max_length, p7gpu_seqdb **ret_db);
void    p7gpu_seqdb_destroy(p7gpu_seqdb *db);
int64_t p7gpu_seqdb_nseq   (const p7gpu_seqdb *db);
int64_t p7gpu_seqdb_nres   (const p7gpu_seqdb *db);

/* ------------------------------------------------------------------------------------------------
 * 5. Batched stages.  `idx` selects targets of `db` (NULL = all nseq targets, n ignored→nseq);
 *    outputs are host arrays of n entries, in idx order.  Scores are nats exactly as the reference
 *    returns them; status is the reference's return code per target.
 * ---------------------------------------------------------------------------------------------- */
/* p7_MSVFilter(dsq,L,om,ox,&sc) (impl_sse/msvfilter.c:75-210): SSV first (ssvfilter.c:877-926), the
 * J-state recursion for targets where SSV answers eslENORESULT.                                     */
int p7gpu_msvfilter_batch(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db,
                          const int64_t *idx, int64_t n, float *sc, int32_t *status);
/* p7_SSVFilter alone (ssvfilter.c:877-926); status may be P7GPU_ENORESULT */
int p7gpu_ssvfilter_batch(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db,
                          const int64_t *idx, int64_t n, float *sc, int32_t *status);
/* p7_ViterbiFilter (impl_sse/vitfilter.c:83-249) */
int p7gpu_vitfilter_batch(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db,
                          const int64_t *idx, int64_t n, float *sc, int32_t *status);
/* p7_ForwardParser (impl_sse/fwdback.c:133-144, 257-465).  If xmx != NULL it receives, for target j,
 * (lengths+1)*6 floats {E,N,J,B,C,SCALE} per row at xmx + xmx_offsets[j] (floats); totscale[j] optional. */
int p7gpu_forward_batch  (p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db,
                          const int64_t *idx, int64_t n, float *sc, int32_t *status,
                          float *xmx, const int64_t *xmx_offsets, float *totscale);
/* p7_BackwardParser (impl_sse/fwdback.c:237-249, 469-735): consumes the forward specials (scale factors) */
int p7gpu_backward_batch (p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db,
                          const int64_t *idx, int64_t n, const float *fwd_xmx, const int64_t *xmx_offsets,
                          float *sc, int32_t *status, float *bck_xmx, float *bck_totscale);

/* The region trigger of the domain-definition step: for each selected target, from its Forward and Backward
 * specials, max_i ddef->mocc[i] as p7_DomainDecoding computes it (impl_sse/decoding.c:166-200) and flag = 1 where
 * p7_domaindef_ByPosteriorHeuristics can open a region (mocc[i] >= rt1 = 0.25 for some i, p7_domaindef.c:109,419), or where
 * the host must decide (Backward used its own scale factors, bck_own_scales[j] != 0; non-finite scale product).  flag = 0
 * means ddef->nregions = 0: p7_Pipeline drops the target (p7_pipeline.c:739).  bck_own_scales may be NULL (all zero). */
int p7gpu_region_trigger_batch(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db,
                               const int64_t *idx, int64_t n, const float *fwd_xmx, const float *bck_xmx,
                               const int64_t *xmx_offsets, const int32_t *bck_own_scales, float *mocc_max, int32_t *flag);
/* p7_bg_FilterScore (p7_bg.c:472-484): the bias-filter null score of the selected targets */
int p7gpu_biasfilter_batch(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db,
                           const int64_t *idx, int64_t n, float *filtersc);

/* device-resident variant of the MSV stage: results stay in HBM (d_sc: nseq floats, d_status: nseq int32, device
 * pointers owned by the caller).  If d_pass_list/d_pass_count are given (nseq int32 / one uint32, device), the targets
 * whose P-value passes F1 (p7_pipeline.c:691-693, decided exactly from a host-built per-length table) are appended
 * there in no particular order.  Asynchronous on the ctx's stream. */
int p7gpu_msvfilter_device(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db, double F1,
                           void *d_sc, void *d_status, void *d_pass_list, void *d_pass_count);

/* The ESL_SQ_BLOCK stream (hmmsearch.c:1317-1361 reader thread -> :1372-1424 workers) as one call: a HOST packed
 * database is copied to the device in chunks while the MSV filter of earlier chunks runs.  Host outputs (any may be
 * NULL): sc/status per target, pass_list (ascending target indices passing F1; room for nseq) and *npass.
 * If opt_ret_db != NULL the uploaded database stays resident and is returned for the later stages. */
int p7gpu_msvfilter_stream(p7gpu_ctx *ctx, const p7gpu_profile *prof, const uint8_t *residues, int64_t nbytes,
                           const int64_t *offsets, const int32_t *lengths, int64_t nseq, double F1,
                           float *sc, int32_t *status, int32_t *pass_list, int64_t *npass, p7gpu_seqdb **opt_ret_db);

/* ------------------------------------------------------------------------------------------------
 * 6. Signature-level wrappers: one target per call, same argument meaning as impl_sse.h:309-344
 *    (P7_OPROFILE* -> p7gpu_profile*, P7_OMX* -> the xmx array it would have filled).
 * ---------------------------------------------------------------------------------------------- */
int p7gpu_SSVFilter     (p7gpu_ctx *ctx, const uint8_t *dsq, int L, const p7gpu_profile *om, float *ret_sc);
int p7gpu_MSVFilter     (p7gpu_ctx *ctx, const uint8_t *dsq, int L, const p7gpu_profile *om, float *ret_sc);
int p7gpu_ViterbiFilter (p7gpu_ctx *ctx, const uint8_t *dsq, int L, const p7gpu_profile *om, float *ret_sc);
int p7gpu_ForwardParser (p7gpu_ctx *ctx, const uint8_t *dsq, int L, const p7gpu_profile *om,
                         float *fwd_xmx /* (L+1)*6 or NULL */, float *fwd_totscale, float *opt_sc);
int p7gpu_BackwardParser(p7gpu_ctx *ctx, const uint8_t *dsq, int L, const p7gpu_profile *om,
                         const float *fwd_xmx, float *bck_xmx /* (L+1)*6 or NULL */,
                         float *bck_totscale, float *opt_sc);
/* same, also reporting P7_OMX.has_own_scales of the Backward matrix (fwdback.c:661-663) */
int p7gpu_BackwardParser2(p7gpu_ctx *ctx, const uint8_t *dsq, int L, const p7gpu_profile *om,
                          const float *fwd_xmx, float *bck_xmx, float *bck_totscale, float *opt_sc, int *opt_own_scales);
/* The one-target wrappers of the two parsers run the exact-order kernels (bit-identical to the reference) unless the ctx
 * option "fb_fast" is set; the batched p7gpu_forward_batch/p7gpu_backward_batch run the fast kernels unless "fb_exact" is set. */

/* ------------------------------------------------------------------------------------------------
 * 7. The batched pipeline: what the loop body at hmmsearch.c:1399-1411 does per target
 *    (p7_pli_NewSeq; p7_bg_SetLength; p7_oprofile_ReconfigLength; p7_Pipeline; Reuse), for a whole
 *    packed database, stage-major (SURVEY.md §8b "stage-major reordering is legal").
 *    It stops where north_star stops: after p7_BackwardParser (p7_pipeline.c:735).  Survivors come
 *    back as records carrying everything p7_domaindef_ByPosteriorHeuristics()+hit creation
 *    (p7_pipeline.c:737-898, unchanged host code) need.
 * ---------------------------------------------------------------------------------------------- */
typedef struct p7gpu_pipeline_opts {   /* the fields of P7_PIPELINE the path reads (hmmer.h:1175-1250) */
  double  F1, F2, F3;                  /* p7_pipeline.c:214-216 defaults 0.02, 1e-3, 1e-5            */
  int32_t do_biasfilter;               /* pli->do_biasfilter                                          */
  int32_t do_max;                      /* --max: F1=F2=F3=1, bias off (p7_pipeline.c:222-229)         */
  int32_t want_specials;               /* 1: return fwd/bck xmx for Forward survivors                  */
  int32_t stop_after_msv;              /* 1: MSV stage only (BASELINE config 2): counters + the MSV survivors */
  int32_t region_trigger;              /* 1: evaluate the region trigger of p7_domaindef_ByPosteriorHeuristics
                                        *    (p7_DomainDecoding, impl_sse/decoding.c:166-200; mocc[i] >= rt1, p7_domaindef.c:419)
                                        *    on the device and return specials only for survivors on which it fires: the
                                        *    others have nregions = 0 and p7_Pipeline drops them (p7_pipeline.c:739).
                                        *    Meant for --max, where every target reaches this point.   */
  int32_t fb_exact;                    /* 1: every Forward/Backward result that reaches the host — fwdsc, bcksc, the specials, the F3
                                        *    decision — is computed in the reference's own order of float operations (the striped
                                        *    SSE order of impl_sse/fwdback.c:257-465, :469-735; csrc/fbx.cu) and is therefore
                                        *    bit-identical to p7_ForwardParser/p7_BackwardParser; the fast kernels (1e-4 relative)
                                        *    then only pre-filter for the F3 test.  0: fast kernels only.                    */
  int32_t decode;                      /* 1 (with want_specials and region_trigger): run p7_DomainDecoding (impl_sse/decoding.c:
                                        *    166-200) on the device for the flagged survivors and return btot/etot/mocc per row —
                                        *    all p7_domaindef_ByPosteriorHeuristics reads of the two parser matrices — plus Forward's
                                        *    SCALE column instead of both sets of specials: 16 bytes per row instead of 48.
                                        *    Survivors the host must decide itself (Backward own scales, non-finite scale product)
                                        *    still get their specials.                                                          */
  int32_t stage_records;               /* 1: also return one p7gpu_msv_survivor per target that passed the MSV filter, with the
                                        *    stage it reached (p7gpu_result_msvpass): what a scan pipeline needs to keep the
                                        *    per-query accounting of hmmscan (p7_pipeline.c:694,705,723,731).  Stage decisions are
                                        *    then evaluated on the host.                                                        */
} p7gpu_pipeline_opts;

typedef struct p7gpu_pipeline_counters { /* P7_PIPELINE accounting (hmmer.h:1224-1231), merged like p7_pipeline_Merge */
  uint64_t nseqs, nres, n_past_msv, n_past_bias, n_past_vit, n_past_fwd;
} p7gpu_pipeline_counters;

typedef struct p7gpu_survivor {        /* one per target that passed the Forward filter (p7_pipeline.c:731) */
  int64_t seqidx;                      /* index into the database                                     */
  int32_t L;
  int32_t msv_status, vit_status;      /* reference return codes of the two filters                   */
  float   usc, vfsc, fwdsc, bcksc;     /* nats; vfsc = NaN when the Viterbi stage was skipped (:716)  */
  float   nullsc, filtersc;            /* p7_bg_NullOne / p7_bg_FilterScore                            */
  float   fwd_totscale, bck_totscale;
  int32_t bck_own_scales;              /* P7_OMX.has_own_scales of the Backward matrix (fwdback.c:661-663)        */
  int64_t xmx_offset;                  /* float offset of this target's (L+1)*6 block in fwd_xmx/bck_xmx, or -1 */
  int32_t region_flag;                 /* -1 not evaluated; 0: mocc[i] < rt1 everywhere, no region (nregions = 0);
                                        *  1: the trigger fires somewhere (or the host must decide): specials returned */
  float   mocc_max;                    /* max_i ddef->mocc[i] (NaN when not evaluated)                            */
  int64_t dec_offset;                  /* opts.decode: float offset of this target's block in p7gpu_result_decoding(), or -1.
                                        *  Block: btot[L+1], etot[L+1], mocc[L+1] (P7_DOMAINDEF arrays, p7_domaindef.c:392-407),
                                        *  fwd SCALE[L+1], fwd xC(L), 0                                                   */
} p7gpu_survivor;

typedef struct p7gpu_msv_survivor {    /* stop_after_msv: one per target with P(MSV) <= F1 (p7_pipeline.c:693-694) */
  int64_t seqidx; int32_t L; int32_t msv_status; float usc;
  int32_t stage;                       /* 1 passed MSV, 2 passed the bias filter, 3 passed Viterbi, 4 passed Forward (1 when stop_after_msv) */
} p7gpu_msv_survivor;

typedef struct p7gpu_pipeline_result p7gpu_pipeline_result;

int  p7gpu_pipeline_run(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_seqdb *db,
                        const p7gpu_pipeline_opts *opts, p7gpu_pipeline_result **ret_res);
/* many profiles against one resident database (BASELINE config 5, hmmscan-style batch): the query loop of
 * hmmsearch.c:471-560 without the per-query rewind and re-read of the sequence file (:480-487).  ret_res: nprof slots. */
int  p7gpu_pipeline_run_many(p7gpu_ctx *ctx, const p7gpu_profile *const *profs, int64_t nprof, const p7gpu_seqdb *db,
                             const p7gpu_pipeline_opts *opts, p7gpu_pipeline_result **ret_res);
/* hmmscan's MSV stage for a block of models at once.  hmmscan streams each model past the queries and reads
 * the rest of a model only when some query gets past its MSV filter (hmmscan.c:1176-1198, p7_pipeline.c:688-713); with a handful
 * of queries almost every model fails, and one upload + launch + read-back per model is what the time goes into.  This call
 * takes the MSV parts of `nmodels` profiles as p7_oprofile_ReadMSV delivers them (impl_sse/io.c:233; only the MSV/SSV fields and
 * evparam of each view are read), runs ONE kernel whose work item is a (model, query) pair with the gain table picked per item,
 * and reports for each model how many queries either pass the MSV threshold F1 (p7_pipeline.c:691-693) by their SSV score or
 * need the J-state recursion (eslENORESULT, msvfilter.c:105-106).  ret_npossible[m] == 0 means that
 * p7gpu_pipeline_run(stop_after_msv) on model m would return no survivor: the model need not be uploaded at all.
 * `queries` is the resident packed query database, `lengths` its nseq target lengths (host). */
int  p7gpu_scan_msv_prefilter(p7gpu_ctx *ctx, const p7gpu_oprofile_view *oms, int64_t nmodels, const p7gpu_seqdb *queries,
                              const int32_t *lengths, double F1, int32_t *ret_npossible);
/* diagnostic (host code, no device needed): the scan table p7gpu_scan_msv_prefilter builds for one model and the per-length
 * word (tjb_b | msv_xj_min << 16) it hands the kernel for target length L.  geom[0..2] = columns P (256 * passes >= M + 1),
 * words per row (P + 96), passes; `table` (may be NULL to ask for the geometry only) receives 29 rows of geom[1] words, word w =
 * the int16 SSV gains of columns (w mod P, (w + 128) mod P), pad columns -20000.  tests/test_cabi.py walks it the way the
 * kernel does and compares the outcome with the CPU oracle's p7_SSVFilter. */
int  p7gpu_scan_table_build(const p7gpu_oprofile_view *om, int32_t geom[3], uint32_t *table, int32_t L, double F1, uint32_t *len_word);
/* same for a HOST packed database: the upload is streamed under the MSV stage (p7gpu_msvfilter_stream) */
int  p7gpu_pipeline_run_host(p7gpu_ctx *ctx, const p7gpu_profile *prof, const uint8_t *residues, int64_t nbytes,
                             const int64_t *offsets, const int32_t *lengths, int64_t nseq,
                             const p7gpu_pipeline_opts *opts, p7gpu_pipeline_result **ret_res);
/* diagnostic (host code, no device needed): the float threshold behind the device-side stage decisions.  Every P-value test
 * of p7_Pipeline reads "P(seq_score) > F" with P non-increasing; *T is the smallest float seq_score that passes, found and
 * verified with the host libm in the reference's expressions (kind 0: esl_gumbel_surv(x, mu, lambda), easel/esl_gumbel.c:
 * 129-138; kind 1: esl_exp_surv, easel/esl_exponential.c:127-132).  P7GPU_ERANGE: no verified threshold (host evaluation). */
int  p7gpu_stage_threshold(int kind, double F, float mu, float lambda, float *T);
/* same with the residues in the 5-bit transfer format (eight codes in five bytes; 31 = sentinel): 5/8 of the bytes cross
 * PCIe and a small kernel on the copy stream expands them in HBM.  Where several GPUs share one PCIe uplink (HGX boards)
 * the streamed search is transfer-bound, and this is what brings it back to the kernels' rate.  p7gpu_pack5 builds the
 * format once per database (nthreads = 0: all host threads); `nbytes` is always the size of the UNPACKED residue array. */
int64_t p7gpu_pack5_bytes(int64_t nbytes);
int  p7gpu_pack5(const uint8_t *residues, int64_t nbytes, uint8_t *out, int nthreads);
int  p7gpu_pipeline_run_host5(p7gpu_ctx *ctx, const p7gpu_profile *prof, const uint8_t *packed5, int64_t nbytes,
                              const int64_t *offsets, const int32_t *lengths, int64_t nseq,
                              const p7gpu_pipeline_opts *opts, p7gpu_pipeline_result **ret_res);
/* residues seen by each stage {MSV, bias, Viterbi, Forward, Backward, exact-order Forward recomputation (fb_exact under the default
 * thresholds; 0 otherwise)}: the denominators of per-stage cells/s
 * (the reference reports one Mc/s figure for the whole pipeline, p7_pipeline.c:1789-1793) */
const uint64_t                *p7gpu_result_stage_residues(const p7gpu_pipeline_result *res);
int64_t                        p7gpu_result_nmsvpass  (const p7gpu_pipeline_result *res);
const p7gpu_msv_survivor      *p7gpu_result_msvpass   (const p7gpu_pipeline_result *res);
const p7gpu_pipeline_counters *p7gpu_result_counters (const p7gpu_pipeline_result *res);
int64_t                        p7gpu_result_nsurvivors(const p7gpu_pipeline_result *res);
const p7gpu_survivor          *p7gpu_result_survivors (const p7gpu_pipeline_result *res);
const float                   *p7gpu_result_fwd_xmx  (const p7gpu_pipeline_result *res);
const float                   *p7gpu_result_bck_xmx  (const p7gpu_pipeline_result *res);
const float                   *p7gpu_result_decoding (const p7gpu_pipeline_result *res);   /* opts.decode blocks, see p7gpu_survivor.dec_offset */
void                           p7gpu_result_destroy  (p7gpu_pipeline_result *res);   /* before p7gpu_ctx_destroy of its context:
                                                                                      a result hands its largest buffer back to the context */

/* host-side helpers restating the reference's per-target scalar arithmetic (host libm, bit-exact):
 * p7_bg_NullOne (p7_bg.c:358-363) and the MSV P-value test of p7_pipeline.c:691-693. */
float  p7gpu_nullsc(int L);
double p7gpu_msv_pvalue(float usc, float nullsc_or_filtersc, float mmu, float mlambda);

/* ------------------------------------------------------------------------------------------------
 * 8. The packed on-disk target database.  Replaces, on the search path, esl_sqfile_OpenDigital +
 *    esl_sqio_ReadBlock (easel/esl_sqio_ascii.c:1466-1500: one reader thread, text parsed again for every query);
 *    the reference's own precedents for a pre-digitised database are src/cachedb.h and src/makehmmerdb.c.
 *    p7gpu_dbfile_from_fasta digitises a FASTA file once, with `nthreads` host threads (0 = all), exactly as easel does
 *    (amino alphabet esl_alphabet.c:269-296; FASTA input map and header rules esl_sqio_ascii.c:2952-3056; the same
 *    error texts in p7gpu_last_error()), into: residues[nbytes] (dsq bytes back to back, one 255 sentinel before the
 *    first and after every target; with align = 16 — recommended — every target starts on a 16-byte boundary, the gap
 *    filled with sentinels, so that the kernels' aligned 16-byte residue loads carry no masked leading rows; align = 1:
 *    exactly one sentinel), offsets[nseq] int64, lengths[nseq] int32, names and descriptions.  A search maps the
 *    file and streams it: p7gpu_pipeline_run_file == p7gpu_pipeline_run_host on the mapped arrays (after
 *    p7gpu_dbfile_pin: == p7gpu_pipeline_run_host5 on the pinned 5-bit copy).
 * ---------------------------------------------------------------------------------------------- */
typedef struct p7gpu_dbfile p7gpu_dbfile;
int            p7gpu_dbfile_from_fasta(const char *fasta_path, const char *out_path, int nthreads, int align, int64_t *ret_nseq, int64_t *ret_nres);
int            p7gpu_dbfile_open      (const char *path, p7gpu_dbfile **ret_db);
void           p7gpu_dbfile_close     (p7gpu_dbfile *db);
int64_t        p7gpu_dbfile_nseq      (const p7gpu_dbfile *db);
int64_t        p7gpu_dbfile_nres      (const p7gpu_dbfile *db);
int64_t        p7gpu_dbfile_nbytes    (const p7gpu_dbfile *db);
int32_t        p7gpu_dbfile_maxlength (const p7gpu_dbfile *db);
const uint8_t *p7gpu_dbfile_residues  (const p7gpu_dbfile *db);
const int64_t *p7gpu_dbfile_offsets   (const p7gpu_dbfile *db);
const int32_t *p7gpu_dbfile_lengths   (const p7gpu_dbfile *db);
const char    *p7gpu_dbfile_name      (const p7gpu_dbfile *db, int64_t s);     /* ESL_SQ.name */
const char    *p7gpu_dbfile_desc      (const p7gpu_dbfile *db, int64_t s);     /* ESL_SQ.desc */
int            p7gpu_dbfile_pin       (p7gpu_dbfile *db);                      /* load into page-locked memory (full-rate H2D),
                                                                                  also in the 5-bit transfer format, which
                                                                                  p7gpu_pipeline_run_file then streams; the
                                                                                  array accessors return the pinned copy */
int            p7gpu_pipeline_run_file(p7gpu_ctx *ctx, const p7gpu_profile *prof, const p7gpu_dbfile *db,
                                       const p7gpu_pipeline_opts *opts, p7gpu_pipeline_result **ret_res);

/* ------------------------------------------------------------------------------------------------
 * 9. Several GPUs from one C host (replaces hmmsearch's worker threads hmmsearch.c:505-516 and --mpi, :806-1283).
 *    The database shards by sequence with no data-path exchange: contiguous ranges balanced by residue count, one
 *    context (device) and one host thread per shard.  first[nshards+1] receives the shard boundaries; ret_res[k] is
 *    shard k's result with seqidx already in DATABASE numbering.  The caller merges exactly as the reference merges
 *    its workers: counters by sum (p7gpu_counters_merge == p7_pipeline_Merge, p7_pipeline.c:597-636), hits through
 *    p7_tophits_Merge (integration/impl_b200/p7_b200_search.c).  profs[k] must have been uploaded through ctxs[k].
 * ---------------------------------------------------------------------------------------------- */
int  p7gpu_shard_plan(const int64_t *offsets, const int32_t *lengths, int64_t nseq, int nshards, int64_t *first);
int  p7gpu_pipeline_run_sharded(p7gpu_ctx *const *ctxs, const p7gpu_profile *const *profs, int nshards,
                                const uint8_t *residues, const uint8_t *packed5, int64_t nbytes,
                                const int64_t *offsets, const int32_t *lengths, int64_t nseq,
                                const p7gpu_pipeline_opts *opts, int64_t *first, p7gpu_pipeline_result **ret_res);
int  p7gpu_pipeline_run_file_sharded(p7gpu_ctx *const *ctxs, const p7gpu_profile *const *profs, int nshards,
                                     const p7gpu_dbfile *db, const p7gpu_pipeline_opts *opts, int64_t *first,
                                     p7gpu_pipeline_result **ret_res);
void p7gpu_counters_merge(p7gpu_pipeline_counters *into, const p7gpu_pipeline_counters *from);

#ifdef __cplusplus
}
#endif
#endif /* P7GPU_H_INCLUDED */
