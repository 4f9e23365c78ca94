/* impl_b200.h — HMMER's fourth DP engine: the B200 implementation behind impl_sse's operator API.
 *
 * HMMER picks its vector engine at configure time (configure.ac:197-218); impl_sse, impl_vmx and impl_dummy export
 * the same symbols (src/impl_sse/impl_sse.h:258-350).  This directory is what a maintainer adds beside them: it keeps
 * impl_sse's TYPES and host code (P7_OPROFILE, P7_OMX, p7_oprofile_Convert, I/O, p7_Forward/p7_Backward/p7_Decoding for
 * the per-envelope rescoring) and replaces the five per-target DP entry points of the acceleration pipeline
 *
 *     p7_SSVFilter       impl_sse.h:311      p7_MSVFilter      impl_sse.h:309
 *     p7_ViterbiFilter   impl_sse.h:344      p7_ForwardParser  impl_sse.h:325      p7_BackwardParser  impl_sse.h:328
 *
 * and, for targets finished from GPU results, p7_DomainDecoding (impl_sse.h:299; decoding.c:166-200),
 * with calls into libp7gpu.so (include/p7gpu.h), plus the batched worker loop that replaces hmmsearch.c:1399-1411.
 * integration/Makefile compiles the reference's own sources where they lie, with the five impl_sse definitions renamed
 * to p7_*_sse on the compiler command line (nothing is edited or copied), so both engines live in one binary and can be
 * compared call by call (impl_b200_utest.c).
 */
#ifndef P7_IMPL_B200_INCLUDED
#define P7_IMPL_B200_INCLUDED

#include "p7_config.h"
#include "easel.h"
#include "esl_sq.h"
#include "hmmer.h"
#include "p7gpu.h"

/* ---- the reference's own SSE definitions, still linked in under these names (integration/Makefile: -Dp7_X=p7_X_sse) ---- */
extern int p7_SSVFilter_sse     (const ESL_DSQ *dsq, int L, const P7_OPROFILE *om, float *ret_sc);
extern int p7_MSVFilter_sse     (const ESL_DSQ *dsq, int L, const P7_OPROFILE *om, P7_OMX *ox, float *ret_sc);
extern int p7_ViterbiFilter_sse (const ESL_DSQ *dsq, int L, const P7_OPROFILE *om, P7_OMX *ox, float *ret_sc);
extern int p7_ForwardParser_sse (const ESL_DSQ *dsq, int L, const P7_OPROFILE *om, P7_OMX *fwd, float *opt_sc);
extern int p7_BackwardParser_sse(const ESL_DSQ *dsq, int L, const P7_OPROFILE *om, const P7_OMX *fwd, P7_OMX *bck, float *opt_sc);
extern int p7_DomainDecoding_sse(const P7_OPROFILE *om, const P7_OMX *oxf, const P7_OMX *oxb, P7_DOMAINDEF *ddef);

/* ---- engine: one per GPU (the counterpart of one WORKER_INFO, hmmsearch.c:36-44) ---- */
typedef struct p7_b200_s P7_B200;

extern int         p7_b200_DeviceCount(void);
extern int         p7_b200_Open (int device, P7_B200 **ret_eng);         /* impl_Init() of this engine; eslESYS without a GPU */
extern void        p7_b200_Close(P7_B200 *eng);
extern const char *p7_b200_ErrorText(void);                              /* p7gpu_last_error() */
/* hand the query over once, after p7_oprofile_Convert + p7_pli_NewModel (hmmsearch.c:500-512): de-stripes and uploads <om>,
 * and the bias-filter HMM that p7_bg_SetFilter built in <bg> (p7_pipeline.c:505) */
extern int         p7_b200_SetModel(P7_B200 *eng, const P7_OPROFILE *om, const P7_BG *bg);
/* same for a profile of which only the MSV part has been read (p7_oprofile_ReadMSV, impl_sse/io.c:233): the MSV/SSV stage only */
extern int         p7_b200_SetModelMSV(P7_B200 *eng, const P7_OPROFILE *om);
/* scan mode: the targets are few and fixed; per-length constants of the current model for exactly these lengths (F1 = 1.0 under --max) */
extern int         p7_b200_PrepareLengths(P7_B200 *eng, const int32_t *lengths, int64_t n, double F1);
extern p7gpu_ctx     *p7_b200_Ctx    (P7_B200 *eng);
extern p7gpu_profile *p7_b200_Profile(P7_B200 *eng);
extern void        p7_b200_ViewOf(const P7_OPROFILE *om, p7gpu_oprofile_view *v);

/* ---- replay: let the UNMODIFIED p7_Pipeline() finish a target whose five DP results the GPU already produced.  While a
 *      record is armed on the calling thread, the five entry points answer from it (for exactly that dsq/L) instead of
 *      launching anything; p7_Pipeline (p7_pipeline.c:665-901) then runs line for line: P-value tests, bias filter, domain
 *      definition, null2, hit creation. ---- */
typedef struct {
  const ESL_DSQ *dsq; int L;
  int   msv_status, vit_status;            /* eslOK / eslERANGE as the filters returned them            */
  float usc, vfsc, fwdsc, bcksc;
  const float *fwd_xmx, *bck_xmx;          /* (L+1)*p7X_NXCELLS floats each, or NULL when <dec> is given */
  const float *dec;                        /* p7_DomainDecoding done on the device: btot[L+1], etot[L+1], mocc[L+1] (p7gpu_survivor.dec_offset) */
  float fwd_totscale, bck_totscale;
  int   bck_own_scales;
} P7_B200_REPLAY;
extern void p7_b200_ReplayArm(const P7_B200_REPLAY *rec);   /* NULL disarms */

/* ---- the batched worker loop (replaces hmmsearch.c:1295-1306 serial_loop / :1372-1424 pipeline_thread) ---- */
typedef struct {
  int      ngpus;            /* engines to use                                                            */
  int      ncpus;            /* host threads finishing survivors through p7_Pipeline (>= 1)               */
  int      fb_exact;         /* 1: specials in the reference's own summation order (bit-identical hits)    */
  int      decode;           /* 1: p7_DomainDecoding on the device; 16 bytes per row come back instead of 48 */
  double   gpu_seconds;      /* out: wall time inside the GPU pipeline calls                              */
  double   post_seconds;     /* out: wall time of the survivors' host post-processing                     */
  uint64_t nsurvivors;       /* out                                                                        */
} P7_B200_SEARCH;

/* Search every target of the packed database <db> with the model set on engs[0..ngpus-1].  Appends hits to <th> and adds
 * the accounting to <pli> exactly as the reference's workers + p7_tophits_Merge/p7_pipeline_Merge would (hmmsearch.c:538-545). */
extern int p7_b200_SearchDatabase(P7_B200 **engs, P7_B200_SEARCH *cfg, ESL_GETOPTS *go, const ESL_ALPHABET *abc,
                                  P7_OPROFILE *om, P7_BG *bg, P7_PIPELINE *pli, P7_TOPHITS *th, p7gpu_dbfile *db);

#endif /* P7_IMPL_B200_INCLUDED */
