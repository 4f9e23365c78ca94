/* p7_b200_engine.c — the five per-target DP entry points of HMMER's acceleration pipeline, bound to libp7gpu.so.
 *
 * Exactly the signatures of src/impl_sse/impl_sse.h:309,311,325,328,344.  A P7_OPROFILE reaches its GPU-resident
 * counterpart through a small registry keyed by the profile's address and a fingerprint of its contents (profiles are
 * cloned, reconfigured and freed by host code this engine never sees): the first call for an unknown profile uploads it.
 * Per-target length configuration (p7_oprofile_ReconfigLength, p7_oprofile.c:1061-1127) is derived from the L argument
 * inside the library, never read from om->L.
 *
 * While a replay record is armed on the calling thread (p7_b200_ReplayArm) the five functions answer from it: that is how
 * the batched search hands the GPU's per-target results to the reference's unmodified p7_Pipeline().
 */
#include "impl_b200.h"
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct p7_b200_s {
  int              device;
  p7gpu_ctx       *ctx;
  p7gpu_profile   *prof;            /* current query */
  uint64_t         finger;          /* fingerprint of the P7_OPROFILE <prof> was built from */
  const void      *om_addr;
  pthread_mutex_t  lock;            /* a p7gpu_ctx serves one caller at a time */
  int              owned_by_registry;
};

static __thread const P7_B200_REPLAY *tl_replay = NULL;

void p7_b200_ReplayArm(const P7_B200_REPLAY *rec) { tl_replay = rec; }

int         p7_b200_DeviceCount(void) { return p7gpu_device_count(); }
const char *p7_b200_ErrorText(void)   { return p7gpu_last_error(); }
p7gpu_ctx     *p7_b200_Ctx    (P7_B200 *e) { return e ? e->ctx  : NULL; }
p7gpu_profile *p7_b200_Profile(P7_B200 *e) { return e ? e->prof : NULL; }

/* the read-only view of a P7_OPROFILE (impl_sse.h:76-143) that p7gpu_profile_upload de-stripes; nothing is copied here */
void
p7_b200_ViewOf(const P7_OPROFILE *om, p7gpu_oprofile_view *v)
{
  int s, t;
  memset(v, 0, sizeof(*v));
  v->M = om->M;  v->Kp = om->abc->Kp;  v->K = om->abc->K;  v->mode = om->mode;  v->nj = om->nj;
  v->rbv = (const uint8_t *) om->rbv[0];  v->sbv = (const uint8_t *) om->sbv[0];
  v->tbm_b = om->tbm_b; v->tec_b = om->tec_b; v->tjb_b = om->tjb_b; v->base_b = om->base_b; v->bias_b = om->bias_b;
  v->scale_b = om->scale_b;
  v->rwv = (const int16_t *) om->rwv[0];  v->twv = (const int16_t *) om->twv;
  v->rfv = (const float *)   om->rfv[0];  v->tfv = (const float *)   om->tfv;
  for (s = 0; s < p7O_NXSTATES; s++)
    for (t = 0; t < p7O_NXTRANS; t++) { v->xw[s][t] = om->xw[s][t];  v->xf[s][t] = om->xf[s][t]; }
  v->scale_w = om->scale_w;  v->base_w = om->base_w;  v->ddbound_w = om->ddbound_w;
  for (s = 0; s < p7_NEVPARAM; s++) v->evparam[s] = om->evparam[s];
  for (s = 0; s < p7_MAXABET;  s++) v->compo[s]   = om->compo[s];
}

/* FNV-1a over the L-independent scalars and the edges of the three striped tables */
static uint64_t
fingerprint(const P7_OPROFILE *om)
{
  uint64_t h = 1469598103934665603ULL;
  const unsigned char *p; size_t n, i;
#define MIX(ptr, len) do { p = (const unsigned char *) (ptr); n = (len); for (i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ULL; } } while (0)
  const int Q16 = p7O_NQB(om->M), Q8 = p7O_NQW(om->M), Q4 = p7O_NQF(om->M);
  MIX(&om->M, sizeof(int));  MIX(&om->mode, sizeof(int));  MIX(&om->nj, sizeof(float));
  MIX(&om->tbm_b, 1); MIX(&om->tec_b, 1); MIX(&om->base_b, 1); MIX(&om->bias_b, 1); MIX(&om->scale_b, sizeof(float));
  MIX(&om->scale_w, sizeof(float)); MIX(&om->base_w, sizeof(int16_t)); MIX(&om->ddbound_w, sizeof(int16_t));
  MIX(om->xw[p7O_E], sizeof(int16_t) * p7O_NXTRANS);  MIX(om->xf[p7O_E], sizeof(float) * p7O_NXTRANS);
  MIX(om->evparam, sizeof(float) * p7_NEVPARAM);
  MIX(om->rbv[0], 16 * (size_t) ESL_MIN(Q16, 8));  MIX(om->rbv[om->abc->K - 1], 16 * (size_t) ESL_MIN(Q16, 8));
  MIX(om->rwv[0], 16 * (size_t) ESL_MIN(Q8, 8));   MIX(om->twv, 16 * (size_t) ESL_MIN(8 * Q8, 16));
  MIX(om->rfv[0], 16 * (size_t) ESL_MIN(Q4, 8));   MIX(om->tfv, 16 * (size_t) ESL_MIN(8 * Q4, 16));
#undef MIX
  return h;
}

int
p7_b200_Open(int device, P7_B200 **ret_eng)
{
  P7_B200 *e = calloc(1, sizeof(P7_B200));
  if (!e) return eslEMEM;
  e->device = device;
  pthread_mutex_init(&e->lock, NULL);
  if (p7gpu_ctx_create(device, &e->ctx) != P7GPU_OK) { free(e); *ret_eng = NULL; return eslESYS; }
  *ret_eng = e;
  return eslOK;
}

void
p7_b200_Close(P7_B200 *e)
{
  if (!e) return;
  if (e->prof) p7gpu_profile_destroy(e->prof);
  if (e->ctx)  p7gpu_ctx_destroy(e->ctx);
  pthread_mutex_destroy(&e->lock);
  free(e);
}

static int set_model(P7_B200 *e, const P7_OPROFILE *om, const P7_BG *bg, int msv_only);

int p7_b200_SetModel   (P7_B200 *e, const P7_OPROFILE *om, const P7_BG *bg) { return set_model(e, om, bg, FALSE); }
int p7_b200_SetModelMSV(P7_B200 *e, const P7_OPROFILE *om)                  { return set_model(e, om, NULL, TRUE); }

int
p7_b200_PrepareLengths(P7_B200 *e, const int32_t *lengths, int64_t n, double F1)
{
  int st;
  if (!e || !e->prof) return eslEINVAL;
  pthread_mutex_lock(&e->lock);
  st = p7gpu_profile_prepare_lengths(e->prof, lengths, n, F1);
  pthread_mutex_unlock(&e->lock);
  return st == P7GPU_OK ? eslOK : eslESYS;
}

static int
set_model(P7_B200 *e, const P7_OPROFILE *om, const P7_BG *bg, int msv_only)
{
  p7gpu_oprofile_view v;
  int st;
  pthread_mutex_lock(&e->lock);
  if (e->prof) { p7gpu_profile_destroy(e->prof); e->prof = NULL; }
  p7_b200_ViewOf(om, &v);
  if (msv_only) { v.rwv = NULL; v.twv = NULL; v.rfv = NULL; v.tfv = NULL; }     /* not read yet: p7gpu_profile_upload skips those parts */
  st = p7gpu_profile_upload(e->ctx, &v, &e->prof);
  if (st == P7GPU_OK && bg && bg->fhmm) {
    float t6[6], pi3[3], eo[P7GPU_KP * 2];
    int k, x;
    for (k = 0; k < 2; k++) for (x = 0; x < 3; x++) t6[k*3+x] = bg->fhmm->t[k][x];
    for (x = 0; x < 3; x++) pi3[x] = bg->fhmm->pi[x];
    for (x = 0; x < om->abc->Kp && x < P7GPU_KP; x++) for (k = 0; k < 2; k++) eo[x*2+k] = bg->fhmm->eo[x][k];
    st = p7gpu_profile_set_biasfilter(e->prof, t6, pi3, eo);
  }
  e->finger  = msv_only ? 0 : fingerprint(om);
  e->om_addr = om;
  pthread_mutex_unlock(&e->lock);
  return (st == P7GPU_OK) ? eslOK : eslESYS;
}

/* ---- registry for per-target callers (evalues.c:264,337,442 calibration, hmmsim, unit-test drivers): profile -> engine ---- */
#define B200_NSLOTS 8
static P7_B200        *g_slot[B200_NSLOTS];
static pthread_mutex_t g_reg_lock = PTHREAD_MUTEX_INITIALIZER;
static unsigned        g_next;

static void
registry_atexit(void)
{
  int i;
  for (i = 0; i < B200_NSLOTS; i++) if (g_slot[i]) { p7_b200_Close(g_slot[i]); g_slot[i] = NULL; }
}

static P7_B200 *
engine_for(const P7_OPROFILE *om)
{
  const uint64_t f = fingerprint(om);
  P7_B200 *e = NULL;
  int i;
  pthread_mutex_lock(&g_reg_lock);
  for (i = 0; i < B200_NSLOTS; i++)
    if (g_slot[i] && g_slot[i]->finger == f && g_slot[i]->prof) { e = g_slot[i]; break; }
  if (!e) {
    static int registered = 0;
    const char *dev = getenv("P7_B200_DEVICE");
    i = (int) (g_next++ % B200_NSLOTS);
    if (!g_slot[i] && p7_b200_Open(dev ? atoi(dev) : 0, &g_slot[i]) != eslOK) { pthread_mutex_unlock(&g_reg_lock); return NULL; }
    if (!registered) { atexit(registry_atexit); registered = 1; }
    e = g_slot[i];
    if (p7_b200_SetModel(e, om, NULL) != eslOK) e = NULL;
  }
  pthread_mutex_unlock(&g_reg_lock);
  return e;
}

#define B200_ENGINE(e, om) \
  P7_B200 *e = engine_for(om); \
  if (!e) ESL_EXCEPTION(eslESYS, "impl_b200: no GPU engine for this profile: %s", p7gpu_last_error())

static int
gpu_status(int st)
{
  if (st == P7GPU_ESYS) ESL_EXCEPTION(eslESYS, "impl_b200: %s", p7gpu_last_error());
  return st;                  /* eslOK, eslERANGE, eslENORESULT, eslEINVAL, eslEMEM are easel's own codes */
}

/* ============================== the five drop-in symbols ============================== */

/* impl_sse.h:311; ssvfilter.c:877-926.  May return eslENORESULT (the caller then runs p7_MSVFilter's J loop). */
int
p7_SSVFilter(const ESL_DSQ *dsq, int L, const P7_OPROFILE *om, float *ret_sc)
{
  int st;
  B200_ENGINE(e, om);
  pthread_mutex_lock(&e->lock);
  st = p7gpu_SSVFilter(e->ctx, dsq, L, e->prof, ret_sc);
  pthread_mutex_unlock(&e->lock);
  return gpu_status(st);
}

/* impl_sse.h:309; msvfilter.c:75-210.  eslOK, or eslERANGE with *ret_sc = +infinity on overflow. */
int
p7_MSVFilter(const ESL_DSQ *dsq, int L, const P7_OPROFILE *om, P7_OMX *ox, float *ret_sc)
{
  const P7_B200_REPLAY *r = tl_replay;
  int st;
  if (ox) ox->M = om->M;                                 /* msvfilter.c:104 */
  if (r && r->dsq == dsq && r->L == L) { *ret_sc = r->usc; return r->msv_status; }
  {
    B200_ENGINE(e, om);
    pthread_mutex_lock(&e->lock);
    st = p7gpu_MSVFilter(e->ctx, dsq, L, e->prof, ret_sc);
    pthread_mutex_unlock(&e->lock);
  }
  return gpu_status(st);
}

/* impl_sse.h:344; vitfilter.c:83-249 */
int
p7_ViterbiFilter(const ESL_DSQ *dsq, int L, const P7_OPROFILE *om, P7_OMX *ox, float *ret_sc)
{
  const P7_B200_REPLAY *r = tl_replay;
  int st;
  if (ox) ox->M = om->M;                                 /* vitfilter.c:113 */
  if (r && r->dsq == dsq && r->L == L && r->vit_status >= 0) { *ret_sc = r->vfsc; return r->vit_status; }
  {
    B200_ENGINE(e, om);
    pthread_mutex_lock(&e->lock);
    st = p7gpu_ViterbiFilter(e->ctx, dsq, L, e->prof, ret_sc);
    pthread_mutex_unlock(&e->lock);
  }
  return gpu_status(st);
}

/* impl_sse.h:325; fwdback.c:133-144.  The caller has grown <fwd> for L rows of specials (p7_pipeline.c:684). */
int
p7_ForwardParser(const ESL_DSQ *dsq, int L, const P7_OPROFILE *om, P7_OMX *fwd, float *opt_sc)
{
  const P7_B200_REPLAY *r = tl_replay;
  int st; float sc = 0.0f;
  if (fwd->allocXR < L + 1) ESL_EXCEPTION(eslEINVAL, "impl_b200: forward matrix has too few rows of specials");
  fwd->M = om->M;  fwd->L = L;  fwd->has_own_scales = TRUE;          /* fwdback.c:283-285 */
  if (r && r->dsq == dsq && r->L == L && (r->fwd_xmx || r->dec)) {
    if (r->fwd_xmx) memcpy(fwd->xmx, r->fwd_xmx, sizeof(float) * (size_t) (L + 1) * p7X_NXCELLS);   /* else: only p7_DomainDecoding would read them */
    fwd->totscale = r->fwd_totscale;
    if (opt_sc) *opt_sc = r->fwdsc;
    return eslOK;
  }
  {
    B200_ENGINE(e, om);
    pthread_mutex_lock(&e->lock);
    st = p7gpu_ForwardParser(e->ctx, dsq, L, e->prof, fwd->xmx, &fwd->totscale, &sc);
    pthread_mutex_unlock(&e->lock);
  }
  if (opt_sc) *opt_sc = sc;
  if (st == P7GPU_ERANGE) ESL_EXCEPTION(eslERANGE, "forward score out of range");      /* fwdback.c:456-458 */
  return gpu_status(st);
}

/* impl_sse.h:328; fwdback.c:237-249.  Uses <fwd>'s scale factors (fwdback.c:600-612). */
int
p7_BackwardParser(const ESL_DSQ *dsq, int L, const P7_OPROFILE *om, const P7_OMX *fwd, P7_OMX *bck, float *opt_sc)
{
  const P7_B200_REPLAY *r = tl_replay;
  int st; float sc = 0.0f; int own = 0;
  if (bck->allocXR < L + 1) ESL_EXCEPTION(eslEINVAL, "impl_b200: backward matrix has too few rows of specials");
  bck->M = om->M;  bck->L = L;
  if (r && r->dsq == dsq && r->L == L && (r->bck_xmx || r->dec)) {
    if (r->bck_xmx) memcpy(bck->xmx, r->bck_xmx, sizeof(float) * (size_t) (L + 1) * p7X_NXCELLS);
    bck->totscale = r->bck_totscale;  bck->has_own_scales = r->bck_own_scales;         /* fwdback.c:661-663 */
    if (opt_sc) *opt_sc = r->bcksc;
    return eslOK;
  }
  {
    B200_ENGINE(e, om);
    pthread_mutex_lock(&e->lock);
    st = p7gpu_BackwardParser2(e->ctx, dsq, L, e->prof, fwd->xmx, bck->xmx, &bck->totscale, &sc, &own);
    pthread_mutex_unlock(&e->lock);
  }
  bck->has_own_scales = own;
  if (opt_sc) *opt_sc = sc;
  if (st == P7GPU_ERANGE) ESL_EXCEPTION(eslERANGE, "backward score out of range");     /* fwdback.c:731-733 */
  return gpu_status(st);
}

/* impl_sse.h:299; decoding.c:166-200.  The only reader of the two parser matrices after the Backward parser
 * (p7_domaindef_ByPosteriorHeuristics, p7_domaindef.c:392-407).  When the replay record carries the arrays the device computed
 * — in the reference's own order of float operations (csrc/decode.cu) — they are handed over; otherwise the reference's code runs. */
int
p7_DomainDecoding(const P7_OPROFILE *om, const P7_OMX *oxf, const P7_OMX *oxb, P7_DOMAINDEF *ddef)
{
  const P7_B200_REPLAY *r = tl_replay;
  if (r && r->dec && oxf->L == r->L) {
    const size_t n = (size_t) r->L + 1;
    memcpy(ddef->btot, r->dec,         sizeof(float) * n);
    memcpy(ddef->etot, r->dec + n,     sizeof(float) * n);
    memcpy(ddef->mocc, r->dec + 2 * n, sizeof(float) * n);
    ddef->L = oxf->L;
    return eslOK;                        /* non-finite scale products never take this path: the host gets their specials */
  }
  return p7_DomainDecoding_sse(om, oxf, oxb, ddef);
}
