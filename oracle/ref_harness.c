/* ref_harness.c — TEST INFRASTRUCTURE, not product code.
 *
 * A thin ctypes-friendly layer over the UNMODIFIED reference (HMMER 3.1b2 under
 * /root/reference/HMMER), linked by oracle/Makefile into oracle/_ref/libp7ref.so.  It exists to
 *   (1) generate golden vectors (tests/golden/make_golden.py),
 *   (2) pin the C restatement in oracle/p7_oracle.c against the real thing,
 *   (3) serve as the CPU baseline (`cpu_baseline.kind = "reference"`) in bench.py,
 *   (4) finish the pipeline after the Backward parser with the reference's own host code
 *       (p7_domaindef_ByPosteriorHeuristics + hit creation) for hit-list parity tests.
 * Nothing under biobench2_b200/ may link or call this file.
 *
 * All the code below is ours; it only CALLS reference functions through the reference's headers.
 */
#include "p7_config.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>

#include "easel.h"
#include "esl_alphabet.h"
#include "esl_random.h"
#include "esl_randomseq.h"
#include "esl_sq.h"
#include "esl_vectorops.h"
#include "esl_gumbel.h"
#include "esl_exponential.h"
#include "hmmer.h"

#include "../include/p7gpu.h"

typedef struct {
  ESL_ALPHABET *abc;
  P7_HMM       *hmm;
  P7_BG        *bg;
  P7_PROFILE   *gm;
  P7_OPROFILE  *om;
  P7_OMX       *ox;    /* one-row matrix + xmx specials (parsers, filters) */
  P7_OMX       *oxb;
} P7REF;

static int g_inited = 0;
static void ref_init_once(void)
{
  if (!g_inited) { p7_FLogsumInit(); g_inited = 1; }
  impl_Init();     /* FTZ/DAZ in this thread, as hmmsearch.c:279,1384 */
}

static P7REF *finish_handle(ESL_ALPHABET *abc, P7_HMM *hmm)
{
  P7REF *h = calloc(1, sizeof(P7REF));
  h->abc = abc;
  h->hmm = hmm;
  h->bg  = p7_bg_Create(abc);
  h->gm  = p7_profile_Create(hmm->M, abc);
  h->om  = p7_oprofile_Create(hmm->M, abc);
  p7_ProfileConfig(hmm, h->bg, h->gm, 100, p7_LOCAL);   /* hmmsearch.c:500 */
  p7_oprofile_Convert(h->gm, h->om);                    /* hmmsearch.c:503 */
  p7_bg_SetFilter(h->bg, h->om->M, h->om->compo);       /* p7_pli_NewModel, p7_pipeline.c:505 */
  h->ox  = p7_omx_Create(hmm->M, 0, 400);
  h->oxb = p7_omx_Create(hmm->M, 0, 400);
  return h;
}

/* Open the first HMM in <path> (HMMER3 or HMMER2 ASCII; p7_hmmfile.c). */
void *p7ref_open_hmm(const char *path)
{
  P7_HMMFILE   *hfp = NULL;
  ESL_ALPHABET *abc = NULL;
  P7_HMM       *hmm = NULL;
  char errbuf[eslERRBUFSIZE];
  ref_init_once();
  if (p7_hmmfile_OpenE((char *) path, NULL, &hfp, errbuf) != eslOK) { fprintf(stderr, "p7ref_open_hmm: %s\n", errbuf); return NULL; }
  if (p7_hmmfile_Read(hfp, &abc, &hmm) != eslOK) { p7_hmmfile_Close(hfp); return NULL; }
  p7_hmmfile_Close(hfp);
  return finish_handle(abc, hmm);
}

/* Synthetic exact-M profile, the recipe of SURVEY.md §8d(i): p7_hmm_Sample + SetComposition (+ p7_Calibrate). */
void *p7ref_sample_hmm(int M, int seed, int calibrate)
{
  ESL_ALPHABET   *abc = esl_alphabet_Create(eslAMINO);
  ESL_RANDOMNESS *r   = esl_randomness_CreateFast(seed);
  P7_HMM         *hmm = NULL;
  char name[64];
  ref_init_once();
  p7_hmm_Sample(r, M, abc, &hmm);
  snprintf(name, sizeof(name), "synth_M%d_s%d", M, seed);
  p7_hmm_SetName(hmm, name);
  p7_hmm_SetComposition(hmm);
  p7_hmm_SetConsensus(hmm, NULL);
  if (calibrate) {
    /* p7_Calibrate with default builder parameters (evalues.c:60-140); deterministic (seed 42 inside) */
    if (p7_Calibrate(hmm, NULL, NULL, NULL, NULL, NULL) != eslOK) { fprintf(stderr, "p7ref_sample_hmm: calibration failed\n"); return NULL; }
  } else {
    hmm->evparam[p7_MMU] = -8.0f; hmm->evparam[p7_MLAMBDA] = 0.693f; hmm->evparam[p7_VMU] = -9.0f;
    hmm->evparam[p7_VLAMBDA] = 0.693f; hmm->evparam[p7_FTAU] = -4.0f; hmm->evparam[p7_FLAMBDA] = 0.693f;
    hmm->flags |= p7H_STATS;
  }
  esl_randomness_Destroy(r);
  return finish_handle(abc, hmm);
}

int p7ref_write_hmm(void *vh, const char *path)
{
  P7REF *h = vh;
  FILE *fp = fopen(path, "w");
  if (!fp) return eslFAIL;
  p7_hmmfile_WriteASCII(fp, -1, h->hmm);
  fclose(fp);
  return eslOK;
}

void p7ref_close(void *vh)
{
  P7REF *h = vh;
  if (!h) return;
  p7_omx_Destroy(h->ox); p7_omx_Destroy(h->oxb);
  p7_oprofile_Destroy(h->om); p7_profile_Destroy(h->gm); p7_bg_Destroy(h->bg);
  p7_hmm_Destroy(h->hmm); esl_alphabet_Destroy(h->abc);
  free(h);
}

int p7ref_M(void *vh)  { return ((P7REF *) vh)->om->M; }
const char *p7ref_name(void *vh) { return ((P7REF *) vh)->om->name; }

/* The reference-side half of the drop-in boundary: fill a p7gpu_oprofile_view from a P7_OPROFILE.
 * (INTEGRATION.md shows this same function as the stub a maintainer adds.) */
void p7ref_get_view(void *vh, p7gpu_oprofile_view *v)
{
  P7REF *h = vh; const P7_OPROFILE *om = h->om;
  int s, t;
  memset(v, 0, sizeof(*v));
  v->M = om->M; v->Kp = om->abc->Kp; v->K = om->abc->K; v->mode = om->mode; v->nj = om->nj;
  v->rbv = (const uint8_t *) om->rbv[0];  v->sbv = (const uint8_t *) om->sbv[0];
  v->tbm_b = om->tbm_b; v->tec_b = om->tec_b; v->tjb_b = om->tjb_b; v->base_b = om->base_b; v->bias_b = om->bias_b;
  v->scale_b = om->scale_b;
  v->rwv = (const int16_t *) om->rwv[0];  v->twv = (const int16_t *) om->twv;
  for (s = 0; s < 4; s++) for (t = 0; t < 2; t++) { v->xw[s][t] = om->xw[s][t]; v->xf[s][t] = om->xf[s][t]; }
  v->scale_w = om->scale_w; v->base_w = om->base_w; v->ddbound_w = om->ddbound_w;
  v->rfv = (const float *) om->rfv[0];    v->tfv = (const float *) om->tfv;
  for (s = 0; s < 6;  s++) v->evparam[s] = om->evparam[s];
  for (s = 0; s < 20; s++) v->compo[s]   = om->compo[s];
}

/* sample a homologous target from the profile (p7_ProfileEmit, emit.c), length configured to Lcfg.
 * Returns the sequence length (<= maxL) or -1; dsq receives sentinel-framed residues. */
int p7ref_emit(void *vh, int seed, int Lcfg, uint8_t *dsq, int maxL)
{
  P7REF *h = vh; ESL_RANDOMNESS *r = esl_randomness_CreateFast(seed);
  ESL_SQ *sq = esl_sq_CreateDigital(h->abc); int n = -1;
  p7_ReconfigLength(h->gm, Lcfg);
  if (p7_ProfileEmit(r, h->hmm, h->gm, h->bg, sq, NULL) == eslOK && sq->n <= maxL) {
    n = (int) sq->n; memcpy(dsq, sq->dsq, n + 2);
  }
  p7_ReconfigLength(h->gm, 100);
  esl_sq_Destroy(sq); esl_randomness_Destroy(r);
  return n;
}

/* digitize text with the reference alphabet (esl_abc_Digitize); dsq must hold L+2 bytes */
int p7ref_digitize(void *vh, const char *seq, uint8_t *dsq)
{
  return esl_abc_Digitize(((P7REF *) vh)->abc, seq, dsq);
}
void p7ref_background(void *vh, float *f20) { esl_vec_FCopy(((P7REF *) vh)->bg->f, 20, f20); }

/* bias-filter HMM parameters after p7_bg_SetFilter: t[2][3], pi[3], eo[Kp][2] (esl_hmm.h) */
void p7ref_get_filter_hmm(void *vh, float *t6, float *pi3, float *eo)
{
  P7REF *h = vh; ESL_HMM *f = h->bg->fhmm; int x, k;
  for (k = 0; k < 2; k++) for (x = 0; x < 3; x++) t6[k*3+x] = f->t[k][x];
  for (x = 0; x < 3; x++) pi3[x] = f->pi[x];
  for (x = 0; x < h->abc->Kp; x++) for (k = 0; k < 2; k++) eo[x*2+k] = f->eo[x][k];
}
float p7ref_omega(void *vh) { return ((P7REF *) vh)->bg->omega; }

/* ---- per-target stage calls; each configures length first as hmmsearch.c:1404-1405 does ---- */
static void set_length(P7REF *h, int L)
{
  p7_bg_SetLength(h->bg, L);
  p7_oprofile_ReconfigLength(h->om, L);
}

int p7ref_ssv(void *vh, const uint8_t *dsq, int L, float *sc)
{ P7REF *h = vh; set_length(h, L); *sc = 0.0f; return p7_SSVFilter(dsq, L, h->om, sc); }

int p7ref_msv(void *vh, const uint8_t *dsq, int L, float *sc)
{ P7REF *h = vh; set_length(h, L); p7_omx_GrowTo(h->ox, h->om->M, 0, L); return p7_MSVFilter(dsq, L, h->om, h->ox, sc); }

int p7ref_vit(void *vh, const uint8_t *dsq, int L, float *sc)
{ P7REF *h = vh; set_length(h, L); p7_omx_GrowTo(h->ox, h->om->M, 0, L); return p7_ViterbiFilter(dsq, L, h->om, h->ox, sc); }

/* xmx (optional) receives (L+1)*6 floats */
int p7ref_fwd(void *vh, const uint8_t *dsq, int L, float *sc, float *xmx, float *totscale)
{
  P7REF *h = vh; int st;
  set_length(h, L); p7_omx_GrowTo(h->ox, h->om->M, 0, L);
  st = p7_ForwardParser(dsq, L, h->om, h->ox, sc);
  if (xmx) memcpy(xmx, h->ox->xmx, sizeof(float) * (size_t)(L+1) * p7X_NXCELLS);
  if (totscale) *totscale = h->ox->totscale;
  return st;
}

/* Forward then Backward parser, as p7_pipeline.c:727-735 */
int p7ref_fwdbck(void *vh, const uint8_t *dsq, int L, float *fsc, float *bsc, float *fxmx, float *bxmx, float *ftot, float *btot, int *b_own_scales)
{
  P7REF *h = vh; int st;
  set_length(h, L); p7_omx_GrowTo(h->ox, h->om->M, 0, L); p7_omx_GrowTo(h->oxb, h->om->M, 0, L);
  st = p7_ForwardParser(dsq, L, h->om, h->ox, fsc);
  if (st != eslOK) return st;
  st = p7_BackwardParser(dsq, L, h->om, h->ox, h->oxb, bsc);
  if (fxmx) memcpy(fxmx, h->ox->xmx,  sizeof(float) * (size_t)(L+1) * p7X_NXCELLS);
  if (bxmx) memcpy(bxmx, h->oxb->xmx, sizeof(float) * (size_t)(L+1) * p7X_NXCELLS);
  if (ftot) *ftot = h->ox->totscale;
  if (btot) *btot = h->oxb->totscale;
  if (b_own_scales) *b_own_scales = h->oxb->has_own_scales;
  return st;
}

/* p7_DomainDecoding (impl_sse/decoding.c:166-200) on externally supplied specials, then the region walk of
 * p7_domaindef_ByPosteriorHeuristics (p7_domaindef.c:412-487) WITHOUT the per-region rescoring: returns mocc/btot/etot
 * [(L+1) each, optional], the number of regions the walk would open, and whether the rt1 trigger ever fired. */
int p7ref_domain_decoding(void *vh, int L, const float *fxmx, const float *bxmx, int b_own_scales,
                          float *btot, float *etot, float *mocc, int *nregions, int *triggered_any)
{
  P7REF *h = vh; int st, i, j, trig = FALSE, any = FALSE, nreg = 0;
  ESL_RANDOMNESS *rng = esl_randomness_CreateFast(42);
  P7_DOMAINDEF *ddef = p7_domaindef_Create(rng);
  set_length(h, L); p7_omx_GrowTo(h->ox, h->om->M, 0, L); p7_omx_GrowTo(h->oxb, h->om->M, 0, L);
  memcpy(h->ox->xmx,  fxmx, sizeof(float) * (size_t)(L+1) * p7X_NXCELLS);
  memcpy(h->oxb->xmx, bxmx, sizeof(float) * (size_t)(L+1) * p7X_NXCELLS);
  h->ox->M = h->oxb->M = h->om->M; h->ox->L = h->oxb->L = L;
  h->ox->has_own_scales = TRUE; h->oxb->has_own_scales = b_own_scales;
  p7_domaindef_GrowTo(ddef, L);
  st = p7_DomainDecoding(h->om, h->ox, h->oxb, ddef);
  if (btot) memcpy(btot, ddef->btot, sizeof(float) * (L+1));
  if (etot) memcpy(etot, ddef->etot, sizeof(float) * (L+1));
  if (mocc) memcpy(mocc, ddef->mocc, sizeof(float) * (L+1));
  i = -1;
  for (j = 1; j <= L; j++) {
    if (! trig) {
      if      (ddef->mocc[j] - (ddef->btot[j] - ddef->btot[j-1]) < ddef->rt2) i = j;
      else if (i == -1)                                                       i = j;
      if      (ddef->mocc[j] >= ddef->rt1) { trig = TRUE; any = TRUE; }
    } else if (ddef->mocc[j] - (ddef->etot[j] - ddef->etot[j-1]) < ddef->rt2) { nreg++; i = -1; trig = FALSE; }
  }
  if (nregions) *nregions = nreg;
  if (triggered_any) *triggered_any = any;
  p7_domaindef_Destroy(ddef); esl_randomness_Destroy(rng);
  return st;
}

/* The reference's own FASTA reader in digital mode (esl_sqfile_OpenDigital + esl_sqio_Read, easel/esl_sqio.c,
 * esl_sqio_ascii.c), packed the way include/p7gpu.h section 4 lays a database out.  Returns eslOK, or the reader's error
 * code with its message copied to errbuf.  names: name\0desc\0 per record (may be NULL). */
#include "esl_sqio.h"
int p7ref_read_fasta(const char *path, uint8_t *residues, int64_t cap_bytes, int64_t *offsets, int32_t *lengths, int64_t cap_seqs,
                     int64_t *ret_nseq, int64_t *ret_nbytes, char *names, int64_t cap_names, char *errbuf, int errbuf_len)
{
  ESL_ALPHABET *abc = esl_alphabet_Create(eslAMINO);
  ESL_SQFILE *sqfp = NULL; ESL_SQ *sq = NULL; int status; int64_t n = 0, pos = 1, npos = 0;
  ref_init_once();
  if (errbuf && errbuf_len) errbuf[0] = 0;
  status = esl_sqfile_OpenDigital(abc, path, eslSQFILE_FASTA, NULL, &sqfp);
  if (status != eslOK) { esl_alphabet_Destroy(abc); return status; }
  sq = esl_sq_CreateDigital(abc);
  if (cap_bytes > 0) residues[0] = eslDSQ_SENTINEL;
  while ((status = esl_sqio_Read(sqfp, sq)) == eslOK) {
    if (n >= cap_seqs || pos + sq->n + 1 > cap_bytes) { status = eslEMEM; break; }
    offsets[n] = pos; lengths[n] = (int32_t) sq->n;
    memcpy(residues + pos, sq->dsq + 1, (size_t) sq->n); residues[pos + sq->n] = eslDSQ_SENTINEL;
    pos += sq->n + 1;
    if (names) {
      int64_t ln = strlen(sq->name) + 1, ld = strlen(sq->desc) + 1;
      if (npos + ln + ld <= cap_names) { memcpy(names + npos, sq->name, ln); memcpy(names + npos + ln, sq->desc, ld); npos += ln + ld; }
    }
    n++;
    esl_sq_Reuse(sq);
  }
  if (status != eslEOF && errbuf && errbuf_len) snprintf(errbuf, errbuf_len, "%s", esl_sqfile_GetErrorBuf(sqfp));
  *ret_nseq = n; *ret_nbytes = pos;
  esl_sq_Destroy(sq); esl_sqfile_Close(sqfp); esl_alphabet_Destroy(abc);
  return status == eslEOF ? eslOK : status;
}

float p7ref_nullsc(void *vh, int L)
{ P7REF *h = vh; float sc; p7_bg_SetLength(h->bg, L); p7_bg_NullOne(h->bg, NULL, L, &sc); return sc; }

float p7ref_filtersc(void *vh, const uint8_t *dsq, int L)
{ P7REF *h = vh; float sc; p7_bg_SetLength(h->bg, L); p7_bg_FilterScore(h->bg, dsq, L, &sc); return sc; }

/* per-L scalars the reference derives in p7_oprofile_ReconfigLength (p7_oprofile.c:1085-1127) */
void p7ref_length_scalars(void *vh, int L, int *tjb_b, int *xw_move, float *pmove, float *ploop)
{
  P7REF *h = vh; set_length(h, L);
  *tjb_b = h->om->tjb_b; *xw_move = h->om->xw[p7O_N][p7O_MOVE];
  *pmove = h->om->xf[p7O_N][p7O_MOVE]; *ploop = h->om->xf[p7O_N][p7O_LOOP];
}

double p7ref_gumbel_surv(double x, double mu, double lambda) { return esl_gumbel_surv(x, mu, lambda); }
double p7ref_exp_surv   (double x, double mu, double lambda) { return esl_exp_surv(x, mu, lambda); }

/* ---- batched stage calls over a packed DB (include/p7gpu.h §4 layout); which: 0 SSV 1 MSV 2 Vit 3 Fwd ---- */
int p7ref_stage_batch(void *vh, int which, const uint8_t *residues, const int64_t *offsets, const int32_t *lengths,
                      const int64_t *idx, int64_t n, float *sc, int32_t *status)
{
  P7REF *h = vh; int64_t j;
  for (j = 0; j < n; j++) {
    int64_t s = idx ? idx[j] : j;
    const uint8_t *dsq = residues + offsets[s] - 1;
    int L = lengths[s];
    switch (which) {
    case 0: status[j] = p7ref_ssv(h, dsq, L, &sc[j]); break;
    case 1: status[j] = p7ref_msv(h, dsq, L, &sc[j]); break;
    case 2: status[j] = p7ref_vit(h, dsq, L, &sc[j]); break;
    case 3: status[j] = p7ref_fwd(h, dsq, L, &sc[j], NULL, NULL); break;
    default: return eslEINVAL;
    }
  }
  return eslOK;
}

/* ---- multi-threaded kernel throughput for the CPU baseline: every thread owns a clone, like hmmsearch.c:505-516 ---- */
typedef struct { P7REF *h; int which; const uint8_t *res; const int64_t *off; const int32_t *len; int64_t lo, hi; float *sc; int32_t *st; } MTJOB;
static void *mt_worker(void *arg)
{
  MTJOB *j = arg; P7REF w = *j->h; int64_t s;
  impl_Init();
  w.om  = p7_oprofile_Clone(j->h->om);
  w.bg  = p7_bg_Clone(j->h->bg);
  w.ox  = p7_omx_Create(w.om->M, 0, 400);
  w.oxb = NULL;
  for (s = j->lo; s < j->hi; s++) {
    const uint8_t *dsq = j->res + j->off[s] - 1; int L = j->len[s];
    switch (j->which) {
    case 0: j->st[s] = p7ref_ssv(&w, dsq, L, &j->sc[s]); break;
    case 1: j->st[s] = p7ref_msv(&w, dsq, L, &j->sc[s]); break;
    case 2: j->st[s] = p7ref_vit(&w, dsq, L, &j->sc[s]); break;
    default: j->st[s] = p7ref_fwd(&w, dsq, L, &j->sc[s], NULL, NULL); break;
    }
  }
  p7_omx_Destroy(w.ox); p7_bg_Destroy(w.bg); p7_oprofile_Destroy(w.om);
  return NULL;
}
int p7ref_stage_batch_mt(void *vh, int which, int nthreads, const uint8_t *residues, const int64_t *offsets,
                         const int32_t *lengths, int64_t n, float *sc, int32_t *status)
{
  pthread_t *th = malloc(sizeof(pthread_t) * nthreads); MTJOB *jobs = malloc(sizeof(MTJOB) * nthreads); int t;
  for (t = 0; t < nthreads; t++) {
    jobs[t] = (MTJOB){ vh, which, residues, offsets, lengths, n * t / nthreads, n * (t+1) / nthreads, sc, status };
    pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
  }
  for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
  return eslOK;
}

/* ---- the whole reference pipeline over a packed DB: p7_pli_NewSeq; p7_bg_SetLength; ReconfigLength; p7_Pipeline
 *      (hmmsearch.c:1399-1411), then SortBySortkey + Threshold (hmmsearch.c:549-550). ---- */
typedef struct {
  int64_t seqidx; float score, pre_score, sum_score; double lnP; int32_t ndom, nreported, nincluded, flags;
  float best_dom_score; int32_t best_dom_ienv, best_dom_jenv;
} P7REF_HIT;

typedef struct { P7_PIPELINE *pli; P7_TOPHITS *th; } P7REF_RUN;

static P7_PIPELINE *make_pli(P7REF *h, double F1, double F2, double F3, int do_max, int nobias)
{
  P7_PIPELINE *pli = p7_pipeline_Create(NULL, h->om->M, 400, FALSE, p7_SEARCH_SEQS);
  pli->F1 = F1; pli->F2 = F2; pli->F3 = F3;
  if (do_max) { pli->do_max = TRUE; pli->do_biasfilter = FALSE; pli->F1 = pli->F2 = pli->F3 = 1.0; }  /* p7_pipeline.c:222-229 */
  if (nobias) pli->do_biasfilter = FALSE;
  p7_pli_NewModel(pli, h->om, h->bg);
  return pli;
}

static int64_t name_to_idx(const char *name) { return atoll(name + 3); }   /* names are "seq%lld" */

static void export_hits(P7_TOPHITS *th, P7REF_HIT *out, int64_t maxhits, int64_t *nhits)
{
  uint64_t i; int64_t n = 0;
  for (i = 0; i < th->N; i++) {
    P7_HIT *hit = th->hit[i];
    if (n >= maxhits) break;
    out[n].seqidx = name_to_idx(hit->name);
    out[n].score = hit->score; out[n].pre_score = hit->pre_score; out[n].sum_score = hit->sum_score; out[n].lnP = hit->lnP;
    out[n].ndom = hit->ndom; out[n].nreported = hit->nreported; out[n].nincluded = hit->nincluded; out[n].flags = hit->flags;
    out[n].best_dom_score = hit->dcl[hit->best_domain].bitscore;
    out[n].best_dom_ienv  = hit->dcl[hit->best_domain].ienv;
    out[n].best_dom_jenv  = hit->dcl[hit->best_domain].jenv;
    n++;
  }
  *nhits = n;
}

/* counters[6] = nseqs, nres, n_past_msv, n_past_bias, n_past_vit, n_past_fwd.  tblout_path may be NULL. */
int p7ref_pipeline(void *vh, const uint8_t *residues, const int64_t *offsets, const int32_t *lengths, int64_t n,
                   double F1, double F2, double F3, int do_max, int nobias,
                   uint64_t *counters, P7REF_HIT *hits, int64_t maxhits, int64_t *nhits, const char *tblout_path)
{
  P7REF *h = vh; int64_t s; char name[32];
  P7_PIPELINE *pli = make_pli(h, F1, F2, F3, do_max, nobias);
  P7_TOPHITS  *th  = p7_tophits_Create();
  for (s = 0; s < n; s++) {
    ESL_SQ *sq;
    snprintf(name, sizeof(name), "seq%lld", (long long) s);
    sq = esl_sq_CreateDigitalFrom(h->abc, name, residues + offsets[s] - 1, lengths[s], NULL, NULL, NULL);
    p7_pli_NewSeq(pli, sq);
    p7_bg_SetLength(h->bg, sq->n);
    p7_oprofile_ReconfigLength(h->om, sq->n);
    p7_Pipeline(pli, h->om, h->bg, sq, th);
    esl_sq_Destroy(sq);
    p7_pipeline_Reuse(pli);
  }
  p7_tophits_SortBySortkey(th);
  p7_tophits_Threshold(th, pli);
  counters[0] = pli->nseqs; counters[1] = pli->nres; counters[2] = pli->n_past_msv;
  counters[3] = pli->n_past_bias; counters[4] = pli->n_past_vit; counters[5] = pli->n_past_fwd;
  export_hits(th, hits, maxhits, nhits);
  if (tblout_path) { FILE *fp = fopen(tblout_path, "w"); if (fp) { p7_tophits_TabularTargets(fp, h->om->name, h->om->acc, th, pli, TRUE); fclose(fp); } }
  p7_tophits_Destroy(th); p7_pipeline_Destroy(pli);
  return eslOK;
}

/* ---- finish the pipeline from externally computed Forward/Backward specials (p7_pipeline.c:737-898):
 * the survivor records come from the GPU path; this is the unchanged reference host code downstream.
 * For each survivor: seqidx, fwdsc, fwd/bck xmx [(L+1)*6], totscales.  Z = total #targets searched. ---- */
int p7ref_finish_from_specials(void *vh, const uint8_t *residues, const int64_t *offsets, const int32_t *lengths,
                               int64_t nsurv, const int64_t *seqidx, const float *fwdsc,
                               const float *fwd_xmx, const float *bck_xmx, const int64_t *xmx_offsets,
                               const float *fwd_totscale, const float *bck_totscale,
                               int64_t nseqs_total, int64_t nres_total, int do_max, int nobias,
                               P7REF_HIT *hits, int64_t maxhits, int64_t *nhits, const char *tblout_path)
{
  P7REF *h = vh; int64_t j; char name[32];
  P7_PIPELINE *pli = make_pli(h, 0.02, 1e-3, 1e-5, do_max, nobias);
  P7_TOPHITS  *th  = p7_tophits_Create();
  extern int p7ref__posthoc(P7_PIPELINE *pli, P7_OPROFILE *om, P7_BG *bg, const ESL_SQ *sq, P7_TOPHITS *hitlist, float fwdsc);
  pli->nseqs = nseqs_total; pli->nres = nres_total; pli->Z = (double) nseqs_total;
  for (j = 0; j < nsurv; j++) {
    int64_t s = seqidx[j]; int L = lengths[s]; ESL_SQ *sq;
    snprintf(name, sizeof(name), "seq%lld", (long long) s);
    sq = esl_sq_CreateDigitalFrom(h->abc, name, residues + offsets[s] - 1, L, NULL, NULL, NULL);
    p7_bg_SetLength(h->bg, L);
    p7_oprofile_ReconfigLength(h->om, L);
    p7_omx_GrowTo(pli->oxf, h->om->M, 0, L); p7_omx_GrowTo(pli->oxb, h->om->M, 0, L);
    memcpy(pli->oxf->xmx, fwd_xmx + xmx_offsets[j], sizeof(float) * (size_t)(L+1) * p7X_NXCELLS);
    memcpy(pli->oxb->xmx, bck_xmx + xmx_offsets[j], sizeof(float) * (size_t)(L+1) * p7X_NXCELLS);
    pli->oxf->M = pli->oxb->M = h->om->M; pli->oxf->L = pli->oxb->L = L;
    pli->oxf->totscale = fwd_totscale[j]; pli->oxb->totscale = bck_totscale[j];
    pli->oxf->has_own_scales = TRUE; pli->oxb->has_own_scales = FALSE;
    p7ref__posthoc(pli, h->om, h->bg, sq, th, fwdsc[j]);
    esl_sq_Destroy(sq);
    p7_pipeline_Reuse(pli);
  }
  p7_tophits_SortBySortkey(th);
  p7_tophits_Threshold(th, pli);
  export_hits(th, hits, maxhits, nhits);
  if (tblout_path) { FILE *fp = fopen(tblout_path, "w"); if (fp) { p7_tophits_TabularTargets(fp, h->om->name, h->om->acc, th, pli, TRUE); fclose(fp); } }
  p7_tophits_Destroy(th); p7_pipeline_Destroy(pli);
  return eslOK;
}

/* Our own sequencing of the reference calls that follow p7_BackwardParser in p7_Pipeline
 * (p7_pipeline.c:737-898): domain definition, null2 correction, reconstruction score, hit creation.
 * Every arithmetic step is a call into, or the documented formula of, the reference. */
int p7ref__posthoc(P7_PIPELINE *pli, P7_OPROFILE *om, P7_BG *bg, const ESL_SQ *sq, P7_TOPHITS *hitlist, float fwdsc)
{
  P7_HIT *hit = NULL; float nullsc, seqbias, seq_score, sum_score, pre_score, pre2_score; double lnP; int Ld, d, status;
  p7_bg_NullOne(bg, sq->dsq, sq->n, &nullsc);
  status = p7_domaindef_ByPosteriorHeuristics(sq, om, pli->oxf, pli->oxb, pli->fwd, pli->bck, pli->ddef, bg, FALSE, NULL, NULL, NULL);
  if (status != eslOK) return status;
  if (pli->ddef->nregions == 0 || pli->ddef->nenvelopes == 0) return eslOK;
  if (pli->do_null2) { seqbias = esl_vec_FSum(pli->ddef->n2sc, sq->n+1); seqbias = p7_FLogsum(0.0, log(bg->omega) + seqbias); }
  else seqbias = 0.0;
  pre_score = (fwdsc - nullsc) / eslCONST_LOG2;
  seq_score = (fwdsc - (nullsc + seqbias)) / eslCONST_LOG2;
  sum_score = 0.0f; seqbias = 0.0f; Ld = 0;
  for (d = 0; d < pli->ddef->ndom; d++) {
    float corr = pli->do_null2 ? pli->ddef->dcl[d].domcorrection : 0.0f;
    if (pli->ddef->dcl[d].envsc - corr > 0.0) {
      sum_score += pli->ddef->dcl[d].envsc;
      Ld        += pli->ddef->dcl[d].jenv - pli->ddef->dcl[d].ienv + 1;
      seqbias   += corr;
    }
  }
  seqbias = pli->do_null2 ? p7_FLogsum(0.0, log(bg->omega) + seqbias) : 0.0f;
  sum_score += (sq->n - Ld) * log((float) sq->n / (float) (sq->n+3));
  pre2_score = (sum_score - nullsc) / eslCONST_LOG2;
  sum_score  = (sum_score - (nullsc + seqbias)) / eslCONST_LOG2;
  if (Ld > 0 && sum_score > seq_score) { seq_score = sum_score; pre_score = pre2_score; }
  lnP = esl_exp_logsurv(seq_score, om->evparam[p7_FTAU], om->evparam[p7_FLAMBDA]);
  if (! p7_pli_TargetReportable(pli, seq_score, lnP)) return eslOK;
  p7_tophits_CreateNextHit(hitlist, &hit);
  esl_strdup(sq->name, -1, &(hit->name));
  hit->ndom = pli->ddef->ndom; hit->nexpected = pli->ddef->nexpected; hit->nregions = pli->ddef->nregions;
  hit->nclustered = pli->ddef->nclustered; hit->noverlaps = pli->ddef->noverlaps; hit->nenvelopes = pli->ddef->nenvelopes;
  hit->pre_score = pre_score; hit->pre_lnP = esl_exp_logsurv(hit->pre_score, om->evparam[p7_FTAU], om->evparam[p7_FLAMBDA]);
  hit->score = seq_score; hit->lnP = lnP; hit->sortkey = pli->inc_by_E ? -lnP : seq_score;
  hit->sum_score = sum_score; hit->sum_lnP = esl_exp_logsurv(hit->sum_score, om->evparam[p7_FTAU], om->evparam[p7_FLAMBDA]);
  hit->dcl = pli->ddef->dcl; pli->ddef->dcl = NULL; hit->best_domain = 0;
  for (d = 0; d < hit->ndom; d++) {
    Ld = hit->dcl[d].jenv - hit->dcl[d].ienv + 1;
    hit->dcl[d].bitscore = hit->dcl[d].envsc + (sq->n - Ld) * log((float) sq->n / (float) (sq->n+3));
    hit->dcl[d].dombias  = (pli->do_null2 ? p7_FLogsum(0.0, log(bg->omega) + hit->dcl[d].domcorrection) : 0.0);
    hit->dcl[d].bitscore = (hit->dcl[d].bitscore - (nullsc + hit->dcl[d].dombias)) / eslCONST_LOG2;
    hit->dcl[d].lnP      = esl_exp_logsurv(hit->dcl[d].bitscore, om->evparam[p7_FTAU], om->evparam[p7_FLAMBDA]);
    if (hit->dcl[d].bitscore > hit->dcl[hit->best_domain].bitscore) hit->best_domain = d;
  }
  return eslOK;
}
