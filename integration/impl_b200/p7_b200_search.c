/* p7_b200_search.c — the batched worker loop of hmmsearch on the B200 engine.
 *
 * Replaces the per-target loop bodies hmmsearch.c:1295-1306 (serial_loop) and :1399-1411 (pipeline_thread):
 *
 *     p7_pli_NewSeq; p7_bg_SetLength; p7_oprofile_ReconfigLength; p7_Pipeline; esl_sq_Reuse; p7_pipeline_Reuse
 *
 * Stage-major instead of target-major (legal: every stage is a pure function of (profile, target)):
 *   1. the whole packed database goes through MSV -> bias -> Viterbi -> Forward -> Backward on the GPU(s)
 *      (p7gpu_pipeline_run_file_sharded: one host thread and one contiguous shard per device);
 *   2. each target that passed the Forward filter is handed, together with its five DP results, to the reference's
 *      UNMODIFIED p7_Pipeline() (replay, p7_b200_engine.c), which redoes the scalar P-value tests and runs domain definition,
 *      null2 correction and hit creation (p7_pipeline.c:737-898) line for line, on <ncpus> host threads with one
 *      (pipeline, profile clone, background, hit list) each — the reference's own worker structure (hmmsearch.c:505-516);
 *   3. hit lists are merged with p7_tophits_Merge and the accounting with p7_pipeline_Merge (hmmsearch.c:538-545); the
 *      counters are the GPU pipeline's (they cover every target, the replay only the survivors).
 */
#include "impl_b200.h"
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double
now_seconds(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

typedef struct {
  /* shared, read-only */
  const ESL_ALPHABET *abc;
  p7gpu_dbfile       *db;
  const uint8_t      *residues;
  const int64_t      *offsets;
  int                 nshards;
  p7gpu_pipeline_result **res;
  int64_t             nsurv_total;
  const int64_t      *surv_base;       /* prefix sums of survivors per shard: global survivor j -> (shard, local index) */
  double              Z_lower;         /* targets searched: the lower bound on Z used while hits are collected */
  /* shared, atomic */
  int64_t            *next;
  /* per worker */
  P7_PIPELINE        *pli;
  P7_OPROFILE        *om;
  P7_BG              *bg;
  P7_TOPHITS         *th;
  int                 status;
  char                errbuf[eslERRBUFSIZE];
} POST_WORKER;

#define POST_GRAB 8

static void *
post_thread(void *arg)
{
  POST_WORKER *w = (POST_WORKER *) arg;
  impl_Init();                                   /* FTZ/DAZ in this thread (hmmsearch.c:1384) */
  if (w->pli->Z_setby == p7_ZSETBY_NTARGETS) w->pli->Z = w->Z_lower;
  for (;;) {
    int64_t j0 = __atomic_fetch_add(w->next, POST_GRAB, __ATOMIC_RELAXED), j;
    if (j0 >= w->nsurv_total) break;
    for (j = j0; j < j0 + POST_GRAB && j < w->nsurv_total; j++) {
      int k = 0;
      const p7gpu_survivor *sv;
      P7_B200_REPLAY rec;
      ESL_SQ *sq;
      int st;
      while (k + 1 < w->nshards && j >= w->surv_base[k + 1]) k++;
      sv = p7gpu_result_survivors(w->res[k]) + (j - w->surv_base[k]);
      if (sv->region_flag == 0) continue;       /* device region trigger: mocc[i] < rt1 everywhere => ddef->nregions == 0,
                                                   p7_Pipeline would return at p7_pipeline.c:739 with nothing added */
      sq = esl_sq_CreateDigitalFrom(w->abc, p7gpu_dbfile_name(w->db, sv->seqidx), w->residues + w->offsets[sv->seqidx] - 1,
                                    sv->L, p7gpu_dbfile_desc(w->db, sv->seqidx), NULL, NULL);
      if (!sq) { w->status = eslEMEM; return NULL; }
      rec.dsq = sq->dsq;  rec.L = sv->L;
      rec.msv_status = sv->msv_status;  rec.vit_status = sv->vit_status;
      rec.usc = sv->usc;  rec.vfsc = sv->vfsc;  rec.fwdsc = sv->fwdsc;  rec.bcksc = sv->bcksc;
      rec.fwd_xmx = (sv->xmx_offset >= 0) ? p7gpu_result_fwd_xmx(w->res[k]) + sv->xmx_offset : NULL;
      rec.bck_xmx = (sv->xmx_offset >= 0) ? p7gpu_result_bck_xmx(w->res[k]) + sv->xmx_offset : NULL;
      rec.dec     = (sv->dec_offset >= 0) ? p7gpu_result_decoding(w->res[k]) + sv->dec_offset : NULL;
      rec.fwd_totscale = sv->fwd_totscale;  rec.bck_totscale = sv->bck_totscale;  rec.bck_own_scales = sv->bck_own_scales;
      p7_b200_ReplayArm(&rec);
      p7_bg_SetLength(w->bg, sq->n);                       /* hmmsearch.c:1404-1405 */
      p7_oprofile_ReconfigLength(w->om, sq->n);
      st = p7_Pipeline(w->pli, w->om, w->bg, sq, w->th);
      p7_b200_ReplayArm(NULL);
      if (st != eslOK && st != eslERANGE) { w->status = st; strncpy(w->errbuf, w->pli->errbuf, eslERRBUFSIZE - 1); esl_sq_Destroy(sq); return NULL; }
      esl_sq_Destroy(sq);
      p7_pipeline_Reuse(w->pli);
    }
  }
  return NULL;
}

int
p7_b200_SearchDatabase(P7_B200 **engs, P7_B200_SEARCH *cfg, ESL_GETOPTS *go, const ESL_ALPHABET *abc,
                       P7_OPROFILE *om, P7_BG *bg, P7_PIPELINE *pli, P7_TOPHITS *th, p7gpu_dbfile *db)
{
  const int ng = ESL_MAX(1, cfg->ngpus), nw = ESL_MAX(1, cfg->ncpus);
  p7gpu_ctx **ctxs = NULL;  const p7gpu_profile **profs = NULL;
  p7gpu_pipeline_result **res = NULL;
  int64_t *first = NULL, *surv_base = NULL, next = 0;
  POST_WORKER *wk = NULL;  pthread_t *tid = NULL;
  p7gpu_pipeline_opts o;
  p7gpu_pipeline_counters tot;
  double t0, t1, t2;
  int k, i, status = eslOK;

  ESL_ALLOC(ctxs,  sizeof(*ctxs)  * ng);   ESL_ALLOC(profs, sizeof(*profs) * ng);
  ESL_ALLOC(res,   sizeof(*res)   * ng);   ESL_ALLOC(first, sizeof(*first) * (ng + 1));
  ESL_ALLOC(surv_base, sizeof(*surv_base) * (ng + 1));
  for (k = 0; k < ng; k++) { ctxs[k] = p7_b200_Ctx(engs[k]); profs[k] = p7_b200_Profile(engs[k]); res[k] = NULL; }

  memset(&o, 0, sizeof(o));
  o.F1 = pli->F1;  o.F2 = pli->F2;  o.F3 = pli->F3;
  o.do_biasfilter = pli->do_biasfilter;  o.do_max = pli->do_max;
  o.want_specials = 1;
  o.fb_exact = cfg->fb_exact;
  o.region_trigger = (cfg->fb_exact || pli->do_max) ? 1 : 0;   /* exact on exact specials; otherwise only where it pays (--max) */
  o.decode = (cfg->decode && o.region_trigger) ? 1 : 0;

  t0 = now_seconds();
  if (p7gpu_pipeline_run_file_sharded(ctxs, profs, ng, db, &o, first, res) != P7GPU_OK)
    ESL_XEXCEPTION(eslESYS, "impl_b200: GPU pipeline failed: %s", p7gpu_last_error());
  t1 = now_seconds();

  memset(&tot, 0, sizeof(tot));
  surv_base[0] = 0;
  for (k = 0; k < ng; k++) {
    p7gpu_counters_merge(&tot, p7gpu_result_counters(res[k]));
    surv_base[k + 1] = surv_base[k] + p7gpu_result_nsurvivors(res[k]);
  }

  /* survivors -> unmodified p7_Pipeline on <nw> host threads */
  ESL_ALLOC(wk,  sizeof(*wk)  * nw);
  ESL_ALLOC(tid, sizeof(*tid) * nw);
  memset(wk, 0, sizeof(*wk) * nw);
  for (i = 0; i < nw; i++) {
    wk[i].abc = abc;  wk[i].db = db;  wk[i].residues = p7gpu_dbfile_residues(db);  wk[i].offsets = p7gpu_dbfile_offsets(db);
    wk[i].nshards = ng;  wk[i].res = res;  wk[i].nsurv_total = surv_base[ng];  wk[i].surv_base = surv_base;  wk[i].next = &next;
    wk[i].Z_lower = (double) (pli->nseqs + tot.nseqs);
    wk[i].status = eslOK;
    if (i == 0) { wk[i].pli = pli; wk[i].om = om; wk[i].bg = bg; wk[i].th = th; }
    else {
      wk[i].th  = p7_tophits_Create();
      wk[i].om  = p7_oprofile_Clone(om);
      wk[i].bg  = p7_bg_Create(abc);
      wk[i].pli = p7_pipeline_Create(go, om->M, 100, FALSE, p7_SEARCH_SEQS);
      if (!wk[i].th || !wk[i].om || !wk[i].bg || !wk[i].pli) ESL_XEXCEPTION(eslEMEM, "allocation failure");
      p7_pli_NewModel(wk[i].pli, wk[i].om, wk[i].bg);
    }
  }
  {
    /* worker 0 works on the caller's pipeline: remember its accounting, the replay only sees survivors */
    const uint64_t s_nseqs = pli->nseqs, s_nres = pli->nres, s_msv = pli->n_past_msv, s_bias = pli->n_past_bias,
                   s_vit = pli->n_past_vit, s_fwd = pli->n_past_fwd;
    const double   s_Z = pli->Z;
    for (i = 1; i < nw; i++) if (pthread_create(&tid[i], NULL, post_thread, &wk[i]) != 0) ESL_XEXCEPTION(eslESYS, "pthread_create failed");
    post_thread(&wk[0]);
    for (i = 1; i < nw; i++) pthread_join(tid[i], NULL);
    for (i = 0; i < nw; i++) if (wk[i].status != eslOK) { status = wk[i].status; strncpy(pli->errbuf, wk[i].errbuf, eslERRBUFSIZE - 1); }
    pli->nseqs = s_nseqs;  pli->nres = s_nres;  pli->n_past_msv = s_msv;  pli->n_past_bias = s_bias;
    pli->n_past_vit = s_vit;  pli->n_past_fwd = s_fwd;  pli->Z = s_Z;
    for (i = 1; i < nw; i++) {
      wk[i].pli->nseqs = wk[i].pli->nres = 0;
      wk[i].pli->n_past_msv = wk[i].pli->n_past_bias = wk[i].pli->n_past_vit = wk[i].pli->n_past_fwd = 0;
      if (wk[i].pli->Z_setby == p7_ZSETBY_NTARGETS) wk[i].pli->Z = 0.;
      p7_tophits_Merge(th, wk[i].th);                      /* hmmsearch.c:540-541 */
      p7_pipeline_Merge(pli, wk[i].pli);
      p7_pipeline_Destroy(wk[i].pli);  p7_tophits_Destroy(wk[i].th);  p7_oprofile_Destroy(wk[i].om);  p7_bg_Destroy(wk[i].bg);
      wk[i].pli = NULL;
    }
    /* the accounting of p7_pli_NewSeq + p7_Pipeline over every target, as the GPU pipeline counted it */
    pli->nseqs       += tot.nseqs;        pli->nres        += tot.nres;
    pli->n_past_msv  += tot.n_past_msv;   pli->n_past_bias += tot.n_past_bias;
    pli->n_past_vit  += tot.n_past_vit;   pli->n_past_fwd  += tot.n_past_fwd;
    if (pli->Z_setby == p7_ZSETBY_NTARGETS) pli->Z = (double) pli->nseqs;      /* p7_pli_NewSeq, p7_pipeline.c:577 */
  }
  t2 = now_seconds();
  cfg->gpu_seconds = t1 - t0;  cfg->post_seconds = t2 - t1;  cfg->nsurvivors = (uint64_t) surv_base[ng];

 ERROR:   /* also the normal exit */
  if (wk) for (i = 1; i < nw; i++) if (wk[i].pli) { p7_pipeline_Destroy(wk[i].pli); p7_tophits_Destroy(wk[i].th); p7_oprofile_Destroy(wk[i].om); p7_bg_Destroy(wk[i].bg); }
  if (res) for (k = 0; k < ng; k++) p7gpu_result_destroy(res[k]);
  free(wk); free(tid); free(ctxs); free(profs); free(res); free(first); free(surv_base);
  return status;
}
