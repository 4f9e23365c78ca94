/* hmmscan_b200.c — hmmscan (query sequences against a pressed profile database) with the acceleration pipeline on a B200.
 *
 * The reference's hmmscan (src/hmmscan.c) loops  for each query: for each model: p7_oprofile_ReadMSV; p7_pli_NewModel;
 * p7_Pipeline(SCAN)  (serial_loop, hmmscan.c:1176-1198), reading the rest of a model only when a query passes its MSV and bias
 * filters (p7_pipeline.c:708-713).  Here the loops are swapped, which is legal because every stage is a pure function of
 * (model, sequence): ALL query sequences are digitised once into a packed database that stays in HBM, and each model of the
 * pressed database is streamed past it:
 *
 *   p7_oprofile_ReadMSV (the reference's reader, impl_sse/io.c:233)      the .h3f part only, a block of models at a time
 *   block prefilter: ONE kernel over every (model, query) pair           p7gpu_scan_msv_prefilter (few queries: most models
 *   of the block, gain table picked per work item                        end here without an upload or a launch of their own)
 *   MSV filter of this model against every query on the GPU              p7gpu_pipeline_run(stop_after_msv)
 *   nothing passed?  next model — the .h3p part is never read            (lazy rest-of-profile, io.c:493)
 *   p7_oprofile_ReadRest; full upload; bias -> Viterbi -> Forward -> Backward -> p7_DomainDecoding for the survivors
 *   per-(query, model) accounting from the stage records; every Forward survivor is finished by the reference's unmodified
 *   p7_Pipeline in SCAN mode (replay, p7_b200_engine.c) into that query's own P7_TOPHITS / P7_PIPELINE.
 *
 * Output: per query, in file order, through the reference's own p7_tophits_* / p7_pli_Statistics — byte-identical to hmmscan
 * (tests/test_integration_hmmscan.py).  One GPU; --cpu host threads finish the Forward survivors of each model (disjoint queries per thread).
 */
#include "impl_b200.h"
#include "esl_getopts.h"
#include "esl_stopwatch.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>
#include "esl_threads.h"
#include <time.h>

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec; }

#define REPOPTS     "-E,-T,--cut_ga,--cut_nc,--cut_tc"
#define DOMREPOPTS  "--domE,--domT,--cut_ga,--cut_nc,--cut_tc"
#define INCOPTS     "--incE,--incT,--cut_ga,--cut_nc,--cut_tc"
#define INCDOMOPTS  "--incdomE,--incdomT,--cut_ga,--cut_nc,--cut_tc"
#define THRESHOPTS  "-E,-T,--domE,--domT,--incE,--incT,--incdomE,--incdomT,--cut_ga,--cut_nc,--cut_tc"

/* hmmscan's option names and defaults (hmmscan.c:64-110): p7_pipeline_Create reads them by name */
static ESL_OPTIONS options[] = {
  { "-h",           eslARG_NONE,    FALSE,  NULL,         NULL,     NULL, NULL, NULL,             "show brief help on version and usage",                        1 },
  { "-o",           eslARG_OUTFILE, NULL,   NULL,         NULL,     NULL, NULL, NULL,             "direct output to file <f>, not stdout",                       2 },
  { "--tblout",     eslARG_OUTFILE, NULL,   NULL,         NULL,     NULL, NULL, NULL,             "save parseable table of per-sequence hits to file <f>",       2 },
  { "--domtblout",  eslARG_OUTFILE, NULL,   NULL,         NULL,     NULL, NULL, NULL,             "save parseable table of per-domain hits to file <f>",         2 },
  { "--pfamtblout", eslARG_OUTFILE, NULL,   NULL,         NULL,     NULL, NULL, NULL,             "save table of hits and domains to file, in Pfam format <f>",  2 },
  { "--acc",        eslARG_NONE,    FALSE,  NULL,         NULL,     NULL, NULL, NULL,             "prefer accessions over names in output",                      2 },
  { "--noali",      eslARG_NONE,    FALSE,  NULL,         NULL,     NULL, NULL, NULL,             "don't output alignments, so output is smaller",               2 },
  { "--notextw",    eslARG_NONE,    NULL,   NULL,         NULL,     NULL, NULL, "--textw",        "unlimit ASCII text output line width",                        2 },
  { "--textw",      eslARG_INT,     "120",  NULL,         "n>=120", NULL, NULL, "--notextw",      "set max width of ASCII text output lines",                    2 },
  { "-E",           eslARG_REAL,    "10.0", NULL,         "x>0",    NULL, NULL, REPOPTS,          "report models <= this E-value threshold in output",           4 },
  { "-T",           eslARG_REAL,    FALSE,  NULL,         NULL,     NULL, NULL, REPOPTS,          "report models >= this score threshold in output",             4 },
  { "--domE",       eslARG_REAL,    "10.0", NULL,         "x>0",    NULL, NULL, DOMREPOPTS,       "report domains <= this E-value threshold in output",          4 },
  { "--domT",       eslARG_REAL,    FALSE,  NULL,         NULL,     NULL, NULL, DOMREPOPTS,       "report domains >= this score cutoff in output",               4 },
  { "--incE",       eslARG_REAL,    "0.01", NULL,         "x>0",    NULL, NULL, INCOPTS,          "consider models <= this E-value threshold as significant",    5 },
  { "--incT",       eslARG_REAL,    FALSE,  NULL,         NULL,     NULL, NULL, INCOPTS,          "consider models >= this score threshold as significant",      5 },
  { "--incdomE",    eslARG_REAL,    "0.01", NULL,         "x>0",    NULL, NULL, INCDOMOPTS,       "consider domains <= this E-value threshold as significant",   5 },
  { "--incdomT",    eslARG_REAL,    FALSE,  NULL,         NULL,     NULL, NULL, INCDOMOPTS,       "consider domains >= this score threshold as significant",     5 },
  { "--cut_ga",     eslARG_NONE,    FALSE,  NULL,         NULL,     NULL, NULL, THRESHOPTS,       "use profile's GA gathering cutoffs to set all thresholding",  6 },
  { "--cut_nc",     eslARG_NONE,    FALSE,  NULL,         NULL,     NULL, NULL, THRESHOPTS,       "use profile's NC noise cutoffs to set all thresholding",      6 },
  { "--cut_tc",     eslARG_NONE,    FALSE,  NULL,         NULL,     NULL, NULL, THRESHOPTS,       "use profile's TC trusted cutoffs to set all thresholding",    6 },
  { "--max",        eslARG_NONE,    FALSE,  NULL,         NULL,     NULL, NULL, "--F1,--F2,--F3", "Turn all heuristic filters off (less speed, more power)",     7 },
  { "--F1",         eslARG_REAL,    "0.02", NULL,         NULL,     NULL, NULL, "--max",          "MSV threshold: promote hits w/ P <= F1",                      7 },
  { "--F2",         eslARG_REAL,    "1e-3", NULL,         NULL,     NULL, NULL, "--max",          "Vit threshold: promote hits w/ P <= F2",                      7 },
  { "--F3",         eslARG_REAL,    "1e-5", NULL,         NULL,     NULL, NULL, "--max",          "Fwd threshold: promote hits w/ P <= F3",                      7 },
  { "--nobias",     eslARG_NONE,    NULL,   NULL,         NULL,     NULL, NULL, "--max",          "turn off composition bias filter",                            7 },
  { "--nonull2",    eslARG_NONE,    NULL,   NULL,         NULL,     NULL, NULL, NULL,             "turn off biased composition score corrections",              12 },
  { "-Z",           eslARG_REAL,    FALSE,  NULL,         "x>0",    NULL, NULL, NULL,             "set # of comparisons done, for E-value calculation",         12 },
  { "--domZ",       eslARG_REAL,    FALSE,  NULL,         "x>0",    NULL, NULL, NULL,             "set # of significant seqs, for domain E-value calculation",  12 },
  { "--seed",       eslARG_INT,     "42",   NULL,         "n>=0",   NULL, NULL, NULL,             "set RNG seed to <n> (if 0: one-time arbitrary seed)",        12 },
  { "--qformat",    eslARG_STRING,  NULL,   NULL,         NULL,     NULL, NULL, NULL,             "assert input <seqfile> is in format <s> (fasta only)",       12 },
  { "--cpu",        eslARG_INT,     NULL,   "HMMER_NCPU", "n>=0",   NULL, NULL, NULL,             "number of host threads finishing the Forward survivors",     12 },
  { "--fbfast",     eslARG_NONE,    FALSE,  NULL,         NULL,     NULL, NULL, NULL,             "Forward/Backward specials within 1e-4 instead of bit-exact", 13 },
  { "--prefilter",  eslARG_INT,     "-1",   NULL,         "n>=-1",  NULL, NULL, NULL,             "models per (model, query) MSV prefilter block; 0 off, -1 auto", 13 },
  { "--times",      eslARG_NONE,    FALSE,  NULL,         NULL,     NULL, NULL, NULL,             "report wall time of each phase on stderr",                   13 },
  {  0, 0, 0, 0, 0, 0, 0, 0, 0, 0 },
};

static char usage[]  = "[-options] <hmmdb> <seqfile>";
static char banner[] = "search sequence(s) against a profile database";

typedef struct {                 /* one per query sequence */
  ESL_SQ      *sq;
  P7_PIPELINE *pli;
  P7_TOPHITS  *th;
} QUERY;

/* Finishing the Forward survivors of one model (the reference's own p7_Pipeline, fed by replay) is host work that is
 * independent between queries: worker t takes the survivors whose query index is t modulo the number of workers, so every
 * query's pipeline and hit list are touched by one thread only and hits arrive in model order as in hmmscan.c:1222. */
typedef struct {
  int                     t, nt;
  const p7gpu_survivor   *sv;
  int64_t                 ns;
  p7gpu_pipeline_result  *r;
  QUERY                  *qs;
  P7_OPROFILE            *om;       /* this worker's clone: p7_oprofile_ReconfigLength writes into it */
  P7_BG                  *bg;       /* this worker's own: p7_bg_SetLength / SetFilter write into it */
  uint64_t                nmodels;
  int                     status;
  char                    errbuf[eslERRBUFSIZE];
} REPLAY_WORKER;

static void *
replay_worker(void *arg)
{
  REPLAY_WORKER *w = (REPLAY_WORKER *) arg;
  int64_t k;
  impl_Init();                                                   /* FTZ/DAZ in this thread (hmmscan.c:1167) */
  w->status = eslOK;
  for (k = 0; k < w->ns; k++) {
    const p7gpu_survivor *sv = &w->sv[k];
    QUERY *Q;  P7_PIPELINE *pli;  P7_B200_REPLAY rec;
    uint64_t c0, c1, c2, c3;
    int status;
    if (sv->seqidx % w->nt != w->t || sv->region_flag == 0) continue;
    Q = &w->qs[sv->seqidx];  pli = Q->pli;
    c0 = pli->n_past_msv;  c1 = pli->n_past_bias;  c2 = pli->n_past_vit;  c3 = pli->n_past_fwd;
    memset(&rec, 0, sizeof(rec));
    rec.dsq = Q->sq->dsq;  rec.L = sv->L;
    rec.msv_status = sv->msv_status;  rec.vit_status = sv->vit_status;
    rec.usc = sv->usc;  rec.vfsc = sv->vfsc;  rec.fwdsc = sv->fwdsc;  rec.bcksc = sv->bcksc;
    rec.fwd_xmx = (sv->xmx_offset >= 0) ? p7gpu_result_fwd_xmx(w->r) + sv->xmx_offset : NULL;
    rec.bck_xmx = (sv->xmx_offset >= 0) ? p7gpu_result_bck_xmx(w->r) + sv->xmx_offset : NULL;
    rec.dec     = (sv->dec_offset >= 0) ? p7gpu_result_decoding(w->r) + sv->dec_offset : NULL;
    rec.fwd_totscale = sv->fwd_totscale;  rec.bck_totscale = sv->bck_totscale;  rec.bck_own_scales = sv->bck_own_scales;
    if (pli->Z_setby == p7_ZSETBY_NTARGETS) pli->Z = (double) w->nmodels;    /* p7_pli_NewModel, p7_pipeline.c:498: a lower bound while hits are collected */
    pli->hfp = NULL;                                             /* <om> is complete already: p7_Pipeline must not read the rest again */
    p7_b200_ReplayArm(&rec);
    p7_bg_SetLength(w->bg, Q->sq->n);                            /* hmmscan.c:1187-1188 */
    p7_oprofile_ReconfigLength(w->om, Q->sq->n);
    status = p7_Pipeline(pli, w->om, w->bg, Q->sq, Q->th);
    p7_b200_ReplayArm(NULL);
    if (status != eslOK && status != eslERANGE) { w->status = status; strncpy(w->errbuf, pli->errbuf, eslERRBUFSIZE - 1); w->errbuf[eslERRBUFSIZE - 1] = 0; return NULL; }
    p7_pipeline_Reuse(pli);
    pli->n_past_msv = c0;  pli->n_past_bias = c1;  pli->n_past_vit = c2;  pli->n_past_fwd = c3;   /* counted by the caller from the stage records */
  }
  return NULL;
}

static int cmp_int32(const void *a, const void *b) { const int32_t x = *(const int32_t *) a, y = *(const int32_t *) b; return (x > y) - (x < y); }

int
main(int argc, char **argv)
{
  ESL_GETOPTS   *go = esl_getopts_Create(options);
  char          *hmmfile, *seqfile, tmp[4096];
  FILE          *ofp = stdout, *tblfp = NULL, *domtblfp = NULL, *pfamtblfp = NULL;
  P7_HMMFILE    *hfp = NULL;
  ESL_ALPHABET  *abc = NULL;
  P7_OPROFILE   *om  = NULL;
  P7_BG         *bg  = NULL;
  P7_B200       *eng = NULL;
  p7gpu_dbfile  *qdb = NULL;
  p7gpu_seqdb   *gdb = NULL;
  QUERY         *qs  = NULL;
  ESL_STOPWATCH *w   = esl_stopwatch_Create();
  const uint8_t *residues; const int64_t *offsets; const int32_t *lengths;
  int64_t        nq = 0, nres = 0, q, k, nqlens = 0;
  REPLAY_WORKER  workers[64];
  pthread_t      tids[64];
  int            ncpus = 1;
  int32_t       *qlens = NULL;
  uint64_t       nmodels = 0, nnodes = 0, nfull = 0, nskipped = 0;
  P7_OPROFILE  **blk = NULL;                 /* the MSV parts of a block of models (p7_oprofile_ReadMSV), in file order */
  p7gpu_oprofile_view *views = NULL;
  int32_t       *npossible = NULL;
  int            blkmax = 1, nb, b, do_max;
  double         F1;
  double         t_read = 0., t_pre = 0., t_msv = 0., t_rest = 0., t_full = 0., t_upload = 0., t_replay = 0., t0;    /* --times: where the scan goes */
  int            textw, times, fb_exact, status, hstatus;
  const char    *dir = getenv("TMPDIR");

  impl_Init();
  p7_FLogsumInit();
  if (esl_opt_ProcessEnvironment(go) != eslOK || esl_opt_ProcessCmdline(go, argc, argv) != eslOK || esl_opt_VerifyConfig(go) != eslOK) {
    printf("Failed to parse command line: %s\n", go->errbuf); esl_usage(stdout, argv[0], usage); exit(1);
  }
  if (esl_opt_GetBoolean(go, "-h")) {
    p7_banner(stdout, argv[0], banner); esl_usage(stdout, argv[0], usage);
    puts("\nOptions:"); esl_opt_DisplayHelp(stdout, go, 0, 2, 80); exit(0);
  }
  if (esl_opt_ArgNumber(go) != 2) { puts("Incorrect number of command line arguments."); esl_usage(stdout, argv[0], usage); exit(1); }
  hmmfile = esl_opt_GetArg(go, 1);  seqfile = esl_opt_GetArg(go, 2);
  textw    = esl_opt_GetBoolean(go, "--notextw") ? 0 : esl_opt_GetInteger(go, "--textw");
  times    = esl_opt_GetBoolean(go, "--times");
  fb_exact = esl_opt_GetBoolean(go, "--fbfast") ? 0 : 1;
  if (esl_opt_IsOn(go, "--cpu")) ncpus = esl_opt_GetInteger(go, "--cpu"); else esl_threads_CPUCount(&ncpus);
  ncpus = ESL_MAX(1, ESL_MIN(ncpus, 64));
  memset(workers, 0, sizeof(workers));
  if (esl_opt_IsOn(go, "--qformat") && strcasecmp(esl_opt_GetString(go, "--qformat"), "fasta") != 0) p7_Fail("%s is not a query format this engine reads (fasta)\n", esl_opt_GetString(go, "--qformat"));

  /* the pressed profile database (hmmscan.c:413: p7_HMMDBENV; binary .h3m/.h3f/.h3p/.h3i set written by hmmpress) */
  status = p7_hmmfile_OpenE(hmmfile, p7_HMMDBENV, &hfp, NULL);
  if      (status == eslENOTFOUND) p7_Fail("File existence/permissions problem in trying to open HMM file %s.\n", hmmfile);
  else if (status == eslEFORMAT)   p7_Fail("File format problem, trying to open HMM file %s.\n", hmmfile);
  else if (status != eslOK)        p7_Fail("Unexpected error %d in opening HMM file %s.\n", status, hmmfile);
  if (!hfp->is_pressed)            p7_Fail("Failed to open binary auxfiles for %s: use hmmpress first\n", hfp->fname);

  if (esl_opt_IsOn(go, "-o")           && (ofp       = fopen(esl_opt_GetString(go, "-o"), "w")) == NULL)           p7_Fail("Failed to open output file %s for writing\n", esl_opt_GetString(go, "-o"));
  if (esl_opt_IsOn(go, "--tblout")     && (tblfp     = fopen(esl_opt_GetString(go, "--tblout"), "w")) == NULL)     p7_Fail("Failed to open tabular per-seq output file %s for writing\n", esl_opt_GetString(go, "--tblout"));
  if (esl_opt_IsOn(go, "--domtblout")  && (domtblfp  = fopen(esl_opt_GetString(go, "--domtblout"), "w")) == NULL)  p7_Fail("Failed to open tabular per-dom output file %s for writing\n", esl_opt_GetString(go, "--domtblout"));
  if (esl_opt_IsOn(go, "--pfamtblout") && (pfamtblfp = fopen(esl_opt_GetString(go, "--pfamtblout"), "w")) == NULL) p7_Fail("Failed to open pfam-style tabular output file %s for writing\n", esl_opt_GetString(go, "--pfamtblout"));

  p7_banner(ofp, go->argv[0], banner);
  fprintf(ofp, "# query sequence file:             %s\n", seqfile);
  fprintf(ofp, "# target HMM database:             %s\n", hmmfile);
  fprintf(ofp, "# - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - -\n\n");

  /* every query, digitised once (easel's rules: p7gpu_dbfile_from_fasta), resident on the device */
  if (getenv("CUDA_VISIBLE_DEVICES") == NULL) setenv("CUDA_VISIBLE_DEVICES", "0", 0);
  if (p7_b200_Open(0, &eng) != eslOK) p7_Fail("Failed to open GPU 0: %s\n", p7_b200_ErrorText());
  snprintf(tmp, sizeof(tmp), "%s/hmmscan_b200.%ld.p7db", dir ? dir : "/tmp", (long) getpid());
  if (p7gpu_dbfile_from_fasta(seqfile, tmp, 0, 16, &nq, &nres) != P7GPU_OK) p7_Fail("Parse failed (sequence file %s):\n%s\n", seqfile, p7gpu_last_error());
  if (p7gpu_dbfile_open(tmp, &qdb) != P7GPU_OK) p7_Fail("%s\n", p7gpu_last_error());
  unlink(tmp);
  residues = p7gpu_dbfile_residues(qdb);  offsets = p7gpu_dbfile_offsets(qdb);  lengths = p7gpu_dbfile_lengths(qdb);
  if (p7gpu_seqdb_upload(p7_b200_Ctx(eng), residues, p7gpu_dbfile_nbytes(qdb), offsets, lengths, nq, &gdb) != P7GPU_OK) p7_Fail("%s\n", p7gpu_last_error());

  /* the distinct query lengths: each model's per-length constants are built for these only */
  ESL_ALLOC(qlens, sizeof(int32_t) * ESL_MAX(nq, 1));
  for (q = 0; q < nq; q++) qlens[q] = lengths[q];
  qsort(qlens, (size_t) nq, sizeof(int32_t), cmp_int32);
  for (q = 0, nqlens = 0; q < nq; q++) if (q == 0 || qlens[q] != qlens[q - 1]) qlens[nqlens++] = qlens[q];

  /* The block prefilter answers "does any query get past MSV of this model" for 512 models in one launch (~30 us per model with
   * the table building) where the per-model MSV stage costs an upload + launch + read-back (~200 us); flagged models go
   * straight to the full upload, so it pays at any number of queries (measured, 2000 models: 1 query 0.32 -> 0.11 s, 10 queries
   * 0.74 -> 0.52 s, 25 queries 1.53 -> 1.17 s).  Off under --max, which passes everything, and beyond 2000 queries, where the
   * per-model kernel (table in shared / Tensor Memory) is the faster way through the cells. */
  do_max = esl_opt_GetBoolean(go, "--max");
  F1     = ESL_MIN(1.0, esl_opt_GetReal(go, "--F1"));                 /* p7_pipeline_Create, p7_pipeline.c:141 */
  blkmax = esl_opt_GetInteger(go, "--prefilter");
  if (blkmax < 0) blkmax = (!do_max && F1 < 1.0 && nq > 0 && nq <= 2000) ? 512 : 0;
  if (do_max || blkmax < 2) blkmax = 1;
  blkmax = ESL_MIN(blkmax, 4096);
  ESL_ALLOC(blk,       sizeof(P7_OPROFILE *) * blkmax);
  ESL_ALLOC(views,     sizeof(p7gpu_oprofile_view) * blkmax);
  ESL_ALLOC(npossible, sizeof(int32_t) * blkmax);

  esl_stopwatch_Start(w);
  /* stream the models past the resident queries */
  for (;;) {
   t0 = now_s();
   for (nb = 0; nb < blkmax && (hstatus = p7_oprofile_ReadMSV(hfp, &abc, &blk[nb])) == eslOK; nb++)
     if (abc->type != eslAMINO) p7_Fail("The B200 engine scans protein profile databases\n");
   t_read += now_s() - t0;
   if (nb == 0) break;
   for (b = 0; b < nb; b++) npossible[b] = 1;
   t0 = now_s();
   if (blkmax > 1 && nq > 0) {                                    /* every (model, query) pair of the block in one launch */
     for (b = 0; b < nb; b++) { p7_b200_ViewOf(blk[b], &views[b]); views[b].rwv = NULL; views[b].twv = NULL; views[b].rfv = NULL; views[b].tfv = NULL; }
     if (p7gpu_scan_msv_prefilter(p7_b200_Ctx(eng), views, nb, gdb, lengths, F1, npossible) != P7GPU_OK) p7_Fail("GPU prefilter failed: %s\n", p7gpu_last_error());
   }
   t_pre += now_s() - t0;
   for (b = 0; b < nb; b++) {
    p7gpu_pipeline_opts o;
    p7gpu_pipeline_result *r = NULL;
    int64_t npass = 0;

    om = blk[b];
    if (!qs) {                                                   /* the alphabet is known now: per-query pipeline + hit list */
      bg = p7_bg_Create(abc);
      ESL_ALLOC(qs, sizeof(QUERY) * ESL_MAX(nq, 1));
      for (q = 0; q < nq; q++) {
        qs[q].sq  = esl_sq_CreateDigitalFrom(abc, p7gpu_dbfile_name(qdb, q), residues + offsets[q] - 1, lengths[q], p7gpu_dbfile_desc(qdb, q), NULL, NULL);
        qs[q].pli = p7_pipeline_Create(go, 100, 100, FALSE, p7_SCAN_MODELS);       /* hmmscan.c:444-448 */
        qs[q].th  = p7_tophits_Create();
        if (!qs[q].sq || !qs[q].pli || !qs[q].th) p7_Fail("allocation failure\n");
        p7_pli_NewSeq(qs[q].pli, qs[q].sq);
      }
    }
    nmodels++;  nnodes += om->M;
    if (npossible[b] == 0) { nskipped++; p7_oprofile_Destroy(om); continue; }   /* no query can get past MSV of this model: never uploaded */

    /* stage 1: the MSV filter of this model against every query, from the .h3f part alone */
    memset(&o, 0, sizeof(o));
    o.F1 = qs ? qs[0].pli->F1 : 0.02;  o.F2 = qs ? qs[0].pli->F2 : 1e-3;  o.F3 = qs ? qs[0].pli->F3 : 1e-5;
    o.do_biasfilter = qs ? qs[0].pli->do_biasfilter : 1;  o.do_max = qs ? qs[0].pli->do_max : 0;
    o.stop_after_msv = 1;
    if (blkmax == 1) {                                           /* (with the block prefilter this question is answered already: a flagged
                                                                  *  model has a query past MSV, or one that needs the J recursion — rare) */
      t0 = now_s();
      if (p7_b200_SetModelMSV(eng, om) != eslOK) p7_Fail("Failed to upload profile %s: %s\n", om->name, p7_b200_ErrorText());
      if (p7_b200_PrepareLengths(eng, qlens, nqlens, o.do_max ? 1.0 : o.F1) != eslOK) p7_Fail("Failed to prepare profile %s: %s\n", om->name, p7_b200_ErrorText());
      if (p7gpu_pipeline_run(p7_b200_Ctx(eng), p7_b200_Profile(eng), gdb, &o, &r) != P7GPU_OK) p7_Fail("GPU pipeline failed: %s\n", p7gpu_last_error());
      npass = p7gpu_result_nmsvpass(r);
      p7gpu_result_destroy(r);  r = NULL;
      t_msv += now_s() - t0;
      if (npass == 0) { p7_oprofile_Destroy(om); continue; }     /* nothing passed: the .h3p part of this model is never read */
    }

    /* stage 2: the rest of the model (p7_pipeline.c:708-713), then the rest of the pipeline for the survivors */
    t0 = now_s();
    if ((status = p7_oprofile_ReadRest(hfp, om)) != eslOK) p7_Fail("Failed to read the rest of profile %s: %s\n", om->name, hfp->errbuf);
    t_rest += now_s() - t0;  t0 = now_s();
    nfull++;
    if (o.do_biasfilter) p7_bg_SetFilter(bg, om->M, om->compo);  /* what p7_pli_NewModel does for the bias filter (p7_pipeline.c:500) */
    if (p7_b200_SetModel(eng, om, bg) != eslOK) p7_Fail("Failed to upload profile %s: %s\n", om->name, p7_b200_ErrorText());
    if (p7_b200_PrepareLengths(eng, qlens, nqlens, o.do_max ? 1.0 : o.F1) != eslOK) p7_Fail("Failed to prepare profile %s: %s\n", om->name, p7_b200_ErrorText());
    t_upload += now_s() - t0;  t0 = now_s();
    o.stop_after_msv = 0;  o.want_specials = 1;  o.stage_records = 1;
    o.fb_exact = fb_exact;  o.region_trigger = (fb_exact || o.do_max) ? 1 : 0;  o.decode = o.region_trigger;
    if (p7gpu_pipeline_run(p7_b200_Ctx(eng), p7_b200_Profile(eng), gdb, &o, &r) != P7GPU_OK) p7_Fail("GPU pipeline failed: %s\n", p7gpu_last_error());
    t_full += now_s() - t0;  t0 = now_s();
    {
      const p7gpu_msv_survivor *mv = p7gpu_result_msvpass(r);
      const p7gpu_survivor     *sv = p7gpu_result_survivors(r);
      const int64_t nm = p7gpu_result_nmsvpass(r), ns = p7gpu_result_nsurvivors(r);
      for (k = 0; k < nm; k++) {                                 /* the accounting of p7_Pipeline for every (query, model) that got past MSV */
        P7_PIPELINE *pli = qs[mv[k].seqidx].pli;
        pli->n_past_msv++;
        if (mv[k].stage >= 2) pli->n_past_bias++;
        if (mv[k].stage >= 3) pli->n_past_vit++;
        if (mv[k].stage >= 4) pli->n_past_fwd++;
      }
      if (ns > 0) {                                              /* Forward survivors: the reference's p7_Pipeline finishes them (replay) */
        const int nt = (int) ESL_MAX(1, ESL_MIN((int64_t) ncpus, ns / 4));
        int t;
        for (t = 0; t < nt; t++) {
          REPLAY_WORKER *W = &workers[t];
          W->t = t;  W->nt = nt;  W->sv = sv;  W->ns = ns;  W->r = r;  W->qs = qs;  W->nmodels = nmodels;
          W->om = (t == 0) ? om : p7_oprofile_Clone(om);
          if (!W->bg) W->bg = (t == 0) ? bg : p7_bg_Clone(bg);
          if (!W->om || !W->bg) p7_Fail("allocation failure\n");
          if (t > 0 && o.do_biasfilter) p7_bg_SetFilter(W->bg, om->M, om->compo);
        }
        for (t = 1; t < nt; t++) if (pthread_create(&tids[t], NULL, replay_worker, &workers[t]) != 0) p7_Fail("Failed to start a worker thread\n");
        replay_worker(&workers[0]);
        for (t = 1; t < nt; t++) pthread_join(tids[t], NULL);
        for (t = 0; t < nt; t++) {
          if (workers[t].status != eslOK) p7_Fail("p7_Pipeline failed (%d): %s\n", workers[t].status, workers[t].errbuf);
          if (t > 0) p7_oprofile_Destroy(workers[t].om);
        }
      }
    }
    p7gpu_result_destroy(r);
    p7_oprofile_Destroy(om);
    t_replay += now_s() - t0;
   }
   if (hstatus != eslOK) break;
  }
  if      (hstatus == eslEFORMAT)   p7_Fail("bad file format in HMM file %s", hmmfile);
  else if (hstatus == eslEINCOMPAT) p7_Fail("HMM file %s contains different alphabets", hmmfile);
  else if (hstatus != eslEOF)       p7_Fail("Unexpected error in reading HMMs from %s", hmmfile);
  esl_stopwatch_Stop(w);
  if (times) fprintf(stderr, "# %lld queries x %llu models: %.3f s; %llu models read in full (%.1f%%); %llu ended in the block prefilter (blocks of %d)\n", (long long) nq, (unsigned long long) nmodels,
                     w->elapsed, (unsigned long long) nfull, nmodels ? 100.0 * (double) nfull / (double) nmodels : 0.0, (unsigned long long) nskipped, blkmax > 1 ? blkmax : 0);
  if (times) fprintf(stderr, "# scan phases: read MSV parts %.3f s, block prefilter %.3f s, per-model MSV stage %.3f s, read rest %.3f s, full upload %.3f s, full pipeline %.3f s, host finish %.3f s\n",
                     t_read, t_pre, t_msv, t_rest, t_upload, t_full, t_replay);

  /* results per query, in file order (hmmscan.c:396-525) */
  for (q = 0; q < nq && qs; q++) {
    P7_PIPELINE *pli = qs[q].pli;  P7_TOPHITS *th = qs[q].th;  ESL_SQ *sq = qs[q].sq;
    pli->nmodels = nmodels;  pli->nnodes = nnodes;                                   /* p7_pli_NewModel, once per model */
    if (pli->Z_setby == p7_ZSETBY_NTARGETS) pli->Z = (double) nmodels;
    fprintf(ofp, "Query:       %s  [L=%ld]\n", sq->name, (long) sq->n);
    if (sq->acc[0]  != 0) fprintf(ofp, "Accession:   %s\n", sq->acc);
    if (sq->desc[0] != 0) fprintf(ofp, "Description: %s\n", sq->desc);
    p7_tophits_SortBySortkey(th);
    p7_tophits_Threshold(th, pli);
    p7_tophits_Targets(ofp, th, pli, textw);  fprintf(ofp, "\n\n");
    p7_tophits_Domains(ofp, th, pli, textw);  fprintf(ofp, "\n\n");
    if (tblfp)     p7_tophits_TabularTargets(tblfp,    sq->name, sq->acc, th, pli, (q == 0));
    if (domtblfp)  p7_tophits_TabularDomains(domtblfp, sq->name, sq->acc, th, pli, (q == 0));
    if (pfamtblfp) p7_tophits_TabularXfam(pfamtblfp,   sq->name, sq->acc, th, pli);
    p7_pli_Statistics(ofp, pli, w);
    fprintf(ofp, "//\n");
    p7_pipeline_Destroy(pli);  p7_tophits_Destroy(th);  esl_sq_Destroy(sq);
  }
  if (tblfp)     p7_tophits_TabularTail(tblfp,     "hmmscan", p7_SCAN_MODELS, seqfile, hmmfile, go);
  if (domtblfp)  p7_tophits_TabularTail(domtblfp,  "hmmscan", p7_SCAN_MODELS, seqfile, hmmfile, go);
  if (pfamtblfp) p7_tophits_TabularTail(pfamtblfp, "hmmscan", p7_SEARCH_SEQS, seqfile, hmmfile, go);
  fprintf(ofp, "[ok]\n");
  if (ofp != stdout) fclose(ofp); else fflush(stdout);
  if (tblfp) fclose(tblfp);
  if (domtblfp) fclose(domtblfp);
  if (pfamtblfp) fclose(pfamtblfp);
  fflush(stderr);
  _exit(0);                      /* outputs are on disk; the process exit releases the device (see hmmsearch_b200.c) */

 ERROR:
  return status;
}
