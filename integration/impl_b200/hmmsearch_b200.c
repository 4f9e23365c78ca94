/* hmmsearch_b200.c — hmmsearch with the per-target acceleration pipeline on B200 GPUs.
 *
 * Same command line, same output as the reference's hmmsearch (src/hmmsearch.c) — every table and alignment is printed
 * by the reference's own p7_tophits_* routines from a P7_TOPHITS filled by the reference's own p7_Pipeline() — but the
 * target database is digitised once by all host cores (p7gpu_dbfile_from_fasta) instead of one esl_sqio reader thread, and
 * the five DP stages of p7_Pipeline run stage-major on the GPU(s) (p7_b200_search.c).  Extra options: --gpus, --fbfast,
 * --p7db, --times.  Not supported: --restrictdb_* (SSI positioning of the text file), --mpi, stdin/gzip targets.
 */
#include "impl_b200.h"
#include "esl_getopts.h"
#include "esl_msa.h"
#include "esl_msafile.h"
#include "esl_stopwatch.h"
#include "esl_threads.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>
#include <time.h>

#define REPOPTS     "-E,-T,--cut_ga,--cut_nc,--cut_tc"
#define DOMREPOPTS  "--domE,--domT,--cut_ga,--cut_nc,--cut_tc"
#define INCOPTS     "--incE,--incT,--cut_ga,--cut_nc,--cut_tc"
#define INCDOMOPTS  "--incdomE,--incdomT,--cut_ga,--cut_nc,--cut_tc"
#define THRESHOPTS  "-E,-T,--domE,--domT,--incE,--incT,--incdomE,--incdomT,--cut_ga,--cut_nc,--cut_tc"

/* the option names, types and defaults p7_pipeline_Create() reads (p7_pipeline.c:66-91) are hmmsearch's own (hmmsearch.c:60-116) */
static ESL_OPTIONS options[] = {
  /* name           type            default  env           range     toggles reqs  incomp            help                                                    docgroup */
  { "-h",           eslARG_NONE,    FALSE,   NULL,         NULL,     NULL,   NULL, NULL,             "show brief help on version and usage",                         1 },
  { "-o",           eslARG_OUTFILE, NULL,    NULL,         NULL,     NULL,   NULL, NULL,             "direct output to file <f>, not stdout",                        2 },
  { "-A",           eslARG_OUTFILE, NULL,    NULL,         NULL,     NULL,   NULL, NULL,             "save multiple alignment of all hits to file <f>",              2 },
  { "--tblout",     eslARG_OUTFILE, NULL,    NULL,         NULL,     NULL,   NULL, NULL,             "save parseable table of per-sequence hits to file <f>",        2 },
  { "--domtblout",  eslARG_OUTFILE, NULL,    NULL,         NULL,     NULL,   NULL, NULL,             "save parseable table of per-domain hits to file <f>",          2 },
  { "--pfamtblout", eslARG_OUTFILE, NULL,    NULL,         NULL,     NULL,   NULL, NULL,             "save table of hits and domains to file, in Pfam format <f>",   2 },
  { "--acc",        eslARG_NONE,    FALSE,   NULL,         NULL,     NULL,   NULL, NULL,             "prefer accessions over names in output",                       2 },
  { "--noali",      eslARG_NONE,    FALSE,   NULL,         NULL,     NULL,   NULL, NULL,             "don't output alignments, so output is smaller",                2 },
  { "--notextw",    eslARG_NONE,    NULL,    NULL,         NULL,     NULL,   NULL, "--textw",        "unlimit ASCII text output line width",                         2 },
  { "--textw",      eslARG_INT,     "120",   NULL,         "n>=120", NULL,   NULL, "--notextw",      "set max width of ASCII text output lines",                     2 },
  { "-E",           eslARG_REAL,    "10.0",  NULL,         "x>0",    NULL,   NULL, REPOPTS,          "report sequences <= this E-value threshold in output",         4 },
  { "-T",           eslARG_REAL,    FALSE,   NULL,         NULL,     NULL,   NULL, REPOPTS,          "report sequences >= this score threshold in output",           4 },
  { "--domE",       eslARG_REAL,    "10.0",  NULL,         "x>0",    NULL,   NULL, DOMREPOPTS,       "report domains <= this E-value threshold in output",           4 },
  { "--domT",       eslARG_REAL,    FALSE,   NULL,         NULL,     NULL,   NULL, DOMREPOPTS,       "report domains >= this score cutoff in output",                4 },
  { "--incE",       eslARG_REAL,    "0.01",  NULL,         "x>0",    NULL,   NULL, INCOPTS,          "consider sequences <= this E-value threshold as significant",  5 },
  { "--incT",       eslARG_REAL,    FALSE,   NULL,         NULL,     NULL,   NULL, INCOPTS,          "consider sequences >= this score threshold as significant",    5 },
  { "--incdomE",    eslARG_REAL,    "0.01",  NULL,         "x>0",    NULL,   NULL, INCDOMOPTS,       "consider domains <= this E-value threshold as significant",    5 },
  { "--incdomT",    eslARG_REAL,    FALSE,   NULL,         NULL,     NULL,   NULL, INCDOMOPTS,       "consider domains >= this score threshold as significant",      5 },
  { "--cut_ga",     eslARG_NONE,    FALSE,   NULL,         NULL,     NULL,   NULL, THRESHOPTS,       "use profile's GA gathering cutoffs to set all thresholding",   6 },
  { "--cut_nc",     eslARG_NONE,    FALSE,   NULL,         NULL,     NULL,   NULL, THRESHOPTS,       "use profile's NC noise cutoffs to set all thresholding",       6 },
  { "--cut_tc",     eslARG_NONE,    FALSE,   NULL,         NULL,     NULL,   NULL, THRESHOPTS,       "use profile's TC trusted cutoffs to set all thresholding",     6 },
  { "--max",        eslARG_NONE,    FALSE,   NULL,         NULL,     NULL,   NULL, "--F1,--F2,--F3", "Turn all heuristic filters off (less speed, more power)",      7 },
  { "--F1",         eslARG_REAL,    "0.02",  NULL,         NULL,     NULL,   NULL, "--max",          "Stage 1 (MSV) threshold: promote hits w/ P <= F1",             7 },
  { "--F2",         eslARG_REAL,    "1e-3",  NULL,         NULL,     NULL,   NULL, "--max",          "Stage 2 (Vit) threshold: promote hits w/ P <= F2",             7 },
  { "--F3",         eslARG_REAL,    "1e-5",  NULL,         NULL,     NULL,   NULL, "--max",          "Stage 3 (Fwd) threshold: promote hits w/ P <= F3",             7 },
  { "--nobias",     eslARG_NONE,    NULL,    NULL,         NULL,     NULL,   NULL, "--max",          "turn off composition bias filter",                             7 },
  { "--nonull2",    eslARG_NONE,    NULL,    NULL,         NULL,     NULL,   NULL, NULL,             "turn off biased composition score corrections",               12 },
  { "-Z",           eslARG_REAL,    FALSE,   NULL,         "x>0",    NULL,   NULL, NULL,             "set # of comparisons done, for E-value calculation",          12 },
  { "--domZ",       eslARG_REAL,    FALSE,   NULL,         "x>0",    NULL,   NULL, NULL,             "set # of significant seqs, for domain E-value calculation",   12 },
  { "--seed",       eslARG_INT,     "42",    NULL,         "n>=0",   NULL,   NULL, NULL,             "set RNG seed to <n> (if 0: one-time arbitrary seed)",         12 },
  { "--tformat",    eslARG_STRING,  NULL,    NULL,         NULL,     NULL,   NULL, NULL,             "assert target <seqfile> is in format <s> (fasta only)",       12 },
  { "--cpu",        eslARG_INT,     NULL,    "HMMER_NCPU", "n>=0",   NULL,   NULL, NULL,             "number of host threads finishing the Forward survivors",      12 },
  /* B200 engine */
  { "--gpus",       eslARG_INT,     "1",     "HMMER_NGPU", "n>=1",   NULL,   NULL, NULL,             "number of GPUs; the database is sharded by sequence",         13 },
  { "--fbfast",     eslARG_NONE,    FALSE,   NULL,         NULL,     NULL,   NULL, NULL,             "Forward/Backward specials within 1e-4 instead of bit-exact",  13 },
  { "--hostdecode", eslARG_NONE,    FALSE,   NULL,         NULL,     NULL,   NULL, NULL,             "p7_DomainDecoding on the host from returned specials",        13 },
  { "--p7db",       eslARG_OUTFILE, NULL,    NULL,         NULL,     NULL,   NULL, NULL,             "keep the packed digital database in file <f> for later runs", 13 },
  { "--teardown",   eslARG_NONE,    FALSE,   NULL,         NULL,     NULL,   NULL, NULL,             "free every resource before exit (leak checkers)",             13 },
  { "--nopin",      eslARG_NONE,    FALSE,   NULL,         NULL,     NULL,   NULL, NULL,             "never page-lock the database (default: from the 2nd query on)", 13 },
  { "--pinfirst",   eslARG_NONE,    FALSE,   NULL,         NULL,     NULL,   NULL, "--nopin",        "page-lock the database before the first query",               13 },
  { "--times",      eslARG_NONE,    FALSE,   NULL,         NULL,     NULL,   NULL, NULL,             "report wall time of each phase on stderr",                    13 },
  {  0, 0, 0, 0, 0, 0, 0, 0, 0, 0 },
};

static double
since_start(void)
{
  static struct timespec t0; static int have = 0;
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  if (!have) { t0 = t; have = 1; }
  return (double) (t.tv_sec - t0.tv_sec) + 1e-9 * (double) (t.tv_nsec - t0.tv_nsec);
}

static char usage[]  = "[options] <hmmfile> <seqdb>";
static char banner[] = "search profile(s) against a sequence database";

static void
help_and_exit(ESL_GETOPTS *go, char *argv0)
{
  static const struct { int group; const char *title; } sect[] = {
    { 1, "Basic options:" }, { 2, "Options directing output:" }, { 4, "Options controlling reporting thresholds:" },
    { 5, "Options controlling inclusion (significance) thresholds:" }, { 6, "Options controlling model-specific thresholding:" },
    { 7, "Options controlling acceleration heuristics:" }, { 12, "Other expert options:" }, { 13, "Options of the B200 engine:" } };
  size_t i;
  p7_banner(stdout, argv0, banner);
  esl_usage(stdout, argv0, usage);
  for (i = 0; i < sizeof(sect) / sizeof(sect[0]); i++) { printf("\n%s\n", sect[i].title); esl_opt_DisplayHelp(stdout, go, sect[i].group, 2, 80); }
  exit(0);
}

/* "# option:   value" lines exactly as hmmsearch prints them (hmmsearch.c:215-268), table-driven */
static void
echo_options(FILE *ofp, const ESL_GETOPTS *go, char *hmmfile, char *seqfile)
{
  static const struct { const char *opt; const char *fmt; char kind; } line[] = {   /* kind: s string, d int, g real, f real %.0f, - none */
    { "-o",           "# output directed to file:         %s\n", 's' },  { "-A",          "# MSA of all hits saved to file:   %s\n", 's' },
    { "--tblout",     "# per-seq hits tabular output:     %s\n", 's' },  { "--domtblout", "# per-dom hits tabular output:     %s\n", 's' },
    { "--pfamtblout", "# pfam-style tabular hit output:   %s\n", 's' },  { "--acc",       "# prefer accessions over names:    yes\n", '-' },
    { "--noali",      "# show alignments in output:       no\n", '-' },  { "--notextw",   "# max ASCII text line length:      unlimited\n", '-' },
    { "--textw",      "# max ASCII text line length:      %d\n", 'd' },  { "-E",          "# sequence reporting threshold:    E-value <= %g\n", 'g' },
    { "-T",           "# sequence reporting threshold:    score >= %g\n", 'g' },   { "--domE",    "# domain reporting threshold:      E-value <= %g\n", 'g' },
    { "--domT",       "# domain reporting threshold:      score >= %g\n", 'g' },   { "--incE",    "# sequence inclusion threshold:    E-value <= %g\n", 'g' },
    { "--incT",       "# sequence inclusion threshold:    score >= %g\n", 'g' },   { "--incdomE", "# domain inclusion threshold:      E-value <= %g\n", 'g' },
    { "--incdomT",    "# domain inclusion threshold:      score >= %g\n", 'g' },   { "--cut_ga",  "# model-specific thresholding:     GA cutoffs\n", '-' },
    { "--cut_nc",     "# model-specific thresholding:     NC cutoffs\n", '-' },    { "--cut_tc",  "# model-specific thresholding:     TC cutoffs\n", '-' },
    { "--max",        "# Max sensitivity mode:            on [all heuristic filters off]\n", '-' },
    { "--F1",         "# MSV filter P threshold:       <= %g\n", 'g' },  { "--F2",        "# Vit filter P threshold:       <= %g\n", 'g' },
    { "--F3",         "# Fwd filter P threshold:       <= %g\n", 'g' },  { "--nobias",    "# biased composition HMM filter:   off\n", '-' },
    { "--nonull2",    "# null2 bias corrections:          off\n", '-' }, { "-Z",          "# sequence search space set to:    %.0f\n", 'f' },
    { "--domZ",       "# domain search space set to:      %.0f\n", 'f' } };
  size_t i;
  p7_banner(ofp, go->argv[0], banner);
  fprintf(ofp, "# query HMM file:                  %s\n", hmmfile);
  fprintf(ofp, "# target sequence database:        %s\n", seqfile);
  for (i = 0; i < sizeof(line) / sizeof(line[0]); i++) {
    if (!esl_opt_IsUsed(go, (char *) line[i].opt)) continue;
    switch (line[i].kind) {
    case 's': fprintf(ofp, line[i].fmt, esl_opt_GetString(go, (char *) line[i].opt));  break;
    case 'd': fprintf(ofp, line[i].fmt, esl_opt_GetInteger(go, (char *) line[i].opt)); break;
    case 'g': case 'f': fprintf(ofp, line[i].fmt, esl_opt_GetReal(go, (char *) line[i].opt)); break;
    default:  fputs(line[i].fmt, ofp); break;
    }
  }
  if (esl_opt_IsUsed(go, "--seed")) {
    if (esl_opt_GetInteger(go, "--seed") == 0) fprintf(ofp, "# random number seed:              one-time arbitrary\n");
    else                                       fprintf(ofp, "# random number seed set to:       %d\n", esl_opt_GetInteger(go, "--seed"));
  }
  if (esl_opt_IsUsed(go, "--tformat")) fprintf(ofp, "# targ <seqfile> format asserted:  %s\n", esl_opt_GetString(go, "--tformat"));
  if (esl_opt_IsUsed(go, "--cpu"))     fprintf(ofp, "# number of worker threads:        %d\n", esl_opt_GetInteger(go, "--cpu"));
  if (esl_opt_IsUsed(go, "--gpus"))    fprintf(ofp, "# number of GPUs:                  %d\n", esl_opt_GetInteger(go, "--gpus"));
  if (esl_opt_IsUsed(go, "--fbfast"))  fprintf(ofp, "# Forward/Backward specials:       fast (1e-4 relative)\n");
  if (esl_opt_IsUsed(go, "--p7db"))    fprintf(ofp, "# packed database kept in:         %s\n", esl_opt_GetString(go, "--p7db"));
  fprintf(ofp, "# - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - -\n\n");
}

static FILE *
open_out(ESL_GETOPTS *go, char *opt, const char *what)
{
  FILE *fp;
  if (!esl_opt_IsOn(go, opt)) return NULL;
  if ((fp = fopen(esl_opt_GetString(go, opt), "w")) == NULL) p7_Fail("Failed to open %s %s for writing\n", what, esl_opt_GetString(go, opt));
  return fp;
}

/* <seqdb> is either a packed database written earlier (--p7db) or FASTA text, digitised now with all host threads.  None of
 * this touches the device, so it runs on its own thread while the main thread waits for the CUDA driver to initialise
 * (0.5-1 s); failures are kept and reported by the main thread at the point where hmmsearch opens its sequence file. */
typedef struct {
  ESL_GETOPTS  *go;
  char         *seqfile;
  p7gpu_dbfile *db;
  double        seconds, done_at;
  char          fail[1024];        /* "" = ok */
} DB_PREP;

static void *
prepare_database(void *arg)
{
  DB_PREP *pr = (DB_PREP *) arg;
  ESL_GETOPTS *go = pr->go;
  char magic[8] = { 0 }, tmp[4096];
  FILE *fp = fopen(pr->seqfile, "rb");
  ESL_STOPWATCH *w = esl_stopwatch_Create();
  int keep = esl_opt_IsOn(go, "--p7db");
  pr->db = NULL;  pr->fail[0] = '\0';
  if (!fp) { snprintf(pr->fail, sizeof(pr->fail), "Failed to open sequence file %s for reading\n", pr->seqfile); goto DONE; }
  if (fread(magic, 1, 8, fp) == 0) { fclose(fp); snprintf(pr->fail, sizeof(pr->fail), "Sequence file %s is empty or misformatted\n", pr->seqfile); goto DONE; }
  fclose(fp);
  esl_stopwatch_Start(w);
  if (memcmp(magic, "P7GPUDB1", 8) == 0) {
    if (p7gpu_dbfile_open(pr->seqfile, &pr->db) != P7GPU_OK) { snprintf(pr->fail, sizeof(pr->fail), "%s\n", p7gpu_last_error()); goto DONE; }
  } else {
    int64_t nseq = 0, nres = 0;
    const char *dir = getenv("TMPDIR");
    if (esl_opt_IsOn(go, "--tformat") && strcasecmp(esl_opt_GetString(go, "--tformat"), "fasta") != 0) {
      snprintf(pr->fail, sizeof(pr->fail), "%s is not a sequence database file format this engine reads (fasta)\n", esl_opt_GetString(go, "--tformat"));
      goto DONE;
    }
    if (keep) snprintf(tmp, sizeof(tmp), "%s", esl_opt_GetString(go, "--p7db"));
    else      snprintf(tmp, sizeof(tmp), "%s/hmmsearch_b200.%ld.p7db", dir ? dir : "/tmp", (long) getpid());
    if (p7gpu_dbfile_from_fasta(pr->seqfile, tmp, 0, 16, &nseq, &nres) != P7GPU_OK) {                /* hmmsearch.c:523-524 */
      snprintf(pr->fail, sizeof(pr->fail), "Parse failed (sequence file %s):\n%s\n", pr->seqfile, p7gpu_last_error());
      goto DONE;
    }
    if (p7gpu_dbfile_open(tmp, &pr->db) != P7GPU_OK) { snprintf(pr->fail, sizeof(pr->fail), "%s\n", p7gpu_last_error()); goto DONE; }
    if (!keep) unlink(tmp);                                                                     /* the mapping keeps it alive */
  }
  esl_stopwatch_Stop(w);
  pr->seconds = w->elapsed;
 DONE:
  pr->done_at = since_start();
  esl_stopwatch_Destroy(w);
  return NULL;
}

/* CUDA start-up (driver initialisation, then one context per GPU, each on its own thread) also runs beside the main thread,
 * which meanwhile reads the first profile; failures are reported by the main thread in hmmsearch's order. */
typedef struct {
  int       ngpus;
  P7_B200 **engs;
  double    seconds, done_at;
  char      fail[1024];        /* "" = ok */
} GPU_OPEN;

typedef struct { int device; P7_B200 **slot; char fail[768]; } GPU_OPEN_ONE;

static void *
open_one_engine(void *arg)
{
  GPU_OPEN_ONE *o = (GPU_OPEN_ONE *) arg;
  o->fail[0] = '\0';
  if (p7_b200_Open(o->device, o->slot) != eslOK) snprintf(o->fail, sizeof(o->fail), "Failed to open GPU %d: %s\n", o->device, p7_b200_ErrorText());
  return NULL;
}

/* first step of the start-up: the driver itself (cuInit behind the device count), 0.5-0.7 s */
static void *
count_devices(void *arg)
{
  GPU_OPEN *g = (GPU_OPEN *) arg;
  int nvis = p7_b200_DeviceCount();
  g->fail[0] = '\0';
  if (g->ngpus > nvis) snprintf(g->fail, sizeof(g->fail), "--gpus %d, but %d CUDA device(s) visible\n%s\n", g->ngpus, nvis, p7_b200_ErrorText());
  g->done_at = since_start();
  return NULL;
}

static void *
open_engines(void *arg)
{
  GPU_OPEN *g = (GPU_OPEN *) arg;
  GPU_OPEN_ONE *one = calloc((size_t) g->ngpus, sizeof(GPU_OPEN_ONE));
  pthread_t    *tid = calloc((size_t) g->ngpus, sizeof(pthread_t));
  double t0 = since_start();
  int k;
  g->fail[0] = '\0';
  if (!one || !tid) { snprintf(g->fail, sizeof(g->fail), "allocation failure\n"); goto DONE; }
  for (k = 0; k < g->ngpus; k++) { one[k].device = k; one[k].slot = &g->engs[k]; }
  for (k = 1; k < g->ngpus; k++) if (pthread_create(&tid[k], NULL, open_one_engine, &one[k]) != 0) { open_one_engine(&one[k]); tid[k] = 0; }
  open_one_engine(&one[0]);
  for (k = 1; k < g->ngpus; k++) if (tid[k]) pthread_join(tid[k], NULL);
  for (k = 0; k < g->ngpus; k++) if (one[k].fail[0]) { snprintf(g->fail, sizeof(g->fail), "%s", one[k].fail); break; }
 DONE:
  g->done_at = since_start();  g->seconds = g->done_at - t0;
  free(one);  free(tid);
  return NULL;
}

int
main(int argc, char **argv)
{
  ESL_GETOPTS   *go    = esl_getopts_Create(options);
  char          *hmmfile, *seqfile;
  FILE          *ofp = stdout, *afp, *tblfp, *domtblfp, *pfamtblfp;
  P7_HMMFILE    *hfp   = NULL;
  P7_HMM        *hmm   = NULL;
  ESL_ALPHABET  *abc   = NULL;
  P7_BG         *bg    = NULL;
  P7_B200      **engs  = NULL;
  p7gpu_dbfile  *db    = NULL;
  ESL_STOPWATCH *w     = esl_stopwatch_Create();
  P7_B200_SEARCH cfg;
  char           errbuf[eslERRBUFSIZE];
  DB_PREP        prep;
  GPU_OPEN       gopen;
  pthread_t      prep_tid, gopen_tid;
  int            textw, nquery = 0, hstatus, status, k, ncpus = 0, times;

  since_start();
  impl_Init();                 /* FTZ/DAZ for the host code that stays (impl_sse.h:356-376) */
  p7_FLogsumInit();

  if (esl_opt_ProcessEnvironment(go) != eslOK || esl_opt_ProcessCmdline(go, argc, argv) != eslOK || esl_opt_VerifyConfig(go) != eslOK) {
    printf("Failed to parse command line: %s\n", go->errbuf);
    esl_usage(stdout, argv[0], usage);
    printf("\nTo see more help on available options, do %s -h\n\n", argv[0]);
    exit(1);
  }
  if (esl_opt_GetBoolean(go, "-h")) help_and_exit(go, argv[0]);
  if (esl_opt_ArgNumber(go) != 2) { puts("Incorrect number of command line arguments."); esl_usage(stdout, argv[0], usage); exit(1); }
  hmmfile = esl_opt_GetArg(go, 1);
  seqfile = esl_opt_GetArg(go, 2);
  if (strcmp(seqfile, "-") == 0) p7_Fail("This engine maps the target database; it cannot be read from stdin\n");
  textw = esl_opt_GetBoolean(go, "--notextw") ? 0 : esl_opt_GetInteger(go, "--textw");
  times = esl_opt_GetBoolean(go, "--times");

  status = p7_hmmfile_OpenE(hmmfile, NULL, &hfp, errbuf);
  if      (status == eslENOTFOUND) p7_Fail("File existence/permissions problem in trying to open HMM file %s.\n%s\n", hmmfile, errbuf);
  else if (status == eslEFORMAT)   p7_Fail("File format problem in trying to open HMM file %s.\n%s\n",                hmmfile, errbuf);
  else if (status != eslOK)        p7_Fail("Unexpected error %d in opening HMM file %s.\n%s\n",               status, hmmfile, errbuf);

  if (esl_opt_IsOn(go, "-o") && (ofp = fopen(esl_opt_GetString(go, "-o"), "w")) == NULL) p7_Fail("Failed to open output file %s for writing\n", esl_opt_GetString(go, "-o"));
  afp       = open_out(go, "-A",           "alignment file");
  tblfp     = open_out(go, "--tblout",     "tabular per-seq output file");
  domtblfp  = open_out(go, "--domtblout",  "tabular per-dom output file");
  pfamtblfp = open_out(go, "--pfamtblout", "pfam-style tabular output file");

  if (esl_opt_IsOn(go, "--cpu")) ncpus = esl_opt_GetInteger(go, "--cpu");
  else                           esl_threads_CPUCount(&ncpus);
  memset(&cfg, 0, sizeof(cfg));
  cfg.ngpus    = esl_opt_GetInteger(go, "--gpus");
  cfg.ncpus    = ESL_MAX(1, ncpus);
  cfg.fb_exact = esl_opt_GetBoolean(go, "--fbfast") ? 0 : 1;
  cfg.decode   = esl_opt_GetBoolean(go, "--hostdecode") ? 0 : 1;
  /* the driver initialises every device it can see (~0.1 s each on an 8-GPU board): show it only the ones this run uses */
  if (getenv("CUDA_VISIBLE_DEVICES") == NULL) {
    char vis[256]; int n = 0;
    for (k = 0; k < cfg.ngpus && n < 240; k++) n += snprintf(vis + n, sizeof(vis) - n, k ? ",%d" : "%d", k);
    setenv("CUDA_VISIBLE_DEVICES", vis, 0);
  }
  ESL_ALLOC(engs, sizeof(*engs) * cfg.ngpus);
  for (k = 0; k < cfg.ngpus; k++) engs[k] = NULL;
  prep.go = go;  prep.seqfile = seqfile;  prep.db = NULL;  prep.seconds = 0.0;  prep.fail[0] = '\0';
  gopen.ngpus = cfg.ngpus;  gopen.engs = engs;  gopen.fail[0] = '\0';
  if (pthread_create(&prep_tid,  NULL, prepare_database, &prep)  != 0) p7_Fail("Failed to start the database thread\n");
  if (pthread_create(&gopen_tid, NULL, count_devices,    &gopen) != 0) p7_Fail("Failed to start the GPU thread\n");
  pthread_join(gopen_tid, NULL);                           /* the database thread works through this wait */
  if (gopen.fail[0]) p7_Fail("%s", gopen.fail);
  if (times) fprintf(stderr, "# CUDA driver initialised: %.3f s since start\n", since_start());
  if (pthread_create(&gopen_tid, NULL, open_engines,     &gopen) != 0) p7_Fail("Failed to start the GPU thread\n");

  /* The first profile, while the contexts are created.  A profile file without E-value parameters (HMMER2 format, as
   * biobench2's tufa420.hmm) is calibrated here by the reference's p7_Calibrate, whose ~600 one-target filter calls go
   * through the drop-in symbols of this library (impl_b200.h) on the registry's own engine. */
  hstatus = p7_hmmfile_Read(hfp, &abc, &hmm);
  if (times) fprintf(stderr, "# first profile read: %.3f s since start\n", since_start());
  pthread_join(gopen_tid, NULL);
  if (gopen.fail[0]) p7_Fail("%s", gopen.fail);
  if (times) fprintf(stderr, "# %d GPU engine(s) open: contexts took %.3f s on their own thread(s), done %.3f s after start; %.3f s since start\n",
                     cfg.ngpus, gopen.seconds, gopen.done_at, since_start());
  if (hstatus == eslOK) {
    if (abc->type != eslAMINO) p7_Fail("The B200 engine searches protein databases; %s is not an amino acid profile\n", hmmfile);
    echo_options(ofp, go, hmmfile, seqfile);
    bg = p7_bg_Create(abc);
    pthread_join(prep_tid, NULL);                            /* once, not once per query (cf. hmmsearch.c:480-487) */
    if (prep.fail[0]) p7_Fail("%s", prep.fail);
    db = prep.db;
    if (times) fprintf(stderr, "# database ready: digitised in %.3f s on its own thread, done %.3f s after start (%lld targets, %lld residues); %.3f s since start\n",
                       prep.seconds, prep.done_at, (long long) p7gpu_dbfile_nseq(db), (long long) p7gpu_dbfile_nres(db), since_start());
  } else pthread_join(prep_tid, NULL);

  while (hstatus == eslOK) {
    P7_PROFILE  *gm  = p7_profile_Create (hmm->M, abc);
    P7_OPROFILE *om  = p7_oprofile_Create(hmm->M, abc);
    P7_PIPELINE *pli;
    P7_TOPHITS  *th  = p7_tophits_Create();

    nquery++;
    /* The first query streams the database from the file mapping; page-locking it (and building the 5-bit transfer form)
     * costs more than one pass saves, so that is done when a second query shows it will be read again. */
    if (nquery == (esl_opt_GetBoolean(go, "--pinfirst") ? 1 : 2) && !esl_opt_GetBoolean(go, "--nopin")) {
      esl_stopwatch_Start(w);
      if (p7gpu_dbfile_pin(db) != P7GPU_OK) p7_Fail("%s\n", p7gpu_last_error());
      esl_stopwatch_Stop(w);
      if (times) fprintf(stderr, "# database page-locked for the remaining queries: %.3f s\n", w->elapsed);
    }
    esl_stopwatch_Start(w);
    fprintf(ofp, "Query:       %s  [M=%d]\n", hmm->name, hmm->M);
    if (hmm->acc)  fprintf(ofp, "Accession:   %s\n", hmm->acc);
    if (hmm->desc) fprintf(ofp, "Description: %s\n", hmm->desc);

    p7_ProfileConfig(hmm, bg, gm, 100, p7_LOCAL);            /* hmmsearch.c:500-503 */
    p7_oprofile_Convert(gm, om);
    pli = p7_pipeline_Create(go, om->M, 100, FALSE, p7_SEARCH_SEQS);
    if ((status = p7_pli_NewModel(pli, om, bg)) != eslOK) p7_Fail("%s\n", pli->errbuf);
    for (k = 0; k < cfg.ngpus; k++)
      if (p7_b200_SetModel(engs[k], om, bg) != eslOK) p7_Fail("Failed to upload the profile to GPU %d: %s\n", k, p7_b200_ErrorText());

    if ((status = p7_b200_SearchDatabase(engs, &cfg, go, abc, om, bg, pli, th, db)) != eslOK)
      p7_Fail("Search failed (%d): %s\n%s\n", status, pli->errbuf, p7_b200_ErrorText());

    p7_tophits_SortBySortkey(th);                            /* hmmsearch.c:548-556 */
    p7_tophits_Threshold(th, pli);
    p7_tophits_Targets(ofp, th, pli, textw);  fprintf(ofp, "\n\n");
    p7_tophits_Domains(ofp, th, pli, textw);  fprintf(ofp, "\n\n");
    if (tblfp)     p7_tophits_TabularTargets(tblfp,    hmm->name, hmm->acc, th, pli, (nquery == 1));
    if (domtblfp)  p7_tophits_TabularDomains(domtblfp, hmm->name, hmm->acc, th, pli, (nquery == 1));
    if (pfamtblfp) p7_tophits_TabularXfam(pfamtblfp,   hmm->name, hmm->acc, th, pli);
    esl_stopwatch_Stop(w);
    p7_pli_Statistics(ofp, pli, w);
    fprintf(ofp, "//\n");
    if (times) fprintf(stderr, "# query %d (%s): %.3f s elapsed; GPU pipeline %.3f s, %llu Forward survivors finished on %d host threads in %.3f s\n",
                       nquery, hmm->name, w->elapsed, cfg.gpu_seconds, (unsigned long long) cfg.nsurvivors, cfg.ncpus, cfg.post_seconds);

    if (afp) {
      ESL_MSA *msa = NULL;
      if (p7_tophits_Alignment(th, abc, NULL, NULL, 0, p7_ALL_CONSENSUS_COLS, &msa) == eslOK) {
        eslx_msafile_Write(afp, msa, textw > 0 ? eslMSAFILE_STOCKHOLM : eslMSAFILE_PFAM);
        fprintf(ofp, "# Alignment of %d hits satisfying inclusion thresholds saved to: %s\n", msa->nseq, esl_opt_GetString(go, "-A"));
      } else fprintf(ofp, "# No hits satisfy inclusion thresholds; no alignment saved\n");
      esl_msa_Destroy(msa);
    }

    p7_pipeline_Destroy(pli);  p7_tophits_Destroy(th);  p7_oprofile_Destroy(om);  p7_profile_Destroy(gm);  p7_hmm_Destroy(hmm);
    hstatus = p7_hmmfile_Read(hfp, &abc, &hmm);
  }
  switch (hstatus) {
  case eslEOD:       p7_Fail("read failed, HMM file %s may be truncated?", hmmfile);              break;
  case eslEFORMAT:   p7_Fail("bad file format in HMM file %s",             hmmfile);              break;
  case eslEINCOMPAT: p7_Fail("HMM file %s contains different alphabets",   hmmfile);              break;
  case eslEOF:       break;
  default:           p7_Fail("Unexpected error (%d) in reading HMMs from %s", hstatus, hmmfile);
  }

  if (tblfp)     p7_tophits_TabularTail(tblfp,     "hmmsearch", p7_SEARCH_SEQS, hmmfile, seqfile, go);
  if (domtblfp)  p7_tophits_TabularTail(domtblfp,  "hmmsearch", p7_SEARCH_SEQS, hmmfile, seqfile, go);
  if (pfamtblfp) p7_tophits_TabularTail(pfamtblfp, "hmmsearch", p7_SEARCH_SEQS, hmmfile, seqfile, go);
  fprintf(ofp, "[ok]\n");

  if (ofp != stdout) fclose(ofp); else fflush(stdout);
  if (afp)       fclose(afp);
  if (tblfp)     fclose(tblfp);
  if (domtblfp)  fclose(domtblfp);
  if (pfamtblfp) fclose(pfamtblfp);
  if (times) fprintf(stderr, "# results written: %.3f s since start\n", since_start());
  if (!esl_opt_GetBoolean(go, "--teardown")) {
    /* every output is on disk; unmapping the database, freeing page-locked memory and destroying the CUDA contexts in an
     * orderly way costs about a second that the process exit does for free */
    fflush(stderr);
    _exit(0);
  }
  for (k = 0; k < cfg.ngpus; k++) p7_b200_Close(engs[k]);
  free(engs);
  if (db) p7gpu_dbfile_close(db);
  p7_bg_Destroy(bg);  p7_hmmfile_Close(hfp);  esl_alphabet_Destroy(abc);  esl_stopwatch_Destroy(w);
  esl_getopts_Destroy(go);
  if (times) fprintf(stderr, "# teardown done: %.3f s since start\n", since_start());
  return eslOK;

 ERROR:
  return status;
}
