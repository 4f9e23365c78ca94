/* impl_b200_utest.c — the five drop-in symbols against the reference's own, in one process.
 *
 * integration/Makefile keeps the reference's SSE implementations in the binary under p7_*_sse names; the unsuffixed
 * symbols are impl_b200's.  For a profile (an HMM file, or p7_hmm_Sample of -M nodes) and a set of targets (iid random
 * sequences plus homologs emitted from the profile, lengths around -L) this driver calls both engines with identical
 * arguments — the shape of the reference's own unit tests (msvfilter.c:708-726, fwdback.c utest) — and requires
 *     p7_SSVFilter, p7_MSVFilter, p7_ViterbiFilter : same return code, same score bits;
 *     p7_ForwardParser, p7_BackwardParser          : same return code, same score bits, every special of every row
 *                                                    (ox->xmx), totscale and has_own_scales identical bit for bit.
 * Exit status 0 = all identical.
 */
#include "impl_b200.h"
#include "esl_getopts.h"
#include "esl_random.h"
#include "esl_randomseq.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static ESL_OPTIONS options[] = {
  { "-h", eslARG_NONE,  FALSE, NULL, NULL, NULL, NULL, NULL, "show help",                                      0 },
  { "-N", eslARG_INT,   "200", NULL, "n>0", NULL, NULL, NULL, "number of random targets",                       0 },
  { "-H", eslARG_INT,    "20", NULL, "n>=0", NULL, NULL, NULL, "number of homologs emitted from the profile",    0 },
  { "-L", eslARG_INT,   "300", NULL, "n>0", NULL, NULL, NULL, "mean target length",                             0 },
  { "-M", eslARG_INT,   "145", NULL, "n>0", NULL, NULL, NULL, "nodes of the sampled profile (no <hmmfile>)",    0 },
  { "-s", eslARG_INT,    "42", NULL, "n>=0", NULL, NULL, NULL, "random seed",                                    0 },
  { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 },
};

static int same_bits(float a, float b) { uint32_t x, y; memcpy(&x, &a, 4); memcpy(&y, &b, 4); return x == y; }

int
main(int argc, char **argv)
{
  ESL_GETOPTS    *go  = esl_getopts_Create(options);
  ESL_RANDOMNESS *r;
  ESL_ALPHABET   *abc = NULL;
  P7_HMM         *hmm = NULL;
  P7_BG          *bg;
  P7_PROFILE     *gm;
  P7_OPROFILE    *om;
  P7_OMX         *fa, *fb, *ba, *bb;
  ESL_SQ         *sq;
  ESL_DSQ        *dsq;
  int N, H, Lmean, t, nbad = 0, nfail[5] = { 0, 0, 0, 0, 0 }, nresc = 0;
  static const char *what[5] = { "p7_SSVFilter", "p7_MSVFilter", "p7_ViterbiFilter", "p7_ForwardParser", "p7_BackwardParser" };

  if (esl_opt_ProcessCmdline(go, argc, argv) != eslOK || esl_opt_VerifyConfig(go) != eslOK) esl_fatal("%s", go->errbuf);
  if (esl_opt_GetBoolean(go, "-h")) { printf("Usage: %s [options] [<hmmfile>]\n", argv[0]); esl_opt_DisplayHelp(stdout, go, 0, 2, 80); return 0; }
  impl_Init();  p7_FLogsumInit();
  r = esl_randomness_CreateFast(esl_opt_GetInteger(go, "-s"));
  N = esl_opt_GetInteger(go, "-N");  H = esl_opt_GetInteger(go, "-H");  Lmean = esl_opt_GetInteger(go, "-L");

  if (esl_opt_ArgNumber(go) >= 1) {
    P7_HMMFILE *hfp = NULL; char errbuf[eslERRBUFSIZE];
    if (p7_hmmfile_OpenE(esl_opt_GetArg(go, 1), NULL, &hfp, errbuf) != eslOK) esl_fatal("%s", errbuf);
    if (p7_hmmfile_Read(hfp, &abc, &hmm) != eslOK) esl_fatal("failed to read an HMM from %s", esl_opt_GetArg(go, 1));
    p7_hmmfile_Close(hfp);
  } else {
    abc = esl_alphabet_Create(eslAMINO);
    p7_hmm_Sample(r, esl_opt_GetInteger(go, "-M"), abc, &hmm);
  }
  bg = p7_bg_Create(abc);
  gm = p7_profile_Create(hmm->M, abc);
  om = p7_oprofile_Create(hmm->M, abc);
  p7_ProfileConfig(hmm, bg, gm, Lmean, p7_LOCAL);
  p7_oprofile_Convert(gm, om);
  fa = p7_omx_Create(hmm->M, 0, Lmean);  fb = p7_omx_Create(hmm->M, 0, Lmean);
  ba = p7_omx_Create(hmm->M, 0, Lmean);  bb = p7_omx_Create(hmm->M, 0, Lmean);
  sq = esl_sq_CreateDigital(abc);

  for (t = 0; t < N + H; t++) {
    int L, st_a, st_b, i, c, bad;
    float sa, sb;
    if (t < N) {                                            /* iid target, length Lmean/2 .. 3*Lmean/2 (every 17th: 1..16) */
      L = (t % 17 == 16) ? 1 + (int) (esl_random(r) * 16) : Lmean / 2 + (int) (esl_random(r) * Lmean);
      dsq = malloc(L + 2);
      esl_rsq_xfIID(r, bg->f, abc->K, L, dsq);
    } else {                                                /* homolog: rescaling rows, high scores */
      p7_ProfileEmit(r, hmm, gm, bg, sq, NULL);
      L = (int) sq->n;
      if (L < 1) { esl_sq_Reuse(sq); continue; }
      dsq = malloc(L + 2); memcpy(dsq, sq->dsq, L + 2);
      esl_sq_Reuse(sq);
    }
    p7_oprofile_ReconfigLength(om, L);                      /* the reference reads L-dependent scalars from <om>; impl_b200 derives them from L */
    p7_omx_GrowTo(fa, om->M, 0, L);  p7_omx_GrowTo(fb, om->M, 0, L);  p7_omx_GrowTo(ba, om->M, 0, L);  p7_omx_GrowTo(bb, om->M, 0, L);

    sa = sb = 0.0f;
    st_a = p7_SSVFilter_sse(dsq, L, om, &sa);  st_b = p7_SSVFilter(dsq, L, om, &sb);
    bad = (st_a != st_b) || (st_a != eslENORESULT && !same_bits(sa, sb));
    if (bad) { nfail[0]++; if (nbad++ < 10) printf("target %d (L=%d) %s: sse status %d sc %g | b200 status %d sc %g\n", t, L, what[0], st_a, sa, st_b, sb); }

    st_a = p7_MSVFilter_sse(dsq, L, om, fa, &sa);  st_b = p7_MSVFilter(dsq, L, om, fb, &sb);
    if (st_a != st_b || !same_bits(sa, sb)) { nfail[1]++; if (nbad++ < 10) printf("target %d (L=%d) %s: sse status %d sc %g | b200 status %d sc %g\n", t, L, what[1], st_a, sa, st_b, sb); }

    st_a = p7_ViterbiFilter_sse(dsq, L, om, fa, &sa);  st_b = p7_ViterbiFilter(dsq, L, om, fb, &sb);
    if (st_a != st_b || !same_bits(sa, sb)) { nfail[2]++; if (nbad++ < 10) printf("target %d (L=%d) %s: sse status %d sc %g | b200 status %d sc %g\n", t, L, what[2], st_a, sa, st_b, sb); }

    st_a = p7_ForwardParser_sse(dsq, L, om, fa, &sa);  st_b = p7_ForwardParser(dsq, L, om, fb, &sb);
    bad = (st_a != st_b) || !same_bits(sa, sb) || !same_bits(fa->totscale, fb->totscale) || fa->has_own_scales != fb->has_own_scales;
    for (i = 0; i <= L && !bad; i++) for (c = 0; c < p7X_NXCELLS; c++) if (!same_bits(fa->xmx[i*p7X_NXCELLS+c], fb->xmx[i*p7X_NXCELLS+c])) { bad = 1; break; }
    for (i = 1; i <= L; i++) if (fa->xmx[i*p7X_NXCELLS+p7X_SCALE] > 1.0f) nresc++;
    if (bad) { nfail[3]++; if (nbad++ < 10) printf("target %d (L=%d) %s: sse status %d sc %.9g | b200 status %d sc %.9g\n", t, L, what[3], st_a, sa, st_b, sb); }

    /* both Backward passes get the reference's Forward matrix, so a Forward difference cannot hide or fake a Backward one */
    st_a = p7_BackwardParser_sse(dsq, L, om, fa, ba, &sa);  st_b = p7_BackwardParser(dsq, L, om, fa, bb, &sb);
    bad = (st_a != st_b) || !same_bits(sa, sb) || !same_bits(ba->totscale, bb->totscale) || ba->has_own_scales != bb->has_own_scales;
    for (i = 0; i <= L && !bad; i++) for (c = 0; c < p7X_NXCELLS; c++) if (!same_bits(ba->xmx[i*p7X_NXCELLS+c], bb->xmx[i*p7X_NXCELLS+c])) { bad = 1; break; }
    if (bad) { nfail[4]++; if (nbad++ < 10) printf("target %d (L=%d) %s: sse status %d sc %.9g | b200 status %d sc %.9g\n", t, L, what[4], st_a, sa, st_b, sb); }
    free(dsq);
  }
  printf("impl_b200_utest: profile %s M=%d, %d targets (%d rescaled Forward rows)\n", hmm->name ? hmm->name : "(sampled)", hmm->M, N + H, nresc);
  for (t = 0; t < 5; t++) printf("  %-18s %s (%d differing)\n", what[t], nfail[t] ? "DIFFERS" : "bit-identical", nfail[t]);

  esl_sq_Destroy(sq);
  p7_omx_Destroy(fa); p7_omx_Destroy(fb); p7_omx_Destroy(ba); p7_omx_Destroy(bb);
  p7_oprofile_Destroy(om); p7_profile_Destroy(gm); p7_bg_Destroy(bg); p7_hmm_Destroy(hmm);
  esl_alphabet_Destroy(abc); esl_randomness_Destroy(r); esl_getopts_Destroy(go);
  return (nfail[0] + nfail[1] + nfail[2] + nfail[3] + nfail[4]) ? 1 : 0;
}
