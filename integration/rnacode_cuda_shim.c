/* integration/rnacode_cuda_shim.c -- the reference-side binding of libRNAcode_cuda.
 *
 * This is the (only) code a maintainer of the reference adds to make RNAcode a GPU program: the two
 * entry points of the scoring core that main() calls,
 *
 *     scoreAln()              src/score.h:116, called at src/RNAcode.c:171
 *     getExtremeValuePars()   src/score.h:103, called at src/RNAcode.c:180
 *
 * are re-implemented on top of the C ABI in include/rnacode_cuda.h.  Nothing else of the reference
 * changes: parsing, PhyML (treeML), getModels, seq-gen (simulateTree/tree2aln/sortAln), the Gumbel fit
 * (EVDMaxLikelyFit), the p-value formula and printResults are the reference's own objects.
 *
 * It is linked with GNU ld's --wrap so that the UNMODIFIED reference objects can be reused as they are
 * (oracle/Makefile, target `cli`):   -Wl,--wrap=scoreAln -Wl,--wrap=getExtremeValuePars
 * In a source tree one would instead delete the two functions from src/score.c and compile this file.
 *
 * Exact mode: the null alignments come from the reference's own seq-gen (same RNG stream), are scored on the
 * GPU, and the maxima feed the reference's own EVD fit, so HSS, scores and p-values are identical.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "RNAcode.h"
#include "extreme_fit.h"
#include "misc.h"
#include "rnaz_utils.h"
#include "score.h"
#include "treeSimulate.h"

#include "rnacode_cuda.h"

extern parameters pars;
extern bgModel *models, *modelsRev;
extern float ****Sk, ****Sk_native, ****Sk_native_rev;

static rc_ctx *g_ctx = NULL;

static rc_ctx *ctx(void) {
  if (!g_ctx) {
    const char *dev = getenv("RNACODE_CUDA_DEVICE");
    if (rc_create(&g_ctx, dev ? atoi(dev) : 0) != RC_OK) {
      fprintf(stderr, "RNAcode: no usable CUDA device (libRNAcode_cuda has no CPU fallback)\n");
      exit(EXIT_FAILURE);
    }
  }
  return g_ctx;
}

static void die(const char *what) {
  fprintf(stderr, "RNAcode: %s: %s\n", what, rc_last_error(g_ctx));
  exit(EXIT_FAILURE);
}

/* main() frees Sk / Sk_native / Sk_native_rev row by row after every block (src/RNAcode.c:192-209); give it
 * something of the right shape to free.  The GPU path never materialises these matrices. */
static float ****tiny_sk(int N, int L) {
  int k, x, i;
  float ****S = (float ****)malloc(sizeof(float ***) * (N + 1));
  for (k = 0; k < N; k++) {
    S[k] = (float ***)malloc(sizeof(float **) * 3);
    for (x = 0; x < 3; x++) {
      S[k][x] = (float **)malloc(sizeof(float *) * (L + 1));
      for (i = 0; i < L + 1; i++) S[k][x][i] = NULL; /* free(NULL) is a no-op */
    }
  }
  return S;
}

static void fill_desc(const struct aln *alignment[], rc_block_desc *d, char **rows, float **sf, float **sr, int **blosum) {
  int N, k, i, j, cols;
  for (N = 0; alignment[N] != NULL; N++);
  cols = (int)strlen(alignment[0]->seq);
  *rows = (char *)malloc((size_t)N * cols);
  *sf = (float *)malloc(sizeof(float) * 4 * N);
  *sr = (float *)malloc(sizeof(float) * 4 * N);
  *blosum = (int *)malloc(sizeof(int) * 576);
  for (k = 0; k < N; k++) {
    memcpy(*rows + (size_t)k * cols, alignment[k]->seq, cols);
    for (i = 0; i < 4; i++) {
      (*sf)[4 * k + i] = models[k].scores[i];
      (*sr)[4 * k + i] = modelsRev[k].scores[i];
    }
  }
  for (i = 0; i < 24; i++)
    for (j = 0; j < 24; j++) (*blosum)[i * 24 + j] = models[0].matrix[i][j];
  d->N = N;
  d->cols = cols;
  d->rows = *rows;
  d->scores_fwd = *sf;
  d->scores_rev = *sr;
  d->n_samples = 0;
  d->samples = NULL;
}

static rc_params current_params(void) {
  rc_params p;
  p.Delta = pars.Delta;
  p.Omega = pars.Omega;
  p.omega = pars.omega;
  p.stopPenalty_0 = pars.stopPenalty_0;
  p.stopPenalty_k = pars.stopPenalty_k;
  return p;
}

segmentStats *__wrap_scoreAln(const struct aln *inputAln[], TTree *tree, float kappa, int backtrack) {
  rc_block_desc d;
  char *rows;
  float *sf, *sr;
  int *blosum, n = 0, cap = 256, i, rc, L;
  rc_hss *h;
  rc_params p = current_params();
  segmentStats *res;
  (void)tree;
  (void)kappa;

  fill_desc(inputAln, &d, &rows, &sf, &sr, &blosum);
  h = (rc_hss *)malloc(sizeof(rc_hss) * cap);
  rc = rc_score_aln(ctx(), &d, &p, blosum, h, cap, &n);
  if (rc == RC_ERR_CAPACITY) {
    cap = n;
    h = (rc_hss *)realloc(h, sizeof(rc_hss) * cap);
    rc = rc_score_aln(ctx(), &d, &p, blosum, h, cap, &n);
  }
  if (rc != RC_OK) die("rc_score_aln");

  /* segmentStats exactly as getHSS fills them (src/score.c:908-940), sentinel as src/score.c:1129-1134 */
  res = (segmentStats *)malloc(sizeof(segmentStats) * (n + 1));
  for (i = 0; i < n; i++) {
    segmentStats *r = &res[i];
    memset(r, 0, sizeof(*r));
    r->name = strdup(inputAln[0]->name);
    r->strand = h[i].strand;
    r->frame = h[i].frame;
    r->startSite = h[i].startSite;
    r->endSite = h[i].endSite;
    r->score = h[i].score;
    r->start = h[i].startSite * 3 + h[i].frame + 1;
    r->end = h[i].endSite * 3 + h[i].frame + 3;
    if ((inputAln[0]->start == 0) && (inputAln[0]->length == 0)) {
      r->startGenomic = r->start;
      r->endGenomic = r->end;
    } else if (h[i].strand == '+') {
      r->startGenomic = inputAln[0]->start + h[i].startSite * 3 + h[i].frame;
      r->endGenomic = inputAln[0]->start + h[i].endSite * 3 + h[i].frame + 2;
    } else {
      r->endGenomic = (inputAln[0]->start + inputAln[0]->length - 1) - h[i].startSite * 3 - h[i].frame;
      r->startGenomic = (inputAln[0]->start + inputAln[0]->length - 1) - h[i].endSite * 3 - h[i].frame - 2;
    }
  }
  if (n == 0) res[0].pvalue = 1.0;
  res[n].score = -1.0;

  if (backtrack) { /* main() expects to free these (see tiny_sk) */
    L = getSeqLength(inputAln[0]->seq);
    if (Sk == NULL) Sk = tiny_sk(d.N, L);
    if (Sk_native == NULL) {
      Sk_native = tiny_sk(d.N, L);
      Sk_native_rev = tiny_sk(d.N, L);
    }
  }
  free(h);
  free(rows);
  free(sf);
  free(sr);
  free(blosum);
  return res;
}

int __wrap_getExtremeValuePars(TTree *tree, const struct aln *alignment[], int sampleN, float maxNativeScore, float *parMu,
                               float *parLambda) {
  rc_block_desc d;
  char *rows, *samples;
  float *sf, *sr;
  int *blosum, i, k, done, betterThanNative = 0, stopCutoff, status = 1;
  double *maxScores, mu, lambda;
  rc_params p = current_params();
  struct aln *sampledAln[MAX_NUM_NAMES];
  /* under --stop-early work in small batches so that little is simulated past the stopping point;
   * seeds depend on (block, sample) only, so batching never changes the drawn alignments */
  const int batch = pars.stopEarly ? 32 : sampleN;
  size_t per;

  stopCutoff = (int)(pars.cutoff * pars.sampleN); /* src/score.c:992 */
  fill_desc(alignment, &d, &rows, &sf, &sr, &blosum);
  per = (size_t)d.N * d.cols;
  samples = (char *)malloc(per * (size_t)(batch > 0 ? batch : 1));
  maxScores = (double *)malloc(sizeof(double) * (sampleN > 0 ? sampleN : 1));

  for (done = 0; done < sampleN && status == 1;) {
    int nb = sampleN - done < batch ? sampleN - done : batch;
    for (i = 0; i < nb; i++) { /* src/score.c:1006-1010; the library re-imposes the native gaps itself */
      simulateTree(tree, models[0].freqs, models[0].kappa, d.cols);
      tree2aln(tree, sampledAln);
      sortAln(alignment, sampledAln);
      for (k = 0; k < d.N; k++) memcpy(samples + per * i + (size_t)k * d.cols, sampledAln[k]->seq, d.cols);
      freeAln((struct aln **)sampledAln);
    }
    d.n_samples = nb;
    d.samples = samples;
    if (rc_score_samples(ctx(), &d, &p, blosum, maxScores + done) != RC_OK) die("rc_score_samples");
    for (i = 0; i < nb; i++) { /* src/score.c:1036-1042, in sample order */
      if ((float)maxScores[done + i] > maxNativeScore) betterThanNative++;
      if (pars.stopEarly && betterThanNative > stopCutoff) {
        status = -1;
        break;
      }
    }
    done += nb;
  }
  if (status == 1) {
    if (EVDMaxLikelyFit(maxScores, NULL, sampleN, &mu, &lambda) == 1) { /* src/score.c:1050-1062 */
      *parMu = mu;
      *parLambda = lambda;
    } else {
      status = -1;
    }
  }
  free(maxScores);
  free(samples);
  free(rows);
  free(sf);
  free(sr);
  free(blosum);
  return status;
}
