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
 * Exact mode: the null alignments are the ones the reference's own seq-gen would draw -- the same MT19937 stream
 * (one CreateSeed() per sample, as src/treeSimulate.c:84), consumed in seq-gen's order, with the transition
 * matrices seq-gen's own SetMatrix() computes -- and the maxima feed the reference's own EVD fit, so HSS, scores
 * and p-values are identical.  By default the simulation itself runs on the GPU (kernel d, rc_score_samples_evolve);
 * RNACODE_CUDA_EVOLVE=host keeps simulateTree/tree2aln/sortAln on the host, RNACODE_CUDA_EVOLVE=philox switches to
 * the counter-based GPU generator (same distribution, different stream: GPU-RNG mode).
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

/* seq-gen's model state, for the on-GPU simulation of the null alignments */
#include "model.h"
#include "nucmodels.h"
#include "twister.h"

#include "rnacode_cuda.h"

extern parameters pars;
extern int numSites, equalTstv, numTaxa;
extern double tstv;
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

/* The tree flattened in the order EvolveSequences visits its nodes (seqgen/evolve.c:400-433): root, subtree of
 * branch1, of branch2 and, for the unrooted trees PhyML writes, of branch0.  cum = what MutateSequence would pass to
 * SetState for the branch above the node (SetMatrix(matrix[0], length0 * 1.0), NoRates, seqgen/evolve.c:291-292). */
static void flatten_node(TTree *tree, TNode *node, int parent, const struct aln *alignment[], int N, int *n, int *par,
                         int *row, double *cum) {
  int me = (*n)++, k;
  par[me] = parent;
  row[me] = -1;
  if (node->tipNo != -1)
    for (k = 0; k < N; k++)
      if (strcmp(alignment[k]->name, tree->names[node->tipNo]) == 0) row[me] = k; /* sortAln, src/misc.c:150-171 */
  if (parent < 0) {
    for (k = 0; k < 16; k++) cum[k] = 0.0;
    for (k = 0; k < 4; k++) cum[k] = addFreq[k]; /* RandomSequence draws from the cumulative frequencies */
  } else {
    SetMatrix(cum + (size_t)me * 16, node->length0 * 1.0);
  }
  if (node->tipNo == -1) {
    flatten_node(tree, node->branch1, me, alignment, N, n, par, row, cum);
    flatten_node(tree, node->branch2, me, alignment, N, n, par, row, cum);
    if (parent < 0 && !tree->rooted) flatten_node(tree, node->branch0, me, alignment, N, n, par, row, cum);
  }
}

/* seq-gen's model set-up as simulateTree does it before evolving (src/treeSimulate.c:59-93), without evolving */
static void setup_seqgen_model(TTree *tree, const float freqs[], float kap, int L) {
  int i;
  double fR, fY;
  isNucModel = 1;
  numStates = 4;
  model = 0; /* HKY */
  equalFreqs = 0;
  equalTstv = 0;
  for (i = 0; i < 4; i++) nucFreq[i] = (double)freqs[i];
  fR = nucFreq[0] + nucFreq[2];
  fY = nucFreq[1] + nucFreq[3];
  tstv = (double)kap * (nucFreq[0] * nucFreq[2] + nucFreq[1] * nucFreq[3]) / (fR * fY);
  numSites = L;
  numTaxa = tree->numTips;
  SetModel(model);
}

/* 0: host seq-gen, 1: GPU MT19937 (exact), 2: GPU Philox */
static int evolve_mode(void) {
  const char *e = getenv("RNACODE_CUDA_EVOLVE");
  if (e && strcmp(e, "host") == 0) return 0;
  if (e && strcmp(e, "philox") == 0) return 2;
  return 1;
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
  const int mode = evolve_mode();
  size_t per;
  rc_tree_desc td;
  int *tpar = NULL, *trow = NULL, nn = 0;
  double *tcum = NULL;
  unsigned int *seeds = NULL;

  stopCutoff = (int)(pars.cutoff * pars.sampleN); /* src/score.c:992 */
  fill_desc(alignment, &d, &rows, &sf, &sr, &blosum);
  per = (size_t)d.N * d.cols;
  samples = (char *)malloc(mode == 0 ? per * (size_t)(batch > 0 ? batch : 1) : 16);
  maxScores = (double *)malloc(sizeof(double) * (sampleN > 0 ? sampleN : 1));
  if (mode != 0 && sampleN > 0) {
    const int max_nodes = 2 * tree->numTips + 2;
    tpar = (int *)malloc(sizeof(int) * max_nodes);
    trow = (int *)malloc(sizeof(int) * max_nodes);
    tcum = (double *)malloc(sizeof(double) * 16 * max_nodes);
    seeds = (unsigned int *)malloc(sizeof(unsigned int) * (batch > 0 ? batch : 1));
    setup_seqgen_model(tree, models[0].freqs, models[0].kappa, d.cols);
    flatten_node(tree, tree->root, -1, alignment, d.N, &nn, tpar, trow, tcum);
    td.n_nodes = nn;
    td.parent = tpar;
    td.row = trow;
    td.cum = tcum;
  }

  for (done = 0; done < sampleN && status == 1;) {
    int nb = sampleN - done < batch ? sampleN - done : batch;
    d.n_samples = nb;
    if (mode == 0) {
      for (i = 0; i < nb; i++) { /* src/score.c:1006-1010; the library re-imposes the native gaps itself */
        simulateTree(tree, models[0].freqs, models[0].kappa, d.cols);
        tree2aln(tree, sampledAln);
        sortAln(alignment, sampledAln);
        for (k = 0; k < d.N; k++) memcpy(samples + per * i + (size_t)k * d.cols, sampledAln[k]->seq, d.cols);
        freeAln((struct aln **)sampledAln);
      }
      d.samples = samples;
      if (rc_score_samples(ctx(), &d, &p, blosum, maxScores + done) != RC_OK) die("rc_score_samples");
    } else {
      /* one CreateSeed() per null alignment, in sample order, exactly where simulateTree would call it */
      for (i = 0; i < nb; i++) seeds[i] = (unsigned int)(CreateSeed() & 0xffffffffUL);
      d.samples = NULL;
      if (rc_score_samples_evolve(ctx(), &d, &td, seeds, mode == 2 ? RC_RNG_PHILOX : RC_RNG_MT19937, &p, blosum,
                                  maxScores + done) != RC_OK)
        die("rc_score_samples_evolve");
    }
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
  free(tpar);
  free(trow);
  free(tcum);
  free(seeds);
  free(rows);
  free(sf);
  free(sr);
  free(blosum);
  return status;
}
