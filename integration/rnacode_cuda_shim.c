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
 * (oracle/Makefile, target `cli`):   -Wl,--wrap=scoreAln -Wl,--wrap=getExtremeValuePars -Wl,--wrap=backtrack
 * (backtrack: --eps only; the Sk_native row it walks is computed on demand, see rnacode_cuda_host.h)
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

#include "rnacode_cuda_host.h"

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

  res = hss_to_segments(inputAln, h, n);

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
