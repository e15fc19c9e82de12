/* integration/rnacode_cuda_host.h -- helpers shared by the two reference-side bindings of libRNAcode_cuda
 * (rnacode_cuda_shim.c: the reference's own main() with scoreAln / getExtremeValuePars replaced;
 *  rnacode_pipeline.c: the batched multi-process driver).  Everything here talks to the reference through its
 * own headers and to the GPU library through include/rnacode_cuda.h. */
#ifndef RNACODE_CUDA_HOST_H
#define RNACODE_CUDA_HOST_H

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

/* Row of the sorted null alignment each tree tip ends up in.  tree2aln (src/treeSimulate.c:254-283) lists the simulated rows in
 * tip order under the names of the tree, and sortAln (src/misc.c:150-171) brings them into input order with a double loop of
 * swaps that never stops at the first match -- for the usual one-to-one names that is the permutation "tip -> row of that name",
 * and for duplicate or unmatched names it still leaves SOME arrangement, which is what the reference then scores.  The same
 * swaps on (name, tip index) pairs reproduce that arrangement for every input.  row_of_tip has tree->numTips entries. */
static void sorted_rows_of_tips(TTree *tree, const struct aln *alignment[], int N, int *row_of_tip) {
  const int T = tree->numTips;
  const char **nm = (const char **)malloc(sizeof(char *) * (size_t)(T > 0 ? T : 1));
  int *src = (int *)malloc(sizeof(int) * (size_t)(T > 0 ? T : 1));
  int i, j;
  for (j = 0; j < T; j++) {
    nm[j] = tree->names[j];
    src[j] = j;
  }
  for (i = 0; i < N && i < T; i++)
    for (j = 0; j < T; j++)
      if (strcmp(alignment[i]->name, nm[j]) == 0) {
        const char *tn = nm[j];
        const int ts = src[j];
        nm[j] = nm[i];
        src[j] = src[i];
        nm[i] = tn;
        src[i] = ts;
      }
  for (i = 0; i < T; i++) row_of_tip[src[i]] = i < N ? i : -1;
  free(nm);
  free(src);
}

/* The tree flattened in the order EvolveSequences visits its nodes (seqgen/evolve.c:400-433): root, subtree of
 * branch1, of branch2 and, for the unrooted trees PhyML writes, of branch0.  cum = what MutateSequence would pass to
 * SetState for the branch above the node (SetMatrix(matrix[0], length0 * 1.0), NoRates, seqgen/evolve.c:291-292). */
static void flatten_rec(TTree *tree, TNode *node, int parent, const int *row_of_tip, int *n, int *par, int *row, double *cum) {
  int me = (*n)++, k;
  par[me] = parent;
  row[me] = node->tipNo != -1 ? row_of_tip[node->tipNo] : -1;
  if (parent < 0) {
    for (k = 0; k < 16; k++) cum[k] = 0.0;
    for (k = 0; k < 4; k++) cum[k] = addFreq[k]; /* RandomSequence draws from the cumulative frequencies */
  } else {
    SetMatrix(cum + (size_t)me * 16, node->length0 * 1.0);
  }
  if (node->tipNo == -1) {
    flatten_rec(tree, node->branch1, me, row_of_tip, n, par, row, cum);
    flatten_rec(tree, node->branch2, me, row_of_tip, n, par, row, cum);
    if (parent < 0 && !tree->rooted) flatten_rec(tree, node->branch0, me, row_of_tip, n, par, row, cum);
  }
}
static void flatten_node(TTree *tree, TNode *node, int parent, const struct aln *alignment[], int N, int *n, int *par,
                         int *row, double *cum) {
  int *row_of_tip = (int *)malloc(sizeof(int) * (size_t)(tree->numTips > 0 ? tree->numTips : 1));
  sorted_rows_of_tips(tree, alignment, N, row_of_tip);
  flatten_rec(tree, node, parent, row_of_tip, n, par, row, cum);
  free(row_of_tip);
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


/* segmentStats exactly as getHSS fills them (src/score.c:908-940), sentinel as src/score.c:1129-1134 */
static segmentStats *hss_to_segments(const struct aln *inputAln[], const rc_hss *h, int n) {
  int i;
  segmentStats *res = (segmentStats *)malloc(sizeof(segmentStats) * (n + 1));
  memset(res, 0, sizeof(segmentStats) * (n + 1));
  for (i = 0; i < n; i++) {
    segmentStats *r = &res[i];
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
  return res;
}

/* ---- --eps: backtrack() on rows computed by the GPU ------------------------------------------------------------
 * colorAln (src/postscript.c:303-305) calls backtrack(b, i, Sk_native or Sk_native_rev, currAln) for up to three
 * regions per plotted hit; backtrack (src/score.c:558-797) reads SSk[k][state][b][.] of that single row b only.  Linked
 * with -Wl,--wrap=backtrack: the row comes from rc_pair_rows(), is hung into an otherwise empty SSk skeleton and the
 * reference's own backtrack() walks it.  `alignment` is already in the strand's orientation (:255-259); which of the
 * two global matrices was passed tells the strand, i.e. whose background model scores apply.
 * The batched pipeline has no global models at print time and points these at the block being printed: */
static const float *g_bt_scores_fwd = NULL, *g_bt_scores_rev = NULL;
static const int *g_bt_blosum = NULL;

backtrackData *__real_backtrack(int opt_b, int opt_i, float ****SSk, const struct aln *alignment[]);

backtrackData *__wrap_backtrack(int opt_b, int opt_i, float ****SSk, const struct aln *alignment[]) {
  const int rev = (SSk == Sk_native_rev);
  rc_block_desc d;
  rc_params p = current_params();
  int N, k, x, i, j, cols, L, nrows, *blosum;
  float *sc, *row, ****tmp;
  char *rows;
  backtrackData *out;

  for (N = 0; alignment[N] != NULL; N++);
  cols = (int)strlen(alignment[0]->seq);
  L = getSeqLength(alignment[0]->seq);
  rows = (char *)malloc((size_t)N * cols);
  sc = (float *)malloc(sizeof(float) * 4 * N);
  blosum = (int *)malloc(sizeof(int) * 576);
  for (k = 0; k < N; k++) {
    memcpy(rows + (size_t)k * cols, alignment[k]->seq, cols);
    for (i = 0; i < 4; i++)
      sc[4 * k + i] = g_bt_scores_fwd ? (rev ? g_bt_scores_rev : g_bt_scores_fwd)[4 * k + i]
                                      : (rev ? modelsRev : models)[k].scores[i];
  }
  for (i = 0; i < 24; i++)
    for (j = 0; j < 24; j++) blosum[i * 24 + j] = g_bt_blosum ? g_bt_blosum[i * 24 + j] : models[0].matrix[i][j];
  d.N = N;
  d.cols = cols;
  d.rows = rows;
  d.scores_fwd = sc; /* the rows are scored as given (strand 0 of what colorAln passes) with that strand's models */
  d.scores_rev = sc;
  d.n_samples = 0;
  d.samples = NULL;

  row = (float *)calloc((size_t)N * 3 * (L + 1) + 8, sizeof(float));
  if (opt_b >= 1 && opt_b <= L && rc_pair_rows(ctx(), &d, &p, blosum, 0, 1, &opt_b, row) != RC_OK) die("rc_pair_rows");

  nrows = (opt_b > L ? opt_b : L) + 1;
  tmp = (float ****)malloc(sizeof(float ***) * (N + 1));
  for (k = 0; k < N; k++) {
    tmp[k] = (float ***)malloc(sizeof(float **) * 3);
    for (x = 0; x < 3; x++) {
      tmp[k][x] = (float **)calloc(nrows, sizeof(float *));
      if (opt_b >= 0) tmp[k][x][opt_b] = row + ((size_t)k * 3 + x) * (L + 1);
    }
  }
  out = __real_backtrack(opt_b, opt_i, tmp, alignment);
  for (k = 0; k < N; k++) {
    for (x = 0; x < 3; x++) free(tmp[k][x]);
    free(tmp[k]);
  }
  free(tmp);
  free(row);
  free(rows);
  free(sc);
  free(blosum);
  return out;
}

#endif
