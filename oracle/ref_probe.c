/* oracle/ref_probe.c -- TEST INFRASTRUCTURE.  Golden-vector generator.
 *
 * Links against the UNMODIFIED reference objects (oracle/Makefile target `ref`) and drives the
 * reference's own functions in the order of its main() (src/RNAcode.c:115-221) and of
 * getExtremeValuePars() (src/score.c:1004-1048), dumping everything the scoring hot path consumes
 * and produces as one JSON document on stdout:
 *
 *   per block: rows (upper-cased as main() does), coordinates of row 0, PhyML tree + kappa,
 *   models[].scores / modelsRev[].scores / freqs / dist, the native HSS list returned by
 *   scoreAln(), and for each of the n null samples: its seed (oracle/ref_wrap.c), the simulated
 *   rows (first --dump-samples K samples only) and the best score the reference assigns to it;
 *   finally the Gumbel fit (mu, lambda) of EVDMaxLikelyFit and the resulting p-values.
 *
 * usage: ref_probe [-n samples] [--dump-samples K] [--max-blocks B] [--pars D,O,o,s] file
 * Floats are printed with %.9g (round-trips IEEE float32).
 */
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "RNAcode.h"
#include "extreme_fit.h"
#include "misc.h"
#include "rnaz_utils.h"
#include "score.h"
#include "treeML.h"
#include "treeSimulate.h"
#include "model.h"

extern parameters pars;
extern bgModel *models, *modelsRev;
extern float ****Sk, ****Sk_native, ****Sk_native_rev;
extern int BLOSUM62[24][24];
extern int BLOSUM90[24][24];
extern int transcode[4][4][4];

unsigned long rc_det_seed(unsigned long base, unsigned long block, unsigned long sample);
long rc_wrap_block(void);

static void json_str(const char *s) {
  putchar('"');
  for (; *s; s++) {
    if (*s == '"' || *s == '\\') putchar('\\');
    putchar(*s);
  }
  putchar('"');
}

/* pre-order walk in the order EvolveSequences visits the nodes (seqgen/evolve.c:400-433) */
static void dump_node(TTree *tree, TNode *node, int parent, int *counter, const struct aln *aln[], int N) {
  int me = (*counter)++, k, row = -1;
  double cum[16];
  if (node->tipNo != -1)
    for (k = 0; k < N; k++)
      if (strcmp(aln[k]->name, tree->names[node->tipNo]) == 0) row = k;
  SetMatrix(cum, node->length0 * 1.0); /* MutateSequence -> SetMatrix(matrix[0], len), NoRates (seqgen/evolve.c:291-292) */
  printf("%s{\"parent\":%d,\"row\":%d,\"len\":%.17g,\"cum\":[", me ? "," : "", parent, row, node->length0);
  for (k = 0; k < 16; k++) printf("%s%.17g", k ? "," : "", parent < 0 ? 0.0 : cum[k]);
  printf("]}");
  if (node->tipNo == -1) {
    dump_node(tree, node->branch1, me, counter, aln, N);
    dump_node(tree, node->branch2, me, counter, aln, N);
    if (parent < 0 && !tree->rooted) dump_node(tree, node->branch0, me, counter, aln, N);
  }
}

static void dump_hss(segmentStats *r) {
  int i;
  printf("[");
  for (i = 0; r[i].score > 0.0; i++) {
    if (i) printf(",");
    printf("{\"strand\":\"%c\",\"frame\":%d,\"startSite\":%d,\"endSite\":%d,\"start\":%d,\"end\":%d,"
           "\"startGenomic\":%d,\"endGenomic\":%d,\"score\":%.9g}",
           r[i].strand, r[i].frame, r[i].startSite, r[i].endSite, r[i].start, r[i].end, r[i].startGenomic,
           r[i].endGenomic, r[i].score);
  }
  printf("]");
}

int main(int argc, char *argv[]) {
  int sampleN = 100, dumpSamples = 0, maxBlocks = 1 << 30, skRows = 0;
  const char *file = NULL;
  int a, i, j, k, N, L, cols, blockIdx = 0, first = 1;
  struct aln *inputAln[MAX_NUM_NAMES];
  struct aln *inputAlnRev[MAX_NUM_NAMES];
  struct aln *sampledAln[MAX_NUM_NAMES];
  int (*readFunction)(FILE *, struct aln *[]) = NULL;
  unsigned long base = getenv("RNACODE_SEED") ? strtoul(getenv("RNACODE_SEED"), NULL, 10) : 1UL;

  /* defaults exactly as src/RNAcode.c:68-88 */
  pars.Delta = -10.0;
  pars.Omega = -4.0;
  pars.omega = -2.0;
  pars.stopPenalty_k = -8.0;
  pars.stopPenalty_0 = -9999.0;
  pars.outputFile = stdout;
  pars.debugFile = stdout;
  pars.bestOnly = 0;
  pars.bestRegion = 0;
  pars.stopEarly = 0;
  pars.postscript = 0;
  pars.postscript_cutoff = 0.05;
  pars.blosum = 62;
  strcpy(pars.limit, "");
  pars.cutoff = 1.0;
  pars.outputFormat = 0;

  for (a = 1; a < argc; a++) {
    if (!strcmp(argv[a], "-n") && a + 1 < argc) sampleN = atoi(argv[++a]);
    else if (!strcmp(argv[a], "--dump-samples") && a + 1 < argc) dumpSamples = atoi(argv[++a]);
    else if (!strcmp(argv[a], "--max-blocks") && a + 1 < argc) maxBlocks = atoi(argv[++a]);
    else if (!strcmp(argv[a], "--sk-rows") && a + 1 < argc) skRows = atoi(argv[++a]);
    else if (!strcmp(argv[a], "--blosum") && a + 1 < argc) pars.blosum = atoi(argv[++a]);
    else if (!strcmp(argv[a], "--pars") && a + 1 < argc) {
      /* same quirk as src/RNAcode.c:318: the 4th number lands in stopPenalty_0 */
      sscanf(argv[++a], "%f,%f,%f,%f", &pars.Delta, &pars.Omega, &pars.omega, &pars.stopPenalty_0);
    } else file = argv[a];
  }
  if (!file) {
    fprintf(stderr, "usage: ref_probe [-n N] [--dump-samples K] [--max-blocks B] [--pars D,O,o,s] file\n");
    return 2;
  }
  pars.sampleN = sampleN;
  pars.inputFile = fopen(file, "r");
  if (!pars.inputFile) {
    perror(file);
    return 2;
  }

  ntMap['A'] = ntMap['a'] = 0;
  ntMap['C'] = ntMap['c'] = 1;
  ntMap['G'] = ntMap['g'] = 2;
  ntMap['T'] = ntMap['t'] = 3;
  ntMap['U'] = ntMap['u'] = 3;

  switch (checkFormat(pars.inputFile)) {
    case CLUSTAL: readFunction = &read_clustal; break;
    case MAF: readFunction = &read_maf; break;
    default: fprintf(stderr, "unknown format\n"); return 2;
  }

  printf("{\"source\":");
  json_str(strrchr(file, '/') ? strrchr(file, '/') + 1 : file);
  printf(",\"sampleN\":%d,\"base_seed\":%lu,\"params\":{\"Delta\":%.9g,\"Omega\":%.9g,\"omega\":%.9g,"
         "\"stopPenalty_0\":%.9g,\"stopPenalty_k\":%.9g},\n",
         sampleN, base, pars.Delta, pars.Omega, pars.omega, pars.stopPenalty_0, pars.stopPenalty_k);
  printf("\"transcode\":[");
  for (i = 0; i < 64; i++) printf("%s%d", i ? "," : "", transcode[i / 16][(i / 4) % 4][i % 4]);
  printf("],\n\"blosum\":[");
  for (i = 0; i < 576; i++) printf("%s%d", i ? "," : "", pars.blosum == 62 ? BLOSUM62[i / 24][i % 24] : BLOSUM90[i / 24][i % 24]);
  printf("],\n\"blocks\":[\n");

  while (readFunction(pars.inputFile, inputAln) != 0) {
    char *treeString;
    float kappa, maxScore, parMu = 0, parLambda = 0;
    TTree *tree;
    segmentStats *results;
    int hssCount, fitStatus;
    double *maxScores, mu, lambda;

    if (blockIdx >= maxBlocks) break;
    for (i = 0; inputAln[i] != NULL; i++)
      for (j = 0; inputAln[i]->seq[j]; j++) inputAln[i]->seq[j] = toupper(inputAln[i]->seq[j]);
    L = getSeqLength(inputAln[0]->seq);
    cols = strlen(inputAln[0]->seq);
    for (N = 0; inputAln[N] != NULL; N++);

    if (!first) printf(",\n");
    first = 0;
    printf("{\"index\":%d,\"N\":%d,\"cols\":%d,\"L\":%d,\"start\":%d,\"length\":%d,\"strand\":\"%c\",\"names\":[", blockIdx,
           N, cols, L, inputAln[0]->start, inputAln[0]->length, inputAln[0]->strand ? inputAln[0]->strand : '?');
    for (i = 0; i < N; i++) {
      if (i) printf(",");
      json_str(inputAln[i]->name);
    }
    printf("],\"rows\":[");
    for (i = 0; i < N; i++) {
      if (i) printf(",");
      json_str(inputAln[i]->seq);
    }
    printf("]");
    blockIdx++;

    if (N <= 2 || L < 3) { /* src/RNAcode.c:142-150 */
      printf(",\"skipped\":true}");
      continue;
    }
    if (treeML((const struct aln **)inputAln, &treeString, &kappa) == 0) {
      printf(",\"skipped\":true,\"tree_failed\":true}");
      continue;
    }
    printf(",\"scored_index\":%ld,\"kappa\":%.9g,\"tree\":", rc_wrap_block(), kappa);
    json_str(treeString);
    tree = string2tree(treeString);
    free(treeString);
    copyAln((struct aln **)inputAln, (struct aln **)inputAlnRev);
    revAln((struct aln **)inputAlnRev);
    models = getModels(tree, (struct aln **)inputAln, kappa);
    modelsRev = getModels(tree, inputAlnRev, kappa);

    printf(",\"freqs_fwd\":[%.9g,%.9g,%.9g,%.9g],\"freqs_rev\":[%.9g,%.9g,%.9g,%.9g],\"dist\":[", models[0].freqs[0],
           models[0].freqs[1], models[0].freqs[2], models[0].freqs[3], modelsRev[0].freqs[0], modelsRev[0].freqs[1],
           modelsRev[0].freqs[2], modelsRev[0].freqs[3]);
    for (i = 0; i < N; i++) printf("%s%.9g", i ? "," : "", models[i].dist);
    printf("],\"scores_fwd\":[");
    for (i = 0; i < N; i++)
      printf("%s[%.9g,%.9g,%.9g,%.9g]", i ? "," : "", models[i].scores[0], models[i].scores[1], models[i].scores[2],
             models[i].scores[3]);
    printf("],\"scores_rev\":[");
    for (i = 0; i < N; i++)
      printf("%s[%.9g,%.9g,%.9g,%.9g]", i ? "," : "", modelsRev[i].scores[0], modelsRev[i].scores[1],
             modelsRev[i].scores[2], modelsRev[i].scores[3]);
    printf("]");

    Sk = NULL;
    Sk_native = NULL;
    Sk_native_rev = NULL;
    results = scoreAln((const struct aln **)inputAln, tree, kappa, skRows > 0 && L <= 700);
    printf(",\"native_hss\":");
    dump_hss(results);
    if (skRows > 0 && L <= 700) { /* rows of Sk_native / Sk_native_rev as backtrack() reads them (src/score.c:558-797) */
      int r, x, s, firstRow = 1;
      printf(",\"sk_rows\":[");
      for (s = 0; s < 2; s++) {
        float ****M = s ? Sk_native_rev : Sk_native;
        for (r = 0; r < skRows; r++) {
          int b = r < 3 ? r + 1 : 1 + (int)((long)(r - 2) * (L - 3) / (skRows - 2));
          if (b > L - 2) b = L - 2;
          printf("%s{\"strand\":%d,\"b\":%d,\"v\":[", firstRow ? "" : ",", s, b);
          firstRow = 0;
          for (k = 1; k < N; k++)
            for (x = 0; x < 3; x++) {
              printf("%s[", (k > 1 || x > 0) ? "," : "");
              for (i = b - 1; i <= L; i += 3) printf("%s%.9g", i > b - 1 ? "," : "", M[k][x][b][i]);
              printf("]");
            }
          printf("]}");
        }
      }
      printf("]");
      freeSk(Sk_native, (const struct aln **)inputAln);
      freeSk(Sk_native_rev, (const struct aln **)inputAln);
      Sk_native = Sk_native_rev = NULL;
    }

    hssCount = 0;
    while (results[hssCount++].score > 0.0);
    qsort(results, hssCount, sizeof(segmentStats), compareScores);
    maxScore = results[0].score;
    printf(",\"maxNativeScore\":%.9g", maxScore);

    /* sampling loop, src/score.c:1004-1048 (no stop-early here: all n maxima are recorded) */
    maxScores = (double *)malloc(sizeof(double) * sampleN);
    printf(",\"seeds\":[");
    for (i = 0; i < sampleN; i++)
      printf("%s%lu", i ? "," : "", rc_det_seed(base, (unsigned long)rc_wrap_block(), (unsigned long)i));
    printf("],\"samples\":[");
    for (i = 0; i < sampleN; i++) {
      segmentStats *r;
      int hc;
      simulateTree(tree, models[0].freqs, models[0].kappa, cols);
      tree2aln(tree, sampledAln);
      sortAln((const struct aln **)inputAln, sampledAln);
      if (i < dumpSamples) { /* rows BEFORE reintroduceGaps, in input order */
        printf("%s[", i ? "," : "");
        for (k = 0; k < N; k++) {
          if (k) printf(",");
          json_str(sampledAln[k]->seq);
        }
        printf("]");
      }
      reintroduceGaps((const struct aln **)inputAln, sampledAln);
      r = scoreAln((const struct aln **)sampledAln, tree, kappa, 0);
      hc = 0;
      while (r[hc].score >= 0) hc++;
      qsort(r, hc, sizeof(segmentStats), compareScores);
      maxScores[i] = r[0].score;
      freeAln((struct aln **)sampledAln);
      freeResults(r);
    }
    { /* everything kernel (d) needs to redraw these samples: flattened tree + cumulative matrices of seq-gen */
      int counter = 0;
      printf("],\"evolve\":{\"rooted\":%d,\"addFreq\":[%.17g,%.17g,%.17g,%.17g],\"nodes\":[", tree->rooted, addFreq[0],
             addFreq[1], addFreq[2], addFreq[3]);
      if (sampleN > 0) dump_node(tree, tree->root, -1, &counter, (const struct aln **)inputAln, N);
      printf("]}");
    }
    printf(",\"maxScores\":[");
    for (i = 0; i < sampleN; i++) printf("%s%.9g", i ? "," : "", maxScores[i]);
    printf("]");
    fitStatus = sampleN > 0 ? EVDMaxLikelyFit(maxScores, NULL, sampleN, &mu, &lambda) : 0;
    if (fitStatus == 1) {
      parMu = mu;
      parLambda = lambda;
    }
    printf(",\"fit\":%d,\"mu\":%.9g,\"lambda\":%.9g,\"pvalues\":[", fitStatus, parMu, parLambda);
    for (i = 0; results[i].score > 0.0; i++) {
      float pv = (fitStatus == 1) ? 1 - exp((-1) * exp((-1) * parLambda * (results[i].score - parMu))) : 99.0;
      printf("%s[%.9g,%.9g]", i ? "," : "", results[i].score, pv);
    }
    printf("]}");
    free(maxScores);
    freeSk(Sk, (const struct aln **)inputAln);
    Sk = NULL;
    freeResults(results);
    freeSeqgenTree(tree);
    freeAln((struct aln **)inputAln);
    freeAln((struct aln **)inputAlnRev);
    /* models / Sk intentionally leaked: short-lived process */
  }
  printf("\n]}\n");
  return 0;
}
