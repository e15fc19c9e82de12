/* integration/rnacode_pipeline.c -- RNAcode_b200: the batched host pipeline around libRNAcode_cuda.
 *
 * Same command line, same input formats and same output as the reference's RNAcode (src/RNAcode.c:52-233); what
 * changes is the order of work.  The reference handles one alignment block per iteration of main()'s loop:
 * parse, PhyML (treeML), models, score, n x (simulate, score), fit, print.  Here blocks are taken in windows:
 *
 *   1. parse a window of blocks with the reference's readers (read_maf / read_clustal), same filters and messages;
 *   2. the per-block host stage for all of them in parallel: treeML, string2tree, getModels (both strands) and
 *      seq-gen's transition matrices for every branch.  PhyML and seq-gen keep global state and are not re-entrant,
 *      so the window is split over forked worker processes (one per host core) that send back plain tables;
 *   3. in the parent, per block: one CreateSeed() per null alignment, in the reference's order;
 *   4. ONE library batch for the whole window: native alignments and all null alignments of all blocks are
 *      simulated (kernel d), packed, scored and reduced on the GPU (rc_batch_*);
 *   5. in input order: sort, EVDMaxLikelyFit, p-values, printResults -- all the reference's own code.
 *
 * Windows overlap: the workers of window w+1 are forked before the parent waits for those of window w, so while the parent
 * collects, scores (GPU), fits and prints window w -- and parses window w+2 -- the host cores are already busy with the
 * trees of window w+1.  A worker that ends abnormally (PhyML's Warn_And_Exit, a signal) is fatal for the run, as it is for
 * the reference; its records are flushed block by block, so nothing a live worker has finished is lost before that.
 *
 * --stop-early (p = 99.0 as soon as more than cutoff*n null alignments beat the native score,
 * src/score.c:1036-1042): the count is monotone in the sample index, so it is evaluated in two rounds -- 32 null
 * alignments for every block, the remaining n-32 only for the blocks still undecided -- with the same outcome.
 *
 * --eps: colorAln / backtrack (src/postscript.c, src/score.c:558-797) are the reference's own; the rows of Sk_native they
 * walk are computed on the GPU when asked for (rc_pair_rows, see __wrap_backtrack in rnacode_cuda_host.h).
 *
 * Link: the reference's objects (RNAcode.c compiled with -Dmain=rnacode_reference_main provides the globals),
 * libRNAcode_cuda; see oracle/Makefile target `pipeline`.
 */
#include <ctype.h>
#include <math.h>
#include <pthread.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include "RNAcode.h"
#include "cmdline.h"
#include "code.h"
#include "extreme_fit.h"
#include "misc.h"
#include "rnaz_utils.h"
#include "score.h"
#include "treeML.h"
#include "treeSimulate.h"
#include "utils.h"

#include "model.h"
#include "nucmodels.h"
#include "twister.h"

#include "rnacode_cuda.h"

#include "rnacode_cuda_host.h"
#include "rnacode_maf_mmap.h"

extern long int hitCounter;
void freeModels(bgModel *models, int N);

/* deterministic test builds (oracle/ref_wrap.c) derive seeds from (scored block index, sample index) */
void rc_wrap_set_block(long b) __attribute__((weak));
void rc_wrap_set_sample(long s) __attribute__((weak));

typedef struct {
  struct aln **aln; /* NULL-terminated, owned */
  int N, L, cols;
  long scored_idx; /* index among the blocks that reach treeML */
  int ok;
  /* what the tree worker sends back (host_prepare) */
  float *sf, *sr;     /* models[k].scores[0..3] / modelsRev[k].scores[0..3], N*4 each */
  int n_nodes;        /* flattened tree, see rc_tree_desc */
  int *tpar, *trow;
  double *tcum;
  char *rows;
  /* scoring results */
  rc_hss *hss;
  int n_hss, hssCount, status, better;
  float maxScore;
  double *maxScores;
  segmentStats *results;
  unsigned int *seeds;
} blk_t;

static void init_gpus(void);

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static int n_workers(int nblocks) {
  const char *e = getenv("RNACODE_CUDA_WORKERS");
  long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
  if (n < 1) n = 1;
  if (n > 256) n = 256;
  if (n > nblocks) n = nblocks;
  return (int)n;
}

/* Everything the reference does on the host between parsing and scoring, for one block: PhyML tree and kappa
 * (treeML, src/RNAcode.c:153), string2tree, the background models of both strands (getModels, :161-165), and seq-gen's
 * model set-up with one cumulative transition matrix per branch (what simulateTree would hand to MutateSequence). */
static int host_prepare(blk_t *b) {
  char *ts = NULL;
  float kappa = 0.0f;
  struct aln *rev[MAX_NUM_NAMES];
  bgModel *mod, *modRev;
  TTree *tree;
  int k, j, max_nodes;
  if (treeML((const struct aln **)b->aln, &ts, &kappa) == 0) return 0;
  tree = string2tree(ts);
  free(ts);
  copyAln(b->aln, rev);
  revAln(rev);
  models = mod = getModels(tree, b->aln, kappa);
  modelsRev = modRev = getModels(tree, rev, kappa);
  freeAln(rev);
  b->sf = (float *)malloc(sizeof(float) * 4 * b->N);
  b->sr = (float *)malloc(sizeof(float) * 4 * b->N);
  for (k = 0; k < b->N; k++)
    for (j = 0; j < 4; j++) {
      b->sf[4 * k + j] = mod[k].scores[j];
      b->sr[4 * k + j] = modRev[k].scores[j];
    }
  max_nodes = 2 * tree->numTips + 2;
  b->tpar = (int *)malloc(sizeof(int) * max_nodes);
  b->trow = (int *)malloc(sizeof(int) * max_nodes);
  b->tcum = (double *)malloc(sizeof(double) * 16 * max_nodes);
  b->n_nodes = 0;
  setup_seqgen_model(tree, mod[0].freqs, mod[0].kappa, b->cols);
  flatten_node(tree, tree->root, -1, (const struct aln **)b->aln, b->N, &b->n_nodes, b->tpar, b->trow, b->tcum);
  freeSeqgenTree(tree);
  freeModels(mod, b->N);
  freeModels(modRev, b->N);
  return 1;
}

/* phase 2: host_prepare in forked workers (PhyML and seq-gen keep global state: processes, not threads); the workers pull
 * the next block from a shared counter and send their results back through one temporary file each.  host_stage_start()
 * forks and returns; host_stage_finish() waits for the workers and reads the results. */
typedef struct {
  blk_t *blk;
  int nb, P;
  FILE **chan;
  pid_t *pid;
  int *next;
  double t_start, t_done;
} host_job;

static int g_in_worker = 0;
/* PhyML leaves through exit() (Warn_And_Exit); in a forked worker that must not run the parent's atexit handlers (the CUDA
 * runtime's teardown among them).  Linked with -Wl,--wrap=exit. */
void __real_exit(int status);
void __wrap_exit(int status) {
  if (g_in_worker) {
    fflush(NULL);
    _exit(status ? status : 3);
  }
  __real_exit(status);
}

static void host_stage_start(host_job *j, blk_t *blk, int nb) {
  int w, i;
  const char *ke = getenv("RNACODE_CUDA_TEST_KILL_WORKER_AT");
  const long kill_at = ke ? atol(ke) : -1;
  memset(j, 0, sizeof(*j));
  j->blk = blk;
  j->nb = nb;
  j->P = nb > 0 ? n_workers(nb) : 0;
  j->t_start = now_s();
  if (j->P <= 1) return; /* no fork: host_stage_finish() prepares the blocks itself */
  j->chan = (FILE **)malloc(sizeof(FILE *) * j->P);
  j->pid = (pid_t *)malloc(sizeof(pid_t) * j->P);
  /* blocks differ a lot in cost (PhyML is about N^2 * cols): the workers pull the next block from a shared counter */
  j->next = (int *)mmap(NULL, 4096, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (j->next == MAP_FAILED) nrerror("ERROR: mmap failed.\n");
  *j->next = 0;
  fflush(NULL);
  for (w = 0; w < j->P; w++) {
    j->chan[w] = tmpfile();
    if (!j->chan[w]) nrerror("ERROR: could not create a temporary file for the tree workers.\n");
    j->pid[w] = fork();
    if (j->pid[w] < 0) nrerror("ERROR: fork failed.\n");
    if (j->pid[w] == 0) {
      g_in_worker = 1;
      while ((i = __sync_fetch_and_add(j->next, 1)) < nb) {
        blk_t *b = &blk[i];
        int ok;
        if (kill_at >= 0 && b->scored_idx == kill_at) raise(SIGKILL); /* test hook: a worker dies mid-window */
        ok = host_prepare(b);
        fwrite(&i, sizeof(int), 1, j->chan[w]);
        fwrite(&ok, sizeof(int), 1, j->chan[w]);
        if (ok) {
          fwrite(&b->n_nodes, sizeof(int), 1, j->chan[w]);
          fwrite(b->sf, sizeof(float), 4 * b->N, j->chan[w]);
          fwrite(b->sr, sizeof(float), 4 * b->N, j->chan[w]);
          fwrite(b->tpar, sizeof(int), b->n_nodes, j->chan[w]);
          fwrite(b->trow, sizeof(int), b->n_nodes, j->chan[w]);
          fwrite(b->tcum, sizeof(double), 16 * (size_t)b->n_nodes, j->chan[w]);
        }
        fflush(j->chan[w]); /* a record per block: what this worker has finished survives its death */
      }
      _exit(0);
    }
  }
}

static void host_stage_finish(host_job *j) {
  blk_t *blk = j->blk;
  const int nb = j->nb;
  int w, i, *seen;
  if (nb == 0) return;
  if (j->P <= 1) {
    for (i = 0; i < nb; i++) blk[i].ok = host_prepare(&blk[i]);
    init_gpus();
    j->t_done = now_s();
    return;
  }
  init_gpus(); /* bring the CUDA contexts up while the workers run PhyML (the children never touch CUDA) */
  seen = (int *)calloc(nb, sizeof(int));
  for (w = 0; w < j->P; w++) {
    int status = 0, idx, ok;
    if (waitpid(j->pid[w], &status, 0) < 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0) {
      /* the reference ends with an error when PhyML gives up on an alignment (Warn_And_Exit) or crashes: so do we,
       * instead of reporting the worker's remaining blocks as failed trees */
      if (WIFSIGNALED(status))
        fprintf(stderr, "ERROR: a tree worker was killed by signal %d.\n", WTERMSIG(status));
      else
        fprintf(stderr, "ERROR: a tree worker ended abnormally (exit status %d).\n", WIFEXITED(status) ? WEXITSTATUS(status) : -1);
      for (i = 0; i < j->P; i++)
        if (i != w) kill(j->pid[i], SIGTERM);
      exit(EXIT_FAILURE);
    }
    rewind(j->chan[w]);
    while (fread(&idx, sizeof(int), 1, j->chan[w]) == 1) {
      blk_t *b;
      int good;
      if (fread(&ok, sizeof(int), 1, j->chan[w]) != 1 || idx < 0 || idx >= nb) nrerror("ERROR: corrupt record from a tree worker.\n");
      seen[idx] = 1;
      if (!ok) continue;
      b = &blk[idx];
      if (fread(&b->n_nodes, sizeof(int), 1, j->chan[w]) != 1 || b->n_nodes < 2 || b->n_nodes > 2 * MAX_NUM_NAMES + 2)
        nrerror("ERROR: corrupt record from a tree worker.\n");
      b->sf = (float *)malloc(sizeof(float) * 4 * b->N);
      b->sr = (float *)malloc(sizeof(float) * 4 * b->N);
      b->tpar = (int *)malloc(sizeof(int) * b->n_nodes);
      b->trow = (int *)malloc(sizeof(int) * b->n_nodes);
      b->tcum = (double *)malloc(sizeof(double) * 16 * b->n_nodes);
      good = fread(b->sf, sizeof(float), 4 * b->N, j->chan[w]) == (size_t)(4 * b->N) &&
             fread(b->sr, sizeof(float), 4 * b->N, j->chan[w]) == (size_t)(4 * b->N) &&
             fread(b->tpar, sizeof(int), b->n_nodes, j->chan[w]) == (size_t)b->n_nodes &&
             fread(b->trow, sizeof(int), b->n_nodes, j->chan[w]) == (size_t)b->n_nodes &&
             fread(b->tcum, sizeof(double), 16 * (size_t)b->n_nodes, j->chan[w]) == 16 * (size_t)b->n_nodes;
      if (!good) nrerror("ERROR: truncated record from a tree worker.\n");
      b->ok = 1;
    }
    fclose(j->chan[w]);
  }
  for (i = 0; i < nb; i++)
    if (!seen[i]) nrerror("ERROR: a block was not processed by any tree worker.\n");
  free(seen);
  munmap(j->next, 4096);
  free(j->chan);
  free(j->pid);
  j->t_done = now_s();
}

/* ---- GPUs: blocks are independent, so a window is dealt out over the devices by cost; no exchange between them ---- */
#define MAX_GPUS 16
static rc_ctx *g_ctxs[MAX_GPUS];
static int g_ngpus = 0;

static void init_gpus(void) {
  const char *e = getenv("RNACODE_CUDA_GPUS"), *d = getenv("RNACODE_CUDA_DEVICE");
  int want = e ? (strcmp(e, "all") == 0 ? -1 : atoi(e)) : 1, have = 0, base = d ? atoi(d) : 0, i;
  if (g_ngpus) return;
  if (rc_device_count(&have) != RC_OK || have < 1) {
    fprintf(stderr, "RNAcode: no usable CUDA device (libRNAcode_cuda has no CPU fallback)\n");
    exit(EXIT_FAILURE);
  }
  if (want < 0 || want > have - base) want = have - base;
  if (want < 1) want = 1;
  if (want > MAX_GPUS) want = MAX_GPUS;
  for (i = 0; i < want; i++)
    if (rc_create(&g_ctxs[i], base + i) != RC_OK) {
      fprintf(stderr, "RNAcode: cannot use CUDA device %d\n", base + i);
      exit(EXIT_FAILURE);
    }
  g_ctx = g_ctxs[0];
  g_ngpus = want;
}

/* unit of GPU work: the native alignment of a block plus its null alignments [s0, s0 + ns) */
typedef struct {
  int blk, s0, ns, want_native;
  double cost;
} unit_t;

typedef struct {
  rc_ctx *c;
  blk_t *blk;
  unit_t *units;
  int nunits, seed0, mode;
  const int *blosum;
} shard_t;

static void shard_die(rc_ctx *c, const char *what) {
  fprintf(stderr, "RNAcode: %s: %s\n", what, rc_last_error(c));
  exit(EXIT_FAILURE);
}

/* one library batch on one device */
static void *run_shard(void *arg) {
  shard_t *sh = (shard_t *)arg;
  rc_params p = current_params();
  rc_block_desc *descs;
  rc_batch *batch = NULL;
  int k;
  if (sh->nunits == 0) return NULL;
  descs = (rc_block_desc *)malloc(sizeof(rc_block_desc) * sh->nunits);
  for (k = 0; k < sh->nunits; k++) {
    blk_t *b = &sh->blk[sh->units[k].blk];
    descs[k].N = b->N;
    descs[k].cols = b->cols;
    descs[k].rows = b->rows;
    descs[k].scores_fwd = b->sf;
    descs[k].scores_rev = b->sr;
    descs[k].n_samples = sh->units[k].ns;
    descs[k].samples = NULL;
  }
  double tt[6];
  tt[0] = now_s();
  if (rc_batch_create(sh->c, descs, sh->nunits, &p, sh->blosum, &batch) != RC_OK) shard_die(sh->c, "rc_batch_create");
  tt[1] = now_s();
  for (k = 0; k < sh->nunits; k++) {
    const unit_t *u = &sh->units[k];
    blk_t *b = &sh->blk[u->blk];
    rc_tree_desc td;
    if (u->ns == 0) continue;
    td.n_nodes = b->n_nodes;
    td.parent = b->tpar;
    td.row = b->trow;
    td.cum = b->tcum;
    if (rc_batch_set_evolve(batch, k, &td, b->seeds + (u->s0 - sh->seed0), sh->mode == 2 ? RC_RNG_PHILOX : RC_RNG_MT19937) != RC_OK)
      shard_die(sh->c, "rc_batch_set_evolve");
  }
  tt[2] = now_s();
  if (rc_batch_upload(batch) != RC_OK) shard_die(sh->c, "rc_batch_upload");
  tt[3] = now_s();
  if (rc_batch_run(batch) != RC_OK) shard_die(sh->c, "rc_batch_run");
  tt[4] = now_s();
  if (rc_batch_download(batch) != RC_OK) shard_die(sh->c, "rc_batch_download");
  tt[5] = now_s();
  if (getenv("RNACODE_CUDA_VERBOSE"))
    fprintf(stderr, "[RNAcode_b200]   device batch of %d units: create %.3f s, trees/thresholds %.3f s, upload %.3f s, run %.3f s, download %.3f s\n",
            sh->nunits, tt[1] - tt[0], tt[2] - tt[1], tt[3] - tt[2], tt[4] - tt[3], tt[5] - tt[4]);
  for (k = 0; k < sh->nunits; k++) {
    const unit_t *u = &sh->units[k];
    blk_t *b = &sh->blk[u->blk];
    if (u->ns > 0 && rc_batch_max_scores(batch, k, b->maxScores + u->s0) != RC_OK) shard_die(sh->c, "rc_batch_max_scores");
    if (u->want_native) {
      int cap = 256, rc;
      b->hss = (rc_hss *)malloc(sizeof(rc_hss) * cap);
      rc = rc_batch_native_hss(batch, k, b->hss, cap, &b->n_hss);
      if (rc == RC_ERR_CAPACITY) {
        cap = b->n_hss;
        b->hss = (rc_hss *)realloc(b->hss, sizeof(rc_hss) * cap);
        rc = rc_batch_native_hss(batch, k, b->hss, cap, &b->n_hss);
      }
      if (rc != RC_OK) shard_die(sh->c, "rc_batch_native_hss");
    }
  }
  rc_batch_destroy(batch);
  free(descs);
  return NULL;
}

/* The native alignment and the null alignments [s0, s0 + ns) of the listed blocks, on all devices.  Fills
 * blk[].maxScores[s0 .. s0+ns) and, when want_native, blk[].hss / blk[].n_hss.  Blocks are dealt out by cost
 * (DP cells ~ (N-1) * L^2 per alignment); a block that alone outweighs a device's fair share is cut along its
 * null alignments, which are independent too (only maxima are gathered). */
static void gpu_batch(blk_t *blk, const int *list, int nlist, int s0, int ns, const int *blosum, int want_native,
                      double *t_seeds, double *t_gpu) {
  const int mode = evolve_mode();
  double ta = now_s(), tb, total = 0, load[MAX_GPUS];
  int k, j, d, G, nunits = 0, pos[MAX_GPUS];
  shard_t sh[MAX_GPUS];
  unit_t *units, *sorted;
  pthread_t th[MAX_GPUS];
  init_gpus();
  G = g_ngpus;
  /* one CreateSeed() per null alignment, block by block in input order (src/treeSimulate.c:84) */
  for (k = 0; k < nlist; k++) {
    blk_t *b = &blk[list[k]];
    b->seeds = (unsigned int *)malloc(sizeof(unsigned int) * (ns > 0 ? ns : 1));
    if (rc_wrap_set_block) rc_wrap_set_block(b->scored_idx);
    if (rc_wrap_set_sample) rc_wrap_set_sample(s0);
    for (j = 0; j < ns; j++) b->seeds[j] = (unsigned int)(CreateSeed() & 0xffffffffUL);
    total += (double)(b->N - 1) * b->L * b->L * (ns + 1);
  }
  units = (unit_t *)malloc(sizeof(unit_t) * ((size_t)nlist * G + 1));
  for (k = 0; k < nlist; k++) {
    blk_t *b = &blk[list[k]];
    const double per_aln = (double)(b->N - 1) * b->L * b->L;
    int parts = 1, p0;
    if (G > 1 && ns >= 2 * G && per_aln * (ns + 1) > total / (2.0 * G)) parts = G;
    for (p0 = 0; p0 < parts; p0++) {
      const int a = (int)((long long)ns * p0 / parts), e = (int)((long long)ns * (p0 + 1) / parts);
      units[nunits].blk = list[k];
      units[nunits].s0 = s0 + a;
      units[nunits].ns = e - a;
      units[nunits].want_native = want_native && p0 == 0;
      units[nunits].cost = per_aln * (e - a + 1);
      nunits++;
    }
  }
  /* heaviest first, each to the device with the least work so far */
  sorted = (unit_t *)malloc(sizeof(unit_t) * (nunits + 1));
  memcpy(sorted, units, sizeof(unit_t) * nunits);
  for (k = 1; k < nunits; k++) { /* insertion sort by descending cost (windows hold at most a few thousand units) */
    unit_t u = sorted[k];
    for (j = k - 1; j >= 0 && sorted[j].cost < u.cost; j--) sorted[j + 1] = sorted[j];
    sorted[j + 1] = u;
  }
  for (d = 0; d < G; d++) {
    memset(&sh[d], 0, sizeof(shard_t));
    sh[d].c = g_ctxs[d];
    sh[d].blk = blk;
    sh[d].units = (unit_t *)malloc(sizeof(unit_t) * (nunits + 1));
    sh[d].seed0 = s0;
    sh[d].mode = mode;
    sh[d].blosum = blosum;
    load[d] = 0;
    pos[d] = 0;
  }
  for (k = 0; k < nunits; k++) {
    int best = 0;
    for (d = 1; d < G; d++)
      if (load[d] < load[best]) best = d;
    sh[best].units[sh[best].nunits++] = sorted[k];
    load[best] += sorted[k].cost;
  }
  (void)pos;
  tb = now_s();
  *t_seeds += tb - ta;
  for (d = 1; d < G; d++)
    if (pthread_create(&th[d], NULL, run_shard, &sh[d]) != 0) nrerror("ERROR: pthread_create failed.\n");
  run_shard(&sh[0]);
  for (d = 1; d < G; d++) pthread_join(th[d], NULL);
  for (d = 0; d < G; d++) free(sh[d].units);
  free(units);
  free(sorted);
  for (k = 0; k < nlist; k++) free(blk[list[k]].seeds);
  *t_gpu += now_s() - tb;
}

static float ***g_sk_fwd_tag[1], ***g_sk_rev_tag[1]; /* stand-ins for Sk_native / Sk_native_rev: colorAln only passes them on */

static double g_wall0 = 0.0;
static int g_window_no = 0;

static void process_window(host_job *job, int *blosum) {
  blk_t *blk = job->blk;
  const int nb = job->nb;
  const int n = pars.sampleN > 0 ? pars.sampleN : 0;
  /* --stop-early: a first round of few null alignments per block decides most non-coding blocks
   * (src/score.c:1036-1042: more than cutoff*n of them beat the native score); only the others get the rest */
  const int n1 = (pars.stopEarly && n > 32) ? 32 : n;
  const int stopCutoff = (int)(pars.cutoff * pars.sampleN); /* src/score.c:992 */
  int *list, nok = 0, nlist2 = 0, i, k, j;
  double t0 = now_s(), t1, t3, t_seeds = 0, t_gpu = 0;
  const int verbose = getenv("RNACODE_CUDA_VERBOSE") != NULL;

  host_stage_finish(job);
  t1 = now_s();

  list = (int *)malloc(sizeof(int) * (nb > 0 ? nb : 1));
  for (i = 0; i < nb; i++) {
    blk_t *b = &blk[i];
    if (!b->ok) continue;
    b->rows = (char *)malloc((size_t)b->N * b->cols);
    for (k = 0; k < b->N; k++) memcpy(b->rows + (size_t)k * b->cols, b->aln[k]->seq, b->cols);
    b->maxScores = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
    list[nok++] = i;
  }
  if (nok > 0) gpu_batch(blk, list, nok, 0, n1, blosum, 1, &t_seeds, &t_gpu);

  /* the reference's bookkeeping on the native result: sort, best score (src/RNAcode.c:173-178), then the first round's count */
  for (k = 0; k < nok; k++) {
    blk_t *b = &blk[list[k]];
    b->results = hss_to_segments((const struct aln **)b->aln, b->hss, b->n_hss);
    free(b->hss);
    b->hssCount = 0;
    while (b->results[b->hssCount++].score > 0.0);
    qsort(b->results, b->hssCount, sizeof(segmentStats), compareScores);
    b->maxScore = b->results[0].score;
    b->status = 1;
    b->better = 0;
    for (j = 0; j < n1 && b->status == 1; j++) {
      if ((float)b->maxScores[j] > b->maxScore) b->better++;
      if (pars.stopEarly && b->better > stopCutoff) b->status = -1;
    }
    if (b->status == 1 && n1 < n) list[nlist2++] = list[k];
  }
  if (nlist2 > 0) {
    /* (the library scores instance 0 = the native alignment of every unit again in this round and the result is dropped:
     * one alignment in n - 31 per block that is still undecided, accepted for the sake of one batch layout) */
    gpu_batch(blk, list, nlist2, n1, n - n1, blosum, 0, &t_seeds, &t_gpu);
    for (k = 0; k < nlist2; k++) {
      blk_t *b = &blk[list[k]];
      for (j = n1; j < n && b->status == 1; j++) {
        if ((float)b->maxScores[j] > b->maxScore) b->better++;
        if (pars.stopEarly && b->better > stopCutoff) b->status = -1;
      }
    }
  }
  t3 = now_s();

  /* the reference's fit and reporting, in input order */
  for (i = 0; i < nb; i++) {
    blk_t *b = &blk[i];
    float parMu = 0, parLambda = 0;
    double mu, lambda;
    if (!b->ok) {
      fprintf(stderr, "\nSkipping alignment. Failed to build ML tree.\n");
      freeAln(b->aln);
      free(b->aln);
      continue;
    }
    if (b->status == 1) { /* src/score.c:1050-1062 */
      if (EVDMaxLikelyFit(b->maxScores, NULL, n, &mu, &lambda) == 1) {
        parMu = mu;
        parLambda = lambda;
      } else {
        b->status = -1;
      }
    }
    for (j = 0; j < b->hssCount; j++)
      b->results[j].pvalue = b->status == 1 ? 1 - exp((-1) * exp((-1) * parLambda * (b->results[j].score - parMu))) : 99.0;
    g_bt_scores_fwd = b->sf; /* --eps: backtrack() rows of this block come from rc_pair_rows (rnacode_cuda_host.h) */
    g_bt_scores_rev = b->sr;
    g_bt_blosum = blosum;
    printResults(pars.outputFile, pars.outputFormat, (const struct aln **)b->aln, b->results);
    freeResults(b->results);
    free(b->maxScores);
    free(b->rows);
    free(b->sf);
    free(b->sr);
    free(b->tpar);
    free(b->trow);
    free(b->tcum);
    freeAln(b->aln);
    free(b->aln);
  }
  free(list);
  if (verbose)
    fprintf(stderr,
            "[RNAcode_b200] window %d of %d blocks (%d scored, %d in the second sampling round): workers forked at %.3f s, done by "
            "%.3f s (waited %.3f s for them), batch set-up+seeds %.3f s, GPU upload+run+download %.3f s (from %.3f s), fit+report "
            "%.3f s, window done at %.3f s\n",
            g_window_no, nb, nok, nlist2, job->t_start - g_wall0, job->t_done - g_wall0, t1 - t0, t_seeds, t_gpu, t1 - g_wall0,
            now_s() - t3, now_s() - g_wall0);
  g_window_no++;
}

int main(int argc, char *argv[]) {
  int i, j, N, L, alnCounter = 0, nb = 0, win_blocks, blosum[576];
  long scored = 0;
  size_t win_bytes = 0, max_bytes;
  clock_t startTime;
  double wall0 = now_s();
  int (*readFunction)(FILE * clust, struct aln * alignedSeqs[]) = NULL;
  struct aln *inputAln[MAX_NUM_NAMES];
  blk_t *blk, *blk_prev;
  host_job job;
  int have_job = 0;
  const char *e;
  rc_maf_map map = {NULL, NULL, NULL, 0};
  int mapped = 0;

  /* the reference's defaults, src/RNAcode.c:68-90 */
  pars.Delta = -10.0;
  pars.Omega = -4.0;
  pars.omega = -2.0;
  pars.stopPenalty_k = -8.0;
  pars.stopPenalty_0 = -9999.0;
  pars.inputFile = stdin;
  pars.outputFile = stdout;
  pars.debugFile = stdout;
  pars.bestOnly = 0;
  pars.bestRegion = 0;
  pars.stopEarly = 0;
  pars.postscript = 0;
  pars.postscript_cutoff = 0.05;
  strcpy(pars.postscriptDir, "eps");
  pars.sampleN = 100;
  pars.blosum = 62;
  strcpy(pars.limit, "");
  pars.cutoff = 1.0;
  pars.outputFormat = 0;
  strcpy(pars.debugFileName, "");
  strcpy(pars.inputFileName, "STDIN");

  read_commandline(argc, argv);
  srand(time(NULL));
  Sk_native = g_sk_fwd_tag; /* told apart by __wrap_backtrack, never dereferenced */
  Sk_native_rev = g_sk_rev_tag;

  ntMap['A'] = ntMap['a'] = 0;
  ntMap['C'] = ntMap['c'] = 1;
  ntMap['G'] = ntMap['g'] = 2;
  ntMap['T'] = ntMap['t'] = 3;
  ntMap['U'] = ntMap['u'] = 3;

  switch (checkFormat(pars.inputFile)) {
    case CLUSTAL: readFunction = &read_clustal; break;
    case MAF: readFunction = &read_maf; break;
    default: nrerror("ERROR: Unknown alignment file format. Use Clustal W or MAF format.\n");
  }
  {
    int **mat = getScoringMatrix(); /* what getModels puts into bgModel.matrix (src/score.c:50-75) */
    for (i = 0; i < 24; i++)
      for (j = 0; j < 24; j++) blosum[i * 24 + j] = mat[i][j];
    for (i = 0; i < 24; i++) free(mat[i]);
    free(mat);
  }
  e = getenv("RNACODE_CUDA_WINDOW");
  win_blocks = e ? atoi(e) : 4096;
  if (win_blocks < 1) win_blocks = 1;
  e = getenv("RNACODE_CUDA_WINDOW_MB"); /* bound on the bytes of simulated alignments per window */
  max_bytes = (size_t)(e ? atol(e) : 4096) << 20;
  blk = (blk_t *)calloc(win_blocks, sizeof(blk_t));
  blk_prev = (blk_t *)calloc(win_blocks, sizeof(blk_t));
  startTime = clock();
  g_wall0 = wall0;

  /* (f3) MAF files are parsed in place from a mapping of the file (rnacode_maf_mmap.h); pipes, Clustal input and
   * RNACODE_CUDA_PARSER=reference keep the reference's read functions */
  if (readFunction == &read_maf && !((e = getenv("RNACODE_CUDA_PARSER")) && strcmp(e, "reference") == 0))
    mapped = rc_maf_map_open(pars.inputFile, &map);
  while ((mapped ? rc_read_maf_mapped(&map, inputAln) : readFunction(pars.inputFile, inputAln)) != 0) {
    alnCounter++;
    for (i = 0; inputAln[i] != NULL; i++)
      for (j = 0; inputAln[i]->seq[j]; j++) inputAln[i]->seq[j] = toupper(inputAln[i]->seq[j]);
    if (strcmp(pars.limit, "") != 0) pruneAln(pars.limit, (struct aln **)inputAln);
    L = getSeqLength(inputAln[0]->seq);
    for (N = 0; inputAln[N] != NULL; N++);
    if (N <= 2) { /* src/RNAcode.c:142-145 */
      fprintf(stderr, "Skipping alignment. There must be at least three sequences in the alignment.\n");
      continue;
    }
    if (L < 3) {
      fprintf(stderr, "Skipping alignment. Too short.\n");
      continue;
    }
    memset(&blk[nb], 0, sizeof(blk_t));
    blk[nb].aln = (struct aln **)malloc(sizeof(struct aln *) * (N + 1));
    memcpy(blk[nb].aln, inputAln, sizeof(struct aln *) * (N + 1));
    blk[nb].N = N;
    blk[nb].L = L;
    blk[nb].cols = (int)strlen(inputAln[0]->seq);
    blk[nb].scored_idx = scored++;
    win_bytes += (size_t)N * blk[nb].cols * (size_t)(pars.sampleN + 1) * 2;
    nb++;
    if (nb == win_blocks || win_bytes >= max_bytes) {
      /* fork the workers of this window, then finish the previous one (its workers have been running since it was parsed) */
      host_job next_job;
      blk_t *tmp;
      host_stage_start(&next_job, blk, nb);
      if (have_job) process_window(&job, blosum);
      job = next_job;
      have_job = 1;
      tmp = blk_prev;
      blk_prev = blk;
      blk = tmp;
      nb = 0;
      win_bytes = 0;
    }
  }
  if (nb > 0) {
    host_job next_job;
    host_stage_start(&next_job, blk, nb);
    if (have_job) process_window(&job, blosum);
    job = next_job;
    have_job = 1;
  }
  if (have_job) process_window(&job, blosum);

  if (pars.outputFormat == 0) {
    float runtime = (float)(clock() - startTime) / CLOCKS_PER_SEC;
    fprintf(pars.outputFile,
            "\n%i alignment(s) scored in %.2f seconds. Parameters used:\nN=%i, Delta=%.2f, Omega=%.2f, omega=%.2f, stop penalty=%.2f\n\n",
            alnCounter, runtime, pars.sampleN, pars.Delta, pars.Omega, pars.omega, pars.stopPenalty_k);
  }
  if (getenv("RNACODE_CUDA_VERBOSE"))
    fprintf(stderr, "[RNAcode_b200] %d alignments in %.3f s wall (%.1f blocks/s)\n", alnCounter, now_s() - wall0,
            alnCounter / (now_s() - wall0));
  free(blk);
  free(blk_prev);
  exit(EXIT_SUCCESS);
}
