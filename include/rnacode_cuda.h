/* include/rnacode_cuda.h -- C ABI of libRNAcode_cuda, the B200 (sm_100a) implementation of RNAcode's
 * scoring hot path.  Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * What it replaces in the reference (/root/reference, RNAcode v0.3.1):
 *
 *   segmentStats* scoreAln(const struct aln *alignment[], TTree*, float kappa, int backtrack)
 *       src/score.h:116, src/score.c:1067-1147, called from src/RNAcode.c:171      -> rc_score_aln()
 *       (its side effect Sk_native / Sk_native_rev, read by backtrack() for --eps    -> rc_pair_rows())
 *   the scoring half of  int getExtremeValuePars(...)  i.e. the n calls of scoreAln() on the null
 *   alignments and the max over their HSS,
 *       src/score.h:103-104, src/score.c:1004-1044, called from src/RNAcode.c:180  -> rc_score_samples()
 *   both at once for many alignment blocks (the reference processes one block per iteration of
 *   main()'s while loop, src/RNAcode.c:115-221)                                    -> rc_batch_*()
 *
 * Everything that stays host C in the reference also stays outside this library: alignment
 * parsing, PhyML tree + kappa (treeML), the background models (getModels -> scores[4] per species,
 * which arrive here as plain float tables), seq-gen simulation in exact mode (null alignments
 * arrive as bytes), the Gumbel fit (EVDMaxLikelyFit), the p-value formula and printing.
 *
 * Conventions
 *   - rows are N*cols bytes, row-major, NOT NUL-terminated, exactly the bytes scoreAln() would see
 *     (main() upper-cases them first, src/RNAcode.c:121-128).  Row 0 is the reference sequence.
 *   - scores_fwd / scores_rev are models[k].scores[0..3] / modelsRev[k].scores[0..3]
 *     (bgModel, src/score.h:34-44), N*4 floats each; entry k=0 is unused.
 *   - blosum is the 24x24 int matrix of bgModel.matrix (src/code.c:39-88), row-major.
 *   - null alignments ("samples") are n*N*cols bytes, sample-major, rows already permuted into
 *     input order (sortAln, src/misc.c:150-171).  The native gap pattern is re-imposed by the
 *     library (reintroduceGaps, src/misc.c:127-148), so callers may pass either form.
 *   - every entry point returns RC_OK (0) or a negative rc_status; the library never exits or
 *     prints.  rc_last_error() gives a message for the last failure on that context.
 *   - there is NO CPU fallback: without a CUDA device rc_create() fails with RC_ERR_CUDA.
 */
#ifndef RNACODE_CUDA_H
#define RNACODE_CUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  RC_OK = 0,
  RC_ERR_ARG = -1,      /* bad argument (N<2, cols<1, NULL pointer, ...) */
  RC_ERR_CUDA = -2,     /* CUDA runtime failure, see rc_last_error() */
  RC_ERR_NOMEM = -3,    /* host or device allocation failed */
  RC_ERR_CAPACITY = -4, /* caller-provided output buffer too small (n_hss still reports the need) */
  RC_ERR_STATE = -5     /* call sequence violated (e.g. results requested before run) */
} rc_status;

/* pars.Delta, .Omega, .omega, .stopPenalty_0, .stopPenalty_k  (src/RNAcode.h:29-35; defaults
 * src/RNAcode.c:68-72: -10, -4, -2, -9999, -8) */
typedef struct {
  float Delta, Omega, omega, stopPenalty_0, stopPenalty_k;
} rc_params;

/* The fields of segmentStats (src/score.h:48-63) decided by the scoring core.  The caller derives
 * start/end/startGenomic/endGenomic/name exactly as src/score.c:914-938 does. */
typedef struct {
  int strand; /* '+' or '-' */
  int frame;  /* 0,1,2 */
  int startSite, endSite; /* codon indices, 0-based, inclusive */
  float score;
} rc_hss;

/* One alignment block and (optionally) its null alignments. */
typedef struct {
  int N;                   /* rows (species), >= 2 */
  int cols;                /* alignment columns, >= 1 */
  const char *rows;        /* N*cols bytes */
  const float *scores_fwd; /* N*4 */
  const float *scores_rev; /* N*4 */
  int n_samples;           /* 0 if none */
  const char *samples;     /* n_samples*N*cols bytes or NULL */
} rc_block_desc;

/* A phylogenetic tree flattened in the order seq-gen evolves it (EvolveSequences / EvolveNode,
 * seqgen/evolve.c:400-433): the root first, then the subtree of branch1, of branch2 and, for the unrooted
 * trees PhyML writes, of branch0.  Input of the on-GPU null-alignment simulation that replaces
 * simulateTree + tree2aln + sortAln (src/treeSimulate.c:52-97, :254-283, src/misc.c:150-171). */
typedef struct {
  int n_nodes;
  const int *parent; /* [n_nodes]; -1 for the root */
  const int *row;    /* [n_nodes]; for a tip the alignment row (input order) it fills, -1 for internal nodes */
  const double *cum; /* [n_nodes][16]: the cumulative transition matrix seq-gen's SetMatrix(matrix, length0) gives
                        for the branch above the node (seqgen/nucmodels.c:187-196, :318-362); for node 0 the first
                        four entries are the cumulative root frequencies addFreq[0..3] (seqgen/model.c:116-119) */
} rc_tree_desc;

enum {
  RC_RNG_MT19937 = 0, /* bit-exact seq-gen: MT19937 seeded per sample, consumed in seq-gen's order (exact mode) */
  RC_RNG_PHILOX = 1   /* counter-based Philox4x32-10: same distribution, different stream (GPU-RNG mode) */
};

typedef struct rc_ctx rc_ctx;
typedef struct rc_batch rc_batch;

/* -- context ---------------------------------------------------------------------------------- */
int rc_create(rc_ctx **ctx, int device);
void rc_destroy(rc_ctx *ctx);
const char *rc_last_error(const rc_ctx *ctx);
void rc_default_params(rc_params *p);
/* Use an existing CUDA stream (a cudaStream_t passed as void*) for all work of this context; NULL
 * restores the context's own stream. */
int rc_set_stream(rc_ctx *ctx, void *cuda_stream);
/* Tunables / test hooks: "force_dense" (0/1: route every alignment through the dense-S fallback),
 * "band_slots" (1..3: tie-band slots per row record before the dense fallback is taken),
 * "scratch_mb" (device scratch budget per chunk), and switches that only choose between kernels with identical results
 * (DESIGN.md section 4): "no_smp", "no_smps", "no_chain", "no_fused", "no_fold", "no_sig_p2", "no_sig_rows3", "no_allf" (0/1), "reg_tu" (-1 auto / 0 / 1), "tail_max" (0..31),
 * "reg_max_nk" (12..16), "smps_max_sites", "smpc_max_sites", "hss_thr_tasks".  Unknown keys are an error. */
int rc_set_option(rc_ctx *ctx, const char *key, long value);

/* -- one block at a time (same call shape as the reference) ----------------------------------- */
/* scoreAln(): native alignment, both strands.  HSS are returned '+' strand first, then '-', each in
 * frame order and in order of discovery, like src/score.c:1107-1127.  *n_hss is always the full count. */
int rc_score_aln(rc_ctx *ctx, const rc_block_desc *block, const rc_params *params, const int *blosum, rc_hss *out,
                 int max_hss, int *n_hss);
/* Rows of the pairwise matrices of the native alignment, for backtrack() (src/score.h:111, src/score.c:558-797; called from
 * colorAln, src/postscript.c:303-305, when --eps plots are drawn).  The reference keeps dense copies Sk_native /
 * Sk_native_rev (3*N*(L+1)^2 floats each, copySk src/score.c:1084-1100) although backtrack(b, i, ...) only reads row b:
 * this call computes the requested rows on the GPU instead.  strand: 0 = '+', 1 = '-' (the alignment is reversed and
 * complemented as revAln does, and scores_rev is used).  b[r]: 1-based start position in that strand's coordinates.
 * out: n_rows * N * 3 * (L+1) floats, out[((r*N + k)*3 + state)*(L+1) + i] = Sk[k][state][b[r]][i] for
 * i = b-1 (0, src/score.c:500-504), b+2, b+5, ...; every other entry (and all of k = 0) is 0. */
int rc_pair_rows(rc_ctx *ctx, const rc_block_desc *block, const rc_params *params, const int *blosum, int strand, int n_rows,
                 const int *b, float *out);

/* maxScores[i] of src/score.c:1044 for every null alignment i: best HSS score over both strands or
 * -1.0 when the sample has none. */
int rc_score_samples(rc_ctx *ctx, const rc_block_desc *block, const rc_params *params, const int *blosum,
                     double *max_scores);

/* Same, with the null alignments simulated on the GPU (block->samples is ignored, block->n_samples of them are drawn). */
int rc_score_samples_evolve(rc_ctx *ctx, const rc_block_desc *block, const rc_tree_desc *tree, const unsigned int *seeds,
                            int rng, const rc_params *params, const int *blosum, double *max_scores);
/* Debug / test access: the simulated rows of sample i of a block (N*cols characters) after rc_batch_run. */
int rc_batch_get_sample_rows(rc_batch *batch, int block, int sample, char *rows);

/* -- many blocks at once ------------------------------------------------------------------------ */
/* The descriptors (and the host memory they point to) must stay valid until rc_batch_upload() returns. */
int rc_batch_create(rc_ctx *ctx, const rc_block_desc *blocks, int n_blocks, const rc_params *params, const int *blosum,
                    rc_batch **batch);
/* Draw the block's n_samples null alignments on the GPU instead of taking them from desc.samples (which may
 * then be NULL).  seeds: one per sample (what SetSeed() would receive, low 32 bits).  Call before rc_batch_upload. */
int rc_batch_set_evolve(rc_batch *batch, int block, const rc_tree_desc *tree, const unsigned int *seeds, int rng);
/* The same for n consecutive blocks first .. first+n-1 in one call (a window of thousands of short blocks: the batched CLI
 * and the benchmark set every block's tree): trees[i] and seeds[i] (n_samples of that block) belong to block first+i. */
int rc_batch_set_evolve_many(rc_batch *batch, int first, int n, const rc_tree_desc *trees, const unsigned int *const *seeds,
                             int rng);
int rc_batch_upload(rc_batch *batch);   /* host -> device copies of rows, samples and score tables */
int rc_batch_run(rc_batch *batch);      /* all kernels; inputs and outputs stay in HBM */
int rc_batch_download(rc_batch *batch); /* device -> host copy of HSS records and per-sample maxima; synchronises */
int rc_batch_native_hss(rc_batch *batch, int block, rc_hss *out, int max_hss, int *n_hss);
int rc_batch_max_scores(rc_batch *batch, int block, double *max_scores /* n_samples of that block */);
/* maxScores of every block, block after block; n_out must be the sum of the blocks' n_samples */
int rc_batch_max_scores_all(rc_batch *batch, double *max_scores, size_t n_out);
void rc_batch_destroy(rc_batch *batch);

/* -- introspection for benchmarks and tests ------------------------------------------------------- */
typedef struct {
  double cells;        /* algorithmic DP cells of the batch: sum (n+1)*2*(N-1)*P(L) */
  long long launches;  /* kernels launched by the last rc_batch_run() */
  long long dense_fallbacks; /* alignments re-scored through the dense-S path in the last run */
  float ms_pack, ms_sigma, ms_dp, ms_hss; /* CUDA-event time of each stage in the last run */
  long long dp_launches;
  size_t h2d_bytes, d2h_bytes; /* bytes moved by upload / download */
  size_t device_bytes;         /* device memory held by the batch */
  float ms_pack_kernel;        /* k_pack alone (ms_pack also covers k_evolve and k_prep) */
  double pack_chars;           /* characters k_pack classifies per run: sum over blocks of instances*N*cols (padded to 16) */
} rc_batch_stats;
int rc_batch_get_stats(rc_batch *batch, rc_batch_stats *stats);

/* Runs a register-only kernel with the DP fast path's instruction mix (4 FADD : 1 FMNMX3) and returns the
 * achieved lane-operations per second: the practical FP32/ALU issue ceiling at the device's current clocks. */
int rc_calibrate_issue(rc_ctx *ctx, double *lane_ops_per_s);

/* Number of CUDA devices visible to the process (0 and RC_ERR_CUDA when there is none). */
int rc_device_count(int *count);

/* Library / build identification, e.g. "libRNAcode_cuda 0.1 sm_100a". */
const char *rc_version(void);

#ifdef __cplusplus
}
#endif
#endif
