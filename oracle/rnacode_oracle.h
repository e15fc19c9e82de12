/* oracle/rnacode_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of RNAcode's scoring hot path (reference v0.3.1, /root/reference).  It is the
 * checker for the CUDA library in rnacode_b200/csrc; nothing in the product may include, link or
 * call it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it.
 *
 * Parity pin: the reference ships no tests or golden vectors for this path, so the
 * restatement is pinned against outputs of the reference itself, compiled unmodified by
 * oracle/Makefile (target `ref`) and dumped by oracle/_ref/ref_probe into tests/golden/ as json
 * (tests/test_oracle_golden.py).
 */
#ifndef RNACODE_ORACLE_H
#define RNACODE_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* pars.Delta/Omega/omega/stopPenalty_0/stopPenalty_k  (reference src/RNAcode.h:29-35, defaults src/RNAcode.c:68-72) */
typedef struct {
  float Delta, Omega, omega, stopPenalty_0, stopPenalty_k;
} orc_params;

/* subset of segmentStats (reference src/score.h:48-63) that the scoring core decides */
typedef struct {
  int strand; /* '+' or '-' */
  int frame;  /* 0..2 */
  int startSite, endSite; /* codon indices, 0-based, inclusive */
  float score;
} orc_hss;

void orc_default_params(orc_params *p);
/* standard genetic code in the reference's encoding (src/code.c:26-35): index a*16+b*4+c, A,C,G,T=0..3,
 * value = amino-acid index in BLOSUM order "ARNDCQEGHILKMFPSTWYV", stop = -1 */
const int *orc_transcode(void);
/* BLOSUM62 24x24 in NCBI order ARNDCQEGHILKMFPSTWYVBZX* (same data as src/code.c:39-62) */
const int *orc_blosum62(void);

/* Length of the ungapped reference row: getSeqLength, src/misc.c:272-289 */
int orc_seq_length(const char *row, int cols);

/* revAln, src/rnaz_utils.c:316-348: reverse every row and complement A,C,G,T,U only */
void orc_rev_aln(const char *rows, int N, int cols, char *out);

/* sigma[k][x], z[k][x] for one strand (rows already in that strand's orientation):
 * getBlock src/misc.c:186-246 + calculateSigma src/score.c:375-426 as driven by src/score.c:488-494.
 * sigma, z: arrays of N*(L+1); entries k>=1, x=3..L are written. */
void orc_sigma_z(const char *rows, int N, int cols, const float *scores /* N*4 */, const int *blosum /* 24*24 */,
                 const orc_params *p, float *sigma, int *z);

/* Row b (1-based start position) of the pairwise matrices Sk[k][state][b][i], src/score.c:496-535, the only part of
 * Sk that backtrack() (src/score.c:558-797) reads for the --eps plots.  out: N*3*(L+1) floats, out[(k*3+x)*(L+1)+i];
 * entries outside i = b-1, b+2, b+5, ... (and all of k = 0) are left 0. */
void orc_pair_row(const char *rows, int N, int cols, const float *scores, const int *blosum, const orc_params *p, int b,
                  float *out);

/* One full strand: getPairwiseScoreMatrix (src/score.c:441-556) + getMultipleScoreMatrix (:811-848) +
 * getHSS (:864-974) without materialising Sk / S.  If S_dense != NULL it receives S[b][i] at
 * S_dense[b*(L+1)+i] (only for small L, used to cross-check against the reference's matrix).
 * Returns the number of HSS appended to out (at most max_out are stored). */
int orc_score_strand(const char *rows, int N, int cols, const float *scores, const int *blosum, const orc_params *p,
                     int strand_char, orc_hss *out, int max_out, float *S_dense);

/* scoreAln, src/score.c:1067-1147: '+' strand with scores_fwd, then revAln and '-' strand with scores_rev */
int orc_score_aln(const char *rows, int N, int cols, const float *scores_fwd, const float *scores_rev,
                  const int *blosum, const orc_params *p, orc_hss *out, int max_out);

/* One iteration of the sampling loop, src/score.c:1008-1044, given the simulated rows in input order:
 * reintroduceGaps (src/misc.c:127-148) then scoreAln with the NATIVE models; returns the best HSS score
 * or -1.0 when the sample has no HSS. */
double orc_sample_max(const char *native_rows, const char *sample_rows, int N, int cols, const float *scores_fwd,
                      const float *scores_rev, const int *blosum, const orc_params *p);

/* countFreqsMono, src/score.c:255-280 */
void orc_count_freqs(const char *rows, int N, int cols, float freqs[4]);
/* probHKY (src/score.c:204-245) + calculateBG (src/score.c:107-193): expected scores per Hamming distance */
void orc_calculate_bg(float dist, const float freqs[4], float kappa, const int *blosum, float scores[4], float probs[4]);

/* ---- null-alignment simulation (kernel d): seq-gen's HKY / no-rate-heterogeneity path ---------------------- */
/* MT19937 as seqgen/twister.c:73-146 (init_genrand, genrand_int32) */
typedef struct {
  unsigned long mt[624];
  int mti;
} orc_mt;
void orc_mt_init(orc_mt *g, unsigned long seed);
unsigned long orc_mt_next(orc_mt *g);
/* One simulated alignment, as simulateTree + tree2aln + sortAln produce it (src/treeSimulate.c:52-97, :254-283,
 * src/misc.c:150-171; seqgen/evolve.c:167-199, :291-308, :400-433): nodes in the order EvolveSequences visits them
 * (root first, then branch1 subtree, branch2 subtree, and branch0 subtree of an unrooted root); every node draws one
 * genrand_real1 per site, in site order; the root picks its state from the cumulative frequencies, every other node
 * from the cumulative row (of its branch's transition matrix) of its parent's state.  row[node] >= 0 marks a tip and
 * names the alignment row it fills.  Output: N*cols characters over ACGT. */
void orc_evolve(unsigned long seed, int n_nodes, const int *parent, const int *row, const double *cum /* n_nodes*16 */,
                const double addFreq[4], int N, int cols, char *out_rows);

/* DP cell count for one alignment: 2*(N-1)*P(L) */
double orc_cells(int N, int L);

#ifdef __cplusplus
}
#endif
#endif
