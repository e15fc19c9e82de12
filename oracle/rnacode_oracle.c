/* oracle/rnacode_oracle.c -- TEST INFRASTRUCTURE ONLY (see rnacode_oracle.h).
 *
 * Streaming CPU restatement of the RNAcode scoring hot path, O(N*L) memory, same float32 operation
 * order as the reference so that results are bit-identical (pinned by tests/test_oracle_golden.py
 * against dumps of the compiled reference).  All reference citations are relative to /root/reference.
 * Compile with -ffp-contract=off (no FMA contraction) on a target with FLT_EVAL_METHOD == 0.
 */
#include "rnacode_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* the reference's comparison macros (src/score.h:28-29): strict '>' keeps the second argument on ties */
#define ORC_MAX(x, y) (((x) > (y)) ? (x) : (y))
#define ORC_MAX3(x, y, z) (ORC_MAX((ORC_MAX((x), (y))), (z)))

void orc_default_params(orc_params *p) {
  /* src/RNAcode.c:68-72 */
  p->Delta = -10.0f;
  p->Omega = -4.0f;
  p->omega = -2.0f;
  p->stopPenalty_k = -8.0f;
  p->stopPenalty_0 = -9999.0f;
}

/* ---- tables ------------------------------------------------------------------------------- */

static int g_transcode[64];
static int g_transcode_ready = 0;

const int *orc_transcode(void) {
  if (!g_transcode_ready) {
    /* standard genetic code, codons enumerated with bases in T,C,A,G order */
    static const char *tcag_aa = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    static const char *aa_order = "ARNDCQEGHILKMFPSTWYV"; /* BLOSUM row order */
    static const int tcag_to_acgt[4] = {3, 1, 0, 2};      /* T,C,A,G -> index in A,C,G,T */
    for (int i = 0; i < 64; i++) {
      int b1 = tcag_to_acgt[i / 16], b2 = tcag_to_acgt[(i / 4) % 4], b3 = tcag_to_acgt[i % 4];
      char aa = tcag_aa[i];
      int idx = -1;
      if (aa != '*') idx = (int)(strchr(aa_order, aa) - aa_order);
      g_transcode[b1 * 16 + b2 * 4 + b3] = idx;
    }
    g_transcode_ready = 1;
  }
  return g_transcode;
}

/* NCBI BLOSUM62, order A R N D C Q E G H I L K M F P S T W Y V B Z X * */
static const int g_blosum62[24 * 24] = {
    4,  -1, -2, -2, 0,  -1, -1, 0,  -2, -1, -1, -1, -1, -2, -1, 1,  0,  -3, -2, 0,  -2, -1, 0,  -4, /* A */
    -1, 5,  0,  -2, -3, 1,  0,  -2, 0,  -3, -2, 2,  -1, -3, -2, -1, -1, -3, -2, -3, -1, 0,  -1, -4, /* R */
    -2, 0,  6,  1,  -3, 0,  0,  0,  1,  -3, -3, 0,  -2, -3, -2, 1,  0,  -4, -2, -3, 3,  0,  -1, -4, /* N */
    -2, -2, 1,  6,  -3, 0,  2,  -1, -1, -3, -4, -1, -3, -3, -1, 0,  -1, -4, -3, -3, 4,  1,  -1, -4, /* D */
    0,  -3, -3, -3, 9,  -3, -4, -3, -3, -1, -1, -3, -1, -2, -3, -1, -1, -2, -2, -1, -3, -3, -2, -4, /* C */
    -1, 1,  0,  0,  -3, 5,  2,  -2, 0,  -3, -2, 1,  0,  -3, -1, 0,  -1, -2, -1, -2, 0,  3,  -1, -4, /* Q */
    -1, 0,  0,  2,  -4, 2,  5,  -2, 0,  -3, -3, 1,  -2, -3, -1, 0,  -1, -3, -2, -2, 1,  4,  -1, -4, /* E */
    0,  -2, 0,  -1, -3, -2, -2, 6,  -2, -4, -4, -2, -3, -3, -2, 0,  -2, -2, -3, -3, -1, -2, -1, -4, /* G */
    -2, 0,  1,  -1, -3, 0,  0,  -2, 8,  -3, -3, -1, -2, -1, -2, -1, -2, -2, 2,  -3, 0,  0,  -1, -4, /* H */
    -1, -3, -3, -3, -1, -3, -3, -4, -3, 4,  2,  -3, 1,  0,  -3, -2, -1, -3, -1, 3,  -3, -3, -1, -4, /* I */
    -1, -2, -3, -4, -1, -2, -3, -4, -3, 2,  4,  -2, 2,  0,  -3, -2, -1, -2, -1, 1,  -4, -3, -1, -4, /* L */
    -1, 2,  0,  -1, -3, 1,  1,  -2, -1, -3, -2, 5,  -1, -3, -1, 0,  -1, -3, -2, -2, 0,  1,  -1, -4, /* K */
    -1, -1, -2, -3, -1, 0,  -2, -3, -2, 1,  2,  -1, 5,  0,  -2, -1, -1, -1, -1, 1,  -3, -1, -1, -4, /* M */
    -2, -3, -3, -3, -2, -3, -3, -3, -1, 0,  0,  -3, 0,  6,  -4, -2, -2, 1,  3,  -1, -3, -3, -1, -4, /* F */
    -1, -2, -2, -1, -3, -1, -1, -2, -2, -3, -3, -1, -2, -4, 7,  -1, -1, -4, -3, -2, -2, -1, -2, -4, /* P */
    1,  -1, 1,  0,  -1, 0,  0,  0,  -1, -2, -2, 0,  -1, -2, -1, 4,  1,  -3, -2, -2, 0,  0,  0,  -4, /* S */
    0,  -1, 0,  -1, -1, -1, -1, -2, -2, -1, -1, -1, -1, -2, -1, 1,  5,  -2, -2, 0,  -1, -1, 0,  -4, /* T */
    -3, -3, -4, -4, -2, -2, -3, -2, -2, -3, -2, -3, -1, 1,  -4, -3, -2, 11, 2,  -3, -4, -3, -2, -4, /* W */
    -2, -2, -2, -3, -2, -1, -2, -3, 2,  -1, -1, -2, -1, 3,  -3, -2, -2, 2,  7,  -1, -3, -2, -1, -4, /* Y */
    0,  -3, -3, -3, -1, -2, -2, -3, -3, 3,  1,  -2, 1,  -1, -2, -2, 0,  -3, -1, 4,  -3, -2, -1, -4, /* V */
    -2, -1, 3,  4,  -3, 0,  1,  -1, 0,  -3, -4, 0,  -3, -3, -2, 0,  -1, -4, -3, -3, 4,  1,  -1, -4, /* B */
    -1, 0,  0,  1,  -3, 3,  4,  -2, 0,  -3, -3, 1,  -1, -3, -1, 0,  -1, -3, -2, -2, 1,  4,  -1, -4, /* Z */
    0,  -1, -1, -1, -2, -1, -1, -1, -1, -1, -1, -1, -1, -1, -2, 0,  0,  -2, -1, -1, -1, -1, -1, -4, /* X */
    -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, 1,  /* * */
};

const int *orc_blosum62(void) { return g_blosum62; }

/* ntMap: zero-initialised global, only A,C,G,T,U (both cases) set (src/score.c:41, src/RNAcode.c:94-98):
 * every other byte, including '-' and 'N', encodes as 0 ('A'). */
static inline int nt_code(unsigned char c) {
  switch (c) {
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    default: return 0;
  }
}

int orc_seq_length(const char *row, int cols) {
  int n = 0;
  for (int c = 0; c < cols; c++)
    if (row[c] != '-') n++;
  return n;
}

void orc_rev_aln(const char *rows, int N, int cols, char *out) {
  for (int k = 0; k < N; k++) {
    const char *src = rows + (size_t)k * cols;
    char *dst = out + (size_t)k * cols;
    for (int j = cols - 1; j >= 0; j--) {
      char letter = src[j];
      switch (letter) { /* src/rnaz_utils.c:327-333 */
        case 'T': letter = 'A'; break;
        case 'U': letter = 'A'; break;
        case 'C': letter = 'G'; break;
        case 'G': letter = 'C'; break;
        case 'A': letter = 'T'; break;
        default: break;
      }
      dst[cols - j - 1] = letter;
    }
  }
}

/* ---- sigma / z ------------------------------------------------------------------------------ */

void orc_sigma_z(const char *rows, int N, int cols, const float *scores, const int *blosum, const orc_params *p,
                 float *sigma, int *z) {
  const int *tc = orc_transcode();
  const char *seq0 = rows;
  int L = orc_seq_length(seq0, cols);
  /* map_0[l] = 1-based column of the l-th non-gap character of row 0 (pos2col, src/misc.c:250-269) */
  int *map0 = (int *)malloc(sizeof(int) * (size_t)(L + 2));
  {
    int l = 0;
    for (int c = 0; c < cols; c++)
      if (seq0[c] != '-') map0[++l] = c + 1;
  }
  for (int k = 1; k < N; k++) {
    const char *seqk = rows + (size_t)k * cols;
    for (int x = 3; x <= L; x++) {
      /* getBlock, src/misc.c:198-207: columns start..end (1-based), from column 1 when x==3 */
      int start = (x > 3) ? map0[x - 3] + 1 : 1;
      int end = map0[x];
      int gap0 = 0, gapk = 0;
      char codonA[3] = {'X', 'X', 'X'}, codonB[3] = {'X', 'X', 'X'};
      int j = 0;
      for (int c = start; c <= end; c++) {
        char a = seq0[c - 1], b = seqk[c - 1];
        if (a == '-') gap0++;
        if (b == '-') gapk++;
        if (a != '-') { /* calculateSigma, src/score.c:384-391 */
          if (j < 3) {
            codonA[j] = a;
            codonB[j] = b;
          }
          j++;
        }
      }
      int diff = gapk - gap0;
      if (diff < 0) diff = -diff;
      int zz = (diff % 3 == 0) ? 0 : ((diff % 3 == 1) ? +1 : -1); /* src/misc.c:230-244 */
      z[(size_t)k * (L + 1) + x] = zz;

      float s;
      if (codonB[0] == 'X' && codonB[1] == 'X' && codonB[2] == 'X') { /* src/score.c:394-396 */
        s = 0.0f;
      } else if (codonA[0] == 'N' || codonA[1] == 'N' || codonA[2] == 'N' || codonB[0] == 'N' || codonB[1] == 'N' ||
                 codonB[2] == 'N') { /* :400-404 */
        s = 0.0f;
      } else {
        int a1 = nt_code(codonA[0]), a2 = nt_code(codonA[1]), a3 = nt_code(codonA[2]);
        int b1 = nt_code(codonB[0]), b2 = nt_code(codonB[1]), b3 = nt_code(codonB[2]);
        int h = (a1 != b1) + (a2 != b2) + (a3 != b3); /* hDist, src/misc.c:303-313 */
        if (h == 0) {
          s = 0.0f; /* :409, tested before the stop tests */
        } else {
          int pepA = tc[a1 * 16 + a2 * 4 + a3], pepB = tc[b1 * 16 + b2 * 4 + b3];
          if (pepA == -1)
            s = p->stopPenalty_0; /* :414-416 */
          else if (pepB == -1)
            s = p->stopPenalty_k; /* :418-420 */
          else {
            float expected = scores[k * 4 + h];
            float observed = (float)blosum[pepA * 24 + pepB];
            s = observed - expected; /* :422-425 */
          }
        }
      }
      sigma[(size_t)k * (L + 1) + x] = s;
    }
  }
  free(map0);
}

/* ---- HSS state machine (getHSS, src/score.c:880-963), fed one entry at a time ------------------ */

typedef struct {
  float currMax;
  int segStart, segEnd;
} hss_machine;

static void hss_reset(hss_machine *m) {
  m->currMax = 0.0f;
  m->segStart = -1;
  m->segEnd = -1;
}

/* returns 1 and fills *emit when a segment is reported */
static int hss_feed(hss_machine *m, int i, int j, float e, int last, int frame, int strand, orc_hss *emit) {
  int emitted = 0;
  if (e > 0.0 || last) {
    if ((m->currMax > 0.0 && m->segEnd < i) || last) {
      if (m->segEnd - m->segStart >= 2) { /* minSegmentLength, :873, :902 */
        emit->strand = strand;
        emit->frame = frame;
        emit->startSite = m->segStart;
        emit->endSite = m->segEnd;
        emit->score = m->currMax;
        emitted = 1;
      }
      m->currMax = e;
      m->segStart = i;
      m->segEnd = j;
    } else {
      /* :953-954; fabs() is the double version applied to a float difference, 0.0001 is a double */
      if (e > m->currMax || ((fabs(e - m->currMax) < 0.0001) && ((j - i) >= (m->segEnd - m->segStart)))) {
        m->currMax = e;
        m->segStart = i;
        m->segEnd = j;
      }
    }
  }
  return emitted;
}

int orc_score_strand(const char *rows, int N, int cols, const float *scores, const int *blosum, const orc_params *p,
                     int strand_char, orc_hss *out, int max_out, float *S_dense) {
  int L = orc_seq_length(rows, cols);
  int count = 0;
  if (L < 3 || N < 2) return 0;
  float *sigma = (float *)calloc((size_t)N * (L + 1), sizeof(float));
  int *z = (int *)calloc((size_t)N * (L + 1), sizeof(int));
  float *st = (float *)malloc(sizeof(float) * 3 * (size_t)N);
  orc_sigma_z(rows, N, cols, scores, blosum, p, sigma, z);

  const float Delta = p->Delta, Omega = p->Omega, omega = p->omega;
  for (int frame = 0; frame <= 2; frame++) {
    int sites = (L - frame) / 3;
    hss_machine m;
    hss_reset(&m);
    for (int i = 0; i < sites; i++) {
      int b = i * 3 + 1 + frame;
      for (int k = 1; k < N; k++) st[3 * k] = st[3 * k + 1] = st[3 * k + 2] = 0.0f; /* src/score.c:500-504 */
      for (int j = i; j < sites; j++) {
        int x = j * 3 + 3 + frame;
        float sum = 0;
        for (int k = 1; k < N; k++) {
          float s0 = st[3 * k], s1 = st[3 * k + 1], s2 = st[3 * k + 2];
          float n0, n1, n2;
          int zz = z[(size_t)k * (L + 1) + x];
          if (zz == 0) { /* :506-510 */
            n0 = s0 + sigma[(size_t)k * (L + 1) + x];
            n1 = s1 + omega;
            n2 = s2 + omega;
          } else if (zz == +1) { /* :512-521 */
            n0 = ORC_MAX(s0 + Delta, s2 + Omega);
            n1 = ORC_MAX(s0 + Omega, s1 + Delta);
            n2 = ORC_MAX(s1 + Omega, s2 + Delta);
          } else { /* :523-533 */
            n0 = ORC_MAX(s0 + Delta, s1 + Omega);
            n1 = ORC_MAX(s1 + Delta, s2 + Omega);
            n2 = ORC_MAX(s2 + Delta, s0 + Omega);
          }
          st[3 * k] = n0;
          st[3 * k + 1] = n1;
          st[3 * k + 2] = n2;
          sum += ORC_MAX3(n0, n1, n2); /* :834-838 */
        }
        /* :841-843.  S[b][x-1] and S[b][x-2] are never written in row b and S is zero-filled. */
        float zero = 0.0f;
        float S = ORC_MAX3(sum, zero + Delta, zero + Delta) / (N - 1);
        if (S_dense) S_dense[(size_t)b * (L + 1) + x] = S;
        orc_hss e;
        if (hss_feed(&m, i, j, S, (i == sites - 1 && j == sites - 1), frame, strand_char, &e)) {
          if (count < max_out) out[count] = e;
          count++;
        }
      }
    }
  }
  free(sigma);
  free(z);
  free(st);
  return count;
}

void orc_pair_row(const char *rows, int N, int cols, const float *scores, const int *blosum, const orc_params *p, int b,
                  float *out) {
  int L = orc_seq_length(rows, cols);
  float *sigma = (float *)calloc((size_t)N * (L + 1), sizeof(float));
  int *z = (int *)calloc((size_t)N * (L + 1), sizeof(int));
  orc_sigma_z(rows, N, cols, scores, blosum, p, sigma, z);
  memset(out, 0, sizeof(float) * (size_t)N * 3 * (L + 1));
  const float Delta = p->Delta, Omega = p->Omega, omega = p->omega;
  for (int k = 1; k < N; k++) {
    float *r0 = out + ((size_t)k * 3 + 0) * (L + 1), *r1 = r0 + (L + 1), *r2 = r1 + (L + 1);
    for (int i = b + 2; i < L + 1; i += 3) {
      int zz = z[(size_t)k * (L + 1) + i];
      if (i - 3 < b) r0[i - 3] = r1[i - 3] = r2[i - 3] = 0.0f; /* src/score.c:500-504 */
      float s0 = r0[i - 3], s1 = r1[i - 3], s2 = r2[i - 3];
      if (zz == 0) { /* :506-510 */
        r0[i] = s0 + sigma[(size_t)k * (L + 1) + i];
        r1[i] = s1 + omega;
        r2[i] = s2 + omega;
      } else if (zz == +1) { /* :512-521 */
        r0[i] = ORC_MAX(s0 + Delta, s2 + Omega);
        r1[i] = ORC_MAX(s0 + Omega, s1 + Delta);
        r2[i] = ORC_MAX(s1 + Omega, s2 + Delta);
      } else { /* :523-533 */
        r0[i] = ORC_MAX(s0 + Delta, s1 + Omega);
        r1[i] = ORC_MAX(s1 + Delta, s2 + Omega);
        r2[i] = ORC_MAX(s2 + Delta, s0 + Omega);
      }
    }
  }
  free(sigma);
  free(z);
}

int orc_score_aln(const char *rows, int N, int cols, const float *scores_fwd, const float *scores_rev,
                  const int *blosum, const orc_params *p, orc_hss *out, int max_out) {
  int n = orc_score_strand(rows, N, cols, scores_fwd, blosum, p, '+', out, max_out, NULL);
  char *rev = (char *)malloc((size_t)N * cols);
  orc_rev_aln(rows, N, cols, rev);
  int room = max_out - n;
  if (room < 0) room = 0;
  n += orc_score_strand(rev, N, cols, scores_rev, blosum, p, '-', out + (n < max_out ? n : max_out), room, NULL);
  free(rev);
  return n;
}

double orc_sample_max(const char *native_rows, const char *sample_rows, int N, int cols, const float *scores_fwd,
                      const float *scores_rev, const int *blosum, const orc_params *p) {
  size_t sz = (size_t)N * cols;
  char *rows = (char *)malloc(sz);
  for (size_t i = 0; i < sz; i++) rows[i] = (native_rows[i] == '-') ? '-' : sample_rows[i]; /* src/misc.c:141-145 */
  int cap = 64, n;
  orc_hss *h = (orc_hss *)malloc(sizeof(orc_hss) * cap);
  n = orc_score_aln(rows, N, cols, scores_fwd, scores_rev, blosum, p, h, cap);
  if (n > cap) {
    cap = n;
    h = (orc_hss *)realloc(h, sizeof(orc_hss) * cap);
    n = orc_score_aln(rows, N, cols, scores_fwd, scores_rev, blosum, p, h, cap);
  }
  /* qsort descending + results[0].score (src/score.c:1034-1044); -1.0 sentinel when empty (:1129-1134) */
  float best = -1.0f;
  for (int i = 0; i < n; i++)
    if (h[i].score > best) best = h[i].score;
  free(h);
  free(rows);
  return (double)best;
}

/* ---- background model (host-side prep in the reference; restated for fixtures and the host mirror) ---- */

void orc_count_freqs(const char *rows, int N, int cols, float freqs[4]) {
  unsigned long counter = 0;
  for (int i = 0; i < 4; i++) freqs[i] = 0.0f;
  for (int k = 0; k < N; k++)
    for (int c = 0; c < cols; c++) {
      char ch = rows[(size_t)k * cols + c];
      if (ch == '-') continue;
      freqs[nt_code((unsigned char)ch)]++;
      counter++;
    }
  for (int i = 0; i < 4; i++) freqs[i] /= (float)counter;
}

static float prob_hky(int i, int j, float d, const float freqs[4], float kappa) {
  /* src/score.c:204-245 -- float variables, double literals: each right-hand side is evaluated in
   * double where a double operand appears and rounded to float on assignment. */
  float piA = freqs[0], piC = freqs[1], piG = freqs[2], piT = freqs[3];
  float piR = piA + piG, piY = piT + piC;
  float r = 1. / (2. * (piA * piC + piC * piG + piA * piT + piG * piT + kappa * (piC * piT + piA * piG)));
  float l = r * d;
  float k1 = kappa * piY + piR;
  float k2 = kappa * piR + piY;
  float exp1 = exp(-l);
  float exp22 = exp(-k2 * l);
  float exp21 = exp(-k1 * l);
  float result[4][4];
  result[0][0] = piA * (1. + (piY / piR) * exp1) + (piG / piR) * exp22;
  result[0][1] = piC * (1. - exp1);
  result[0][2] = piG * (1. + (piY / piR) * exp1) - (piG / piR) * exp22;
  result[0][3] = piT * (1. - exp1);
  result[1][0] = piA * (1. - exp1);
  result[1][1] = piC * (1. + (piR / piY) * exp1) + (piT / piY) * exp21;
  result[1][2] = piG * (1. - exp1);
  result[1][3] = piT * (1. + (piR / piY) * exp1) - (piT / piY) * exp21;
  result[2][0] = piA * (1. + (piY / piR) * exp1) - (piA / piR) * exp22;
  result[2][1] = piC * (1. - exp1);
  result[2][2] = piG * (1. + (piY / piR) * exp1) + (piA / piR) * exp22;
  result[2][3] = piT * (1. - exp1);
  result[3][0] = piA * (1. - exp1);
  result[3][1] = piC * (1. + (piR / piY) * exp1) - (piC / piY) * exp21;
  result[3][2] = piG * (1. - exp1);
  result[3][3] = piT * (1. + (piR / piY) * exp1) + (piC / piY) * exp21;
  return result[i][j];
}

void orc_calculate_bg(float dist, const float freqs[4], float kappa, const int *blosum, float scores[4],
                      float probsOut[4]) {
  const int *tc = orc_transcode();
  float probs[4][4];
  float counts[4] = {0, 0, 0, 0};
  float f, prob, score, probStop;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) probs[i][j] = prob_hky(i, j, dist, freqs, kappa);
  scores[0] = scores[1] = scores[2] = scores[3] = 0.0f;
  probStop = 0;
  for (int a1 = 0; a1 < 4; a1++)
    for (int a2 = 0; a2 < 4; a2++)
      for (int a3 = 0; a3 < 4; a3++)
        for (int b1 = 0; b1 < 4; b1++)
          for (int b2 = 0; b2 < 4; b2++)
            for (int b3 = 0; b3 < 4; b3++) {
              int pepA = tc[a1 * 16 + a2 * 4 + a3], pepB = tc[b1 * 16 + b2 * 4 + b3];
              if (pepA != -1 && pepB != -1) continue;
              f = (freqs[a1]) * (freqs[a2]) * (freqs[a3]);
              prob = probs[a1][b1] * probs[a2][b2] * probs[a3][b3];
              prob *= f;
              probStop += prob;
            }
  for (int a1 = 0; a1 < 4; a1++)
    for (int a2 = 0; a2 < 4; a2++)
      for (int a3 = 0; a3 < 4; a3++) {
        int pepA = tc[a1 * 16 + a2 * 4 + a3];
        if (pepA == -1) continue;
        for (int b1 = 0; b1 < 4; b1++)
          for (int b2 = 0; b2 < 4; b2++)
            for (int b3 = 0; b3 < 4; b3++) {
              int pepB = tc[b1 * 16 + b2 * 4 + b3];
              if (pepB == -1) continue;
              int h = (a1 != b1) + (a2 != b2) + (a3 != b3);
              f = (freqs[a1]) * (freqs[a2]) * (freqs[a3]);
              prob = probs[a1][b1] * probs[a2][b2] * probs[a3][b3];
              prob *= f;
              prob /= (1 - probStop);
              score = blosum[pepA * 24 + pepB];
              counts[h] += prob;
              scores[h] += score * prob;
            }
      }
  for (int i = 0; i < 4; i++) {
    scores[i] /= counts[i];
    if (probsOut) probsOut[i] = counts[i];
  }
}

/* ---- seq-gen restatement --------------------------------------------------------------------------------- */

void orc_mt_init(orc_mt *g, unsigned long s) { /* seqgen/twister.c:73-86 */
  g->mt[0] = s & 0xffffffffUL;
  for (g->mti = 1; g->mti < 624; g->mti++) {
    g->mt[g->mti] = (1812433253UL * (g->mt[g->mti - 1] ^ (g->mt[g->mti - 1] >> 30)) + g->mti);
    g->mt[g->mti] &= 0xffffffffUL;
  }
}

unsigned long orc_mt_next(orc_mt *g) { /* seqgen/twister.c:118-146 */
  static const unsigned long mag01[2] = {0x0UL, 0x9908b0dfUL};
  unsigned long y;
  if (g->mti >= 624) {
    int kk;
    for (kk = 0; kk < 624 - 397; kk++) {
      y = (g->mt[kk] & 0x80000000UL) | (g->mt[kk + 1] & 0x7fffffffUL);
      g->mt[kk] = g->mt[kk + 397] ^ (y >> 1) ^ mag01[y & 0x1UL];
    }
    for (; kk < 623; kk++) {
      y = (g->mt[kk] & 0x80000000UL) | (g->mt[kk + 1] & 0x7fffffffUL);
      g->mt[kk] = g->mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 0x1UL];
    }
    y = (g->mt[623] & 0x80000000UL) | (g->mt[0] & 0x7fffffffUL);
    g->mt[623] = g->mt[396] ^ (y >> 1) ^ mag01[y & 0x1UL];
    g->mti = 0;
  }
  y = g->mt[g->mti++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680UL;
  y ^= (y << 15) & 0xefc60000UL;
  y ^= (y >> 18);
  return y;
}

/* SetState, seqgen/evolve.c:167-175.  The reference can return 4 when r exceeds the last cumulative value (the
 * rows end a few 1e-10 below 1); that state indexes past its tables (undefined behaviour), here it is clamped to 3. */
static int set_state(orc_mt *g, const double *P) {
  double r = orc_mt_next(g) * (1.0 / 4294967295.0); /* genrand_real1, seqgen/twister.c:162-166 */
  int j;
  for (j = 0; j < 4 && r > P[j]; j++);
  return j > 3 ? 3 : j;
}

void orc_evolve(unsigned long seed, int n_nodes, const int *parent, const int *row, const double *cum,
                const double addFreq[4], int N, int cols, char *out_rows) {
  static const char nucleotides[] = "ACGT"; /* seqgen/nucmodels.c:47 */
  orc_mt g;
  unsigned char *seqs = (unsigned char *)malloc((size_t)n_nodes * cols);
  (void)N;
  orc_mt_init(&g, seed); /* SetSeed(randomSeed), src/treeSimulate.c:84-85 */
  for (int nd = 0; nd < n_nodes; nd++) {
    unsigned char *me = seqs + (size_t)nd * cols;
    if (parent[nd] < 0) {
      for (int i = 0; i < cols; i++) me[i] = (unsigned char)set_state(&g, addFreq); /* RandomSequence */
    } else {
      const unsigned char *anc = seqs + (size_t)parent[nd] * cols;
      const double *M = cum + (size_t)nd * 16;
      for (int i = 0; i < cols; i++) me[i] = (unsigned char)set_state(&g, M + anc[i] * 4); /* MutateSequence, NoRates */
    }
    if (row[nd] >= 0) /* tree2aln + sortAln */
      for (int i = 0; i < cols; i++) out_rows[(size_t)row[nd] * cols + i] = nucleotides[me[i]];
  }
  free(seqs);
}

double orc_cells(int N, int L) {
  double P = 0;
  for (int b = 1; b <= L; b++) P += (L - b + 1) / 3;
  return 2.0 * (N - 1) * P;
}
