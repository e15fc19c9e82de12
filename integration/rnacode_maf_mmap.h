/* integration/rnacode_maf_mmap.h -- (f3) MAF reader over a memory-mapped file for the batched driver.
 *
 * Same blocks, fields and error messages as the reference's read_maf (src/rnaz_utils.c:132-234) driven by
 * get_line / splitFields (src/utils.c:172-189, src/rnaz_utils.c:480-563), without their per-line and per-field
 * malloc / realloc / strdup churn: lines and fields are located in place in the mapping, and the only copies made are
 * the name and the sequence of each 's' line into the struct aln the rest of the host code works on (the sequence is
 * upper-cased on the way, as main() does right after reading, src/RNAcode.c:127-133).
 *
 * Conventions kept: a block is the run of 's' lines up to the next 'a' line or the end of the file; empty lines, '#'
 * comments and 'i' / 'e' / 'q' lines are skipped, as is any other line; fields are maximal runs of non-space characters;
 * an 's' line must have exactly 7 fields; integers are read with sscanf("%d"); the strand must be '+' or '-'; all
 * sequences of a block must have the same length.  checkFormat() (src/rnaz_utils.c:430-475) has already consumed the
 * first 'a' line: mapping starts at the stream's current offset.
 */
#ifndef RNACODE_MAF_MMAP_H
#define RNACODE_MAF_MMAP_H

#include <ctype.h>
#include <sys/mman.h>
#include <sys/stat.h>

typedef struct {
  const char *base, *p, *end;
  size_t len;
} rc_maf_map;

/* 1 if `f` is a regular file and could be mapped; the cursor is placed at the stream's current position */
static int rc_maf_map_open(FILE *f, rc_maf_map *m) {
  struct stat st;
  long pos = ftell(f);
  void *a;
  if (pos < 0 || fstat(fileno(f), &st) != 0 || !S_ISREG(st.st_mode) || st.st_size <= 0) return 0;
  a = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fileno(f), 0);
  if (a == MAP_FAILED) return 0;
  madvise(a, (size_t)st.st_size, MADV_SEQUENTIAL);
  m->base = (const char *)a;
  m->len = (size_t)st.st_size;
  m->p = m->base + ((size_t)pos < m->len ? (size_t)pos : m->len);
  m->end = m->base + m->len;
  return 1;
}

static void rc_maf_map_close(rc_maf_map *m) {
  if (m->base) munmap((void *)m->base, m->len);
  m->base = m->p = m->end = NULL;
}

static int rc_maf_int(const char *s, int n, const char *what, const char *fmt_msg) {
  char buf[64];
  int v;
  if (n > 63) n = 63;
  memcpy(buf, s, n);
  buf[n] = '\0';
  if (sscanf(buf, "%d", &v) != 1) {
    fprintf(stderr, fmt_msg, buf);
    (void)what;
    exit(EXIT_FAILURE);
  }
  return v;
}

/* read_maf() on the mapping: fills alignedSeqs (NULL-terminated), returns the number of rows, 0 at the end of the file */
static int rc_read_maf_mapped(rc_maf_map *m, struct aln *alignedSeqs[]) {
  int num_seq = 0, nn;
  size_t n;
again:
  if (m->p >= m->end) return 0;
  while (m->p < m->end) {
    const char *line = m->p, *eol = (const char *)memchr(line, '\n', (size_t)(m->end - line));
    const char *stop = eol ? eol : m->end, *q = line;
    const char *fs[8];
    int fl[8], nf = 0;
    m->p = eol ? eol + 1 : m->end;
    while (q < stop) { /* fields = maximal runs of non-space characters (splitFields) */
      while (q < stop && isspace((unsigned char)*q)) q++;
      if (q >= stop) break;
      {
        const char *b = q;
        while (q < stop && !isspace((unsigned char)*q)) q++;
        if (nf < 8) {
          fs[nf] = b;
          fl[nf] = (int)(q - b);
        }
        nf++;
      }
    }
    if (nf == 0) continue;         /* empty line */
    if (fs[0][0] == '#') continue; /* comment */
    if (fl[0] == 1 && (fs[0][0] == 'i' || fs[0][0] == 'e' || fs[0][0] == 'q')) continue;
    if (fl[0] == 1 && fs[0][0] == 's') {
      char *name, *seq, strand;
      int start, length, fullLength, k;
      if (nf != 7) nrerror("ERROR: Invalid MAF format (number of fields in 's' line not correct)");
      if (num_seq >= MAX_NUM_NAMES - 1) nrerror("ERROR: Too many sequences in MAF block");
      start = rc_maf_int(fs[2], fl[2], "start", "ERROR: Invalid MAF format (start position '%s' is not an integer)\n");
      length = rc_maf_int(fs[3], fl[3], "length", "ERROR: Invalid MAF format (length '%s' is not an integer)\n");
      fullLength = rc_maf_int(fs[5], fl[5], "srcSize", "ERROR: Invalid MAF format (source sequence length '%s' is not an integer)\n");
      strand = fs[4][0];
      if (strand != '+' && strand != '-') {
        char buf[64];
        int c = fl[4] > 63 ? 63 : fl[4];
        memcpy(buf, fs[4], c);
        buf[c] = '\0';
        fprintf(stderr, "ERROR: Invalid MAF format (strand field '%s' is not '+' or '-')\n", buf);
        exit(EXIT_FAILURE);
      }
      name = (char *)malloc((size_t)fl[1] + 1);
      memcpy(name, fs[1], (size_t)fl[1]);
      name[fl[1]] = '\0';
      seq = (char *)malloc((size_t)fl[6] + 1);
      for (k = 0; k < fl[6]; k++) seq[k] = (char)toupper((unsigned char)fs[6][k]);
      seq[fl[6]] = '\0';
      alignedSeqs[num_seq++] = createAlnEntry(name, seq, start, length, fullLength, strand);
      continue;
    }
    if (fl[0] == 1 && fs[0][0] == 'a') break; /* next block */
  }
  alignedSeqs[num_seq] = NULL;
  if (num_seq == 0) {
    /* an 'a' line without any 's' line (the reference's read_maf hands an empty block to main(), which then crashes on it):
     * not the end of the input -- go on with the next block instead of silently dropping the rest of the file */
    if (m->p < m->end) goto again;
    return 0; /* nothing but blank / comment lines were left */
  }
  n = strlen(alignedSeqs[0]->seq);
  for (nn = 1; nn < num_seq; nn++)
    if (strlen(alignedSeqs[nn]->seq) != n) nrerror("ERROR: Sequences are of unequal length.");
  return num_seq;
}

#endif
