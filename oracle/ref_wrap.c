/* oracle/ref_wrap.c -- TEST INFRASTRUCTURE. Link-time wrappers (GNU ld --wrap) that make the
 * UNMODIFIED reference deterministic without touching its sources.
 *
 *  - CreateSeed (reference: seqgen/twister.c:195-224, called once per null sample from
 *    src/treeSimulate.c:84) normally hashes time()/clock().  The wrapper returns
 *        rc_det_seed(base, block, sample)
 *    where base = $RNACODE_SEED (default 1), block = number of treeML calls so far - 1,
 *    sample = number of CreateSeed calls since the last treeML call.  Deriving the seed from
 *    (block, sample) instead of a running counter keeps seeds stable under --stop-early.
 *  - treeML (reference: src/treeML.c:35, called once per scored block from src/RNAcode.c:153)
 *    is wrapped only to advance the block counter.
 *
 *  The same rc_det_seed() is restated in rnacode_oracle.c and in the host code of the product
 *  so that all three draw identical null alignments.
 */
#include <stdlib.h>

struct aln;
int __real_treeML(const struct aln *alignment[], char **treeString, float *kappa);

static long g_block = -1;
static long g_sample = 0;

unsigned long rc_det_seed(unsigned long base, unsigned long block, unsigned long sample) {
  /* 32-bit mix (MT19937's init_genrand only keeps the low 32 bits) */
  unsigned long x = (base * 2654435761UL) ^ (block * 40503UL + 0x9E3779B9UL) ^ (sample * 2246822519UL);
  x ^= x >> 15; x *= 2246822519UL; x &= 0xffffffffUL;
  x ^= x >> 13; x *= 3266489917UL; x &= 0xffffffffUL;
  x ^= x >> 16;
  return x & 0xffffffffUL;
}

static unsigned long base_seed(void) {
  const char *e = getenv("RNACODE_SEED");
  return e ? strtoul(e, NULL, 10) : 1UL;
}

/* the probe may set these explicitly */
void rc_wrap_set_block(long b) { g_block = b; g_sample = 0; }
void rc_wrap_set_sample(long s) { g_sample = s; }
long rc_wrap_block(void) { return g_block; }
long rc_wrap_sample(void) { return g_sample; }

unsigned long __wrap_CreateSeed(void) {
  unsigned long s = rc_det_seed(base_seed(), (unsigned long)(g_block < 0 ? 0 : g_block), (unsigned long)g_sample);
  g_sample++;
  return s;
}

int __wrap_treeML(const struct aln *alignment[], char **treeString, float *kappa) {
  g_block++;
  g_sample = 0;
  return __real_treeML(alignment, treeString, kappa);
}
