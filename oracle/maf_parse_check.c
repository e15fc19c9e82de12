/* oracle/maf_parse_check.c -- TEST INFRASTRUCTURE.  Parses a MAF file twice, with the reference's read_maf
 * (src/rnaz_utils.c:132-234) and with the memory-mapped reader of the batched driver (integration/rnacode_maf_mmap.h),
 * and compares every block field by field.  Prints "OK <blocks> <rows>" or the first difference. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "RNAcode.h"
#include "rnaz_utils.h"
#include "utils.h"

#include "../integration/rnacode_maf_mmap.h"

int main(int argc, char *argv[]) {
  struct aln *a[MAX_NUM_NAMES], *b[MAX_NUM_NAMES];
  FILE *fa, *fb;
  rc_maf_map map;
  long blocks = 0, rows = 0;
  int na, nb, i, j;
  if (argc < 2) return 2;
  if (argc > 2) { /* --reference-only / --mapped-only FILE: one parser alone (error paths: both exit with a message) */
    const int mapped_only = strcmp(argv[1], "--mapped-only") == 0;
    fa = fopen(argv[2], "r");
    if (!fa || checkFormat(fa) != MAF) return 2;
    if (mapped_only && !rc_maf_map_open(fa, &map)) return 2;
    while ((na = mapped_only ? rc_read_maf_mapped(&map, a) : read_maf(fa, a)) != 0) {
      rows += na;
      blocks++;
    }
    printf("OK %ld %ld\n", blocks, rows);
    return 0;
  }
  fa = fopen(argv[1], "r");
  fb = fopen(argv[1], "r");
  if (!fa || !fb) return 2;
  if (checkFormat(fa) != MAF || checkFormat(fb) != MAF) {
    printf("not MAF\n");
    return 2;
  }
  if (!rc_maf_map_open(fb, &map)) {
    printf("cannot map\n");
    return 2;
  }
  for (;;) {
    na = read_maf(fa, a);
    nb = rc_read_maf_mapped(&map, b);
    if (na != nb) {
      printf("block %ld: %d rows vs %d rows\n", blocks, na, nb);
      return 1;
    }
    if (na == 0) break;
    for (i = 0; i < na; i++) {
      for (j = 0; a[i]->seq[j]; j++) a[i]->seq[j] = toupper(a[i]->seq[j]); /* src/RNAcode.c:127-133 */
      if (strcmp(a[i]->name, b[i]->name) || strcmp(a[i]->seq, b[i]->seq) || a[i]->start != b[i]->start ||
          a[i]->length != b[i]->length || a[i]->fullLength != b[i]->fullLength || a[i]->strand != b[i]->strand) {
        printf("block %ld row %d differs\n", blocks, i);
        return 1;
      }
    }
    rows += na;
    blocks++;
    freeAln(a);
    freeAln(b);
  }
  rc_maf_map_close(&map);
  printf("OK %ld %ld\n", blocks, rows);
  return 0;
}
