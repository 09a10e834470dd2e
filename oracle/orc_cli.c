/*
 * orc_cli.c — command-line face of the CPU oracle.  TEST INFRASTRUCTURE ONLY (see kmc_oracle.h).
 *   orc_cli compat FASTA            → the reference's stdout (main.rs:88-90)
 *   orc_cli counts FASTA K CANON    → "kmer\tcount\n" (contiguous mode, parity unpinned)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "kmc_oracle.h"
int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s compat FASTA | counts FASTA K CANON [THREADS]\n", argv[0]); return 2; }
  uint8_t *bases; uint64_t *off, nrec;
  int rc = orc_parse_fasta(argv[2], &bases, &off, &nrec);
  if (rc) { fprintf(stderr, "parse error %d\n", rc); return 101; }
  orc_table t;
  if (!strcmp(argv[1], "compat")) {
    char *text; uint64_t len;
    rc = orc_compat_lr(bases, off, nrec, &t, &text, &len);
    if (rc) { fprintf(stderr, "oracle error %d\n", rc); return 101; }
    fwrite(text, 1, len, stdout);
    free(text);
  } else {
    if (argc < 5) return 2;
    int thr = argc > 5 ? atoi(argv[5]) : 1;
    rc = orc_contiguous_mt(bases, off, nrec, (uint32_t)atoi(argv[3]), atoi(argv[4]), thr, &t);
    if (rc) { fprintf(stderr, "oracle error %d\n", rc); return 101; }
    orc_emit_counts(&t, (uint32_t)atoi(argv[3]), NULL);
  }
  orc_table_free(&t);
  return 0;
}
