/*
 * kmc_oracle.h — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libkmc.so, the kmer-count CLI) never links or calls it.
 *
 * Parity status
 *   lr-gapped (compat) mode : PINNED  — follows k-mer-count/src/main.rs:48-90 and is checked against
 *                             the unchanged test.py run on k-mer-count/sample.fasta
 *                             (sha256 00f3e1ea…, tests/golden/compat_golden.json).
 *   contiguous mode         : PARITY UNPINNED — the reference has no contiguous k-mer mode, no k,
 *                             no canonical form and no N handling.  The rules
 *                             are this build's own definition, stated at orc_contiguous_def().
 */
#ifndef KMC_ORACLE_H
#define KMC_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
  ORC_OK = 0,
  ORC_E_IO = -1,          /* main.rs:44  File::open(..).expect(..)                      */
  ORC_E_FORMAT = -2,      /* main.rs:59  reader.read(..).unwrap(): "Expected > at record start." */
  ORC_E_BADBASE = -3,     /* main.rs:23  panic!("Unexpected charactor ..")              */
  ORC_E_EMPTY = -4,       /* main.rs:35  source[0] on an empty Vec                      */
  ORC_E_NOMEM = -5,
  ORC_E_ARG = -6,
  ORC_E_BADBASE_OFFSET0 = -7 /* non-ACGT byte only at chunk offset 0: main.rs:36 never looks
                                there and prints it verbatim; a 2-bit key cannot hold it, so
                                this build refuses (documented divergence)        */
};

/* Sorted (ascending by (key_hi,key_lo)) table of distinct keys with multiplicities. */
typedef struct {
  uint64_t n_distinct;
  uint64_t n_total;
  uint64_t *key_hi; /* all zero when the key fits 64 bits */
  uint64_t *key_lo;
  uint64_t *count;
} orc_table;

void orc_table_free(orc_table *t);

/* FASTA → concatenated sequence bytes + record offsets (n_recs+1 entries).
 * Restates bio 0.41 fasta::Reader::read as called at main.rs:45,59-62.            */
int orc_parse_fasta(const char *path, uint8_t **bases, uint64_t **rec_off, uint64_t *n_recs);
void orc_free(void *p);

/* lr-gapped mode, literal: strings, memcmp sort.  main.rs:48-90.  If `text` is non-NULL it
 * receives the exact stdout bytes of the reference (malloc'd, *text_len bytes).       */
int orc_compat_lr(const uint8_t *bases, const uint64_t *rec_off, uint64_t n_recs,
                  orc_table *out, char **text, uint64_t *text_len);

/* contiguous mode, definitional (string level, single thread, small inputs). */
int orc_contiguous_def(const uint8_t *bases, const uint64_t *rec_off, uint64_t n_recs,
                       uint32_t k, int canonical, orc_table *out);

/* contiguous mode, packed + multi-threaded (same results; used for large parity cases and as
 * the timed CPU baseline: extract → partition by key prefix → radix sort → group equals,
 * the algorithm class of main.rs:87).                                                  */
int orc_contiguous_mt(const uint8_t *bases, const uint64_t *rec_off, uint64_t n_recs,
                      uint32_t k, int canonical, int n_threads, orc_table *out);

/* generalised gapped mode: L/R lengths and chunk-size range as parameters;
 * l=r=27, dmin=80, dmax=140 is the reference.  Packed, multi-threaded.                  */
int orc_gapped_mt(const uint8_t *bases, const uint64_t *rec_off, uint64_t n_recs,
                  uint32_t l_len, uint32_t r_len, uint32_t d_min, uint32_t d_max,
                  int n_threads, orc_table *out);

/* Order-independent digest: sum over rows of mix(key_hi,key_lo,count) mod 2^64. */
uint64_t orc_digest(const orc_table *t);
uint64_t orc_mix(uint64_t hi, uint64_t lo, uint64_t count);

/* Text emitters. compat: each key repeated `count` times, 54 chars + '\n' (main.rs:88-90).
 * counts: "kmer\tcount\n".  Return bytes written or <0.                                */
int64_t orc_emit_expanded(const orc_table *t, uint32_t key_bases, const char *path);
int64_t orc_emit_counts(const orc_table *t, uint32_t key_bases, const char *path);

#ifdef __cplusplus
}
#endif
#endif
