/*
 * kmc_oracle.c — CPU restatement of jaxonwang/k-mer-count's hot path.  TEST INFRASTRUCTURE ONLY:
 * see kmc_oracle.h for who may load it and for the parity status of each mode.
 *
 * Citations are file:line under /root/reference (the reference is NOT read at run time).
 */
#define _GNU_SOURCE
#include "kmc_oracle.h"

#include <pthread.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ encoding ------------- */
/* A=0 C=1 G=2 T=3, first base most significant: numeric order == the bytewise String order
 * of main.rs:87 for equal-length ACGT strings.                                              */
static inline int orc_code_strict(uint8_t c) {
  switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return -1; /* main.rs:23 */
  }
}
static inline int orc_code_fold(uint8_t c) { return orc_code_strict((uint8_t)(c & 0xDF)); }

static inline uint64_t fmix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}
uint64_t orc_mix(uint64_t hi, uint64_t lo, uint64_t count) {
  uint64_t m = fmix64(lo ^ fmix64(hi ^ 0x9E3779B97F4A7C15ULL));
  return fmix64(m + count * 0xD6E8FEB86659FD93ULL);
}
uint64_t orc_digest(const orc_table *t) {
  uint64_t s = 0;
  for (uint64_t i = 0; i < t->n_distinct; i++) s += orc_mix(t->key_hi[i], t->key_lo[i], t->count[i]);
  return s;
}

void orc_table_free(orc_table *t) {
  if (!t) return;
  free(t->key_hi); free(t->key_lo); free(t->count);
  memset(t, 0, sizeof *t);
}
void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------ FASTA ---------------- */
/* bio 0.41 fasta::Reader::read (third-party, not in /root/reference; Cargo.lock:36-37), as used at
 * main.rs:45,59-62: a record starts at a line beginning '>'; header = rest of that line, trailing
 * whitespace trimmed, split at the first whitespace into id / desc; following lines up to the next
 * '>' line or EOF are appended with trailing whitespace trimmed; a first line that does not start
 * with '>' is an error; the loop at main.rs:60 stops at the first record for which
 * id=="" && desc==None && seq=="" (which is how EOF is signalled).                            */
static int is_space(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

int orc_parse_fasta(const char *path, uint8_t **bases_out, uint64_t **rec_off_out, uint64_t *n_recs_out) {
  FILE *f = fopen(path, "rb");
  if (!f) return ORC_E_IO;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  uint8_t *buf = (uint8_t *)malloc(sz > 0 ? sz : 1);
  if (sz > 0 && fread(buf, 1, sz, f) != (size_t)sz) { fclose(f); free(buf); return ORC_E_IO; }
  fclose(f);
  uint8_t *bases = (uint8_t *)malloc(sz > 0 ? sz : 1);
  uint64_t cap = 1024, n = 0, nb = 0;
  uint64_t *off = (uint64_t *)malloc((cap + 1) * sizeof(uint64_t));
  off[0] = 0;
  long p = 0;
  int rc = ORC_OK;
  while (p < sz) {
    /* header line */
    long e = p;
    while (e < sz && buf[e] != '\n') e++;
    if (buf[p] != '>') { rc = ORC_E_FORMAT; break; }
    long he = e;
    while (he > p + 1 && is_space(buf[he - 1])) he--;
    int header_empty = (he == p + 1); /* id=="" and desc==None */
    p = (e < sz) ? e + 1 : sz;
    uint64_t start = nb;
    while (p < sz && buf[p] != '>') {
      long le = p;
      while (le < sz && buf[le] != '\n') le++;
      long te = le;
      while (te > p && is_space(buf[te - 1])) te--;
      memcpy(bases + nb, buf + p, te - p);
      nb += te - p;
      p = (le < sz) ? le + 1 : sz;
    }
    if (header_empty && nb == start) break; /* record.is_empty() → main.rs:60-62 break */
    if (n == cap) { cap *= 2; off = (uint64_t *)realloc(off, (cap + 1) * sizeof(uint64_t)); }
    off[++n] = nb;
  }
  free(buf);
  if (rc != ORC_OK) { free(bases); free(off); return rc; }
  *bases_out = bases; *rec_off_out = off; *n_recs_out = n;
  return ORC_OK;
}

/* ------------------------------------------------------------------ lr-gapped, literal ---- */
static int cmp54(const void *a, const void *b) { return memcmp(a, b, 54); }

static void pack_str(const char *s, uint32_t nb, uint64_t *hi, uint64_t *lo) {
  unsigned __int128 v = 0;
  for (uint32_t i = 0; i < nb; i++) v = (v << 2) | (unsigned)orc_code_strict((uint8_t)s[i]);
  *lo = (uint64_t)v; *hi = (uint64_t)(v >> 64);
}

int orc_compat_lr(const uint8_t *bases, const uint64_t *rec_off, uint64_t n_recs, orc_table *out,
                  char **text, uint64_t *text_len) {
  const uint64_t l_len = 27, r_len = 27;              /* main.rs:48-49 */
  uint64_t n = 0;
  for (uint64_t r = 0; r < n_recs; r++) {
    uint64_t len = rec_off[r + 1] - rec_off[r];
    for (uint64_t d = 80; d < 141; d++) if (len >= d) n += len - d + 1;
  }
  char *lr = (char *)malloc((n ? n : 1) * 54);
  if (!lr) return ORC_E_NOMEM;
  uint64_t m = 0;
  for (uint64_t r = 0; r < n_recs; r++) {              /* main.rs:58-62 */
    const uint8_t *seq = bases + rec_off[r];
    uint64_t len = rec_off[r + 1] - rec_off[r];
    for (uint64_t dna_chunk_size = 80; dna_chunk_size < 141; dna_chunk_size++) { /* main.rs:63 */
      uint64_t window_start = 0;
      for (;;) {
        uint64_t m_len = dna_chunk_size - l_len - r_len;  /* main.rs:66 */
        uint64_t l_start = window_start;                  /* :67 */
        uint64_t l_end = l_start + l_len;                 /* :68 */
        uint64_t r_start = l_end + m_len;                 /* :69 */
        uint64_t r_end = r_start + r_len;                 /* :70 */
        window_start += 1;                                /* :71 */
        if (r_end > len) break;                           /* :73-75 */
        memcpy(lr + m * 54, seq + l_start, 27);           /* :76,78 */
        memcpy(lr + m * 54 + 27, seq + r_start, 27);      /* :77,78 */
        m++;                                              /* :79 */
      }
    }
  }
  memset(out, 0, sizeof *out);
  if (m == 0) { free(lr); return ORC_E_EMPTY; }          /* main.rs:35 source[0] */
  /* radix_sort's only observable effect: panic on a non-ACGT char at offsets 53..1 (main.rs:36,23) */
  int bad0 = 0;
  for (uint64_t i = 0; i < m; i++) {
    for (int o = 1; o < 54; o++)
      if (orc_code_strict((uint8_t)lr[i * 54 + o]) < 0) { free(lr); return ORC_E_BADBASE; }
    if (orc_code_strict((uint8_t)lr[i * 54]) < 0) bad0 = 1;
  }
  if (bad0) { free(lr); return ORC_E_BADBASE_OFFSET0; }
  qsort(lr, m, 54, cmp54);                               /* main.rs:87 */
  if (text) {                                            /* main.rs:88-90 */
    char *t = (char *)malloc(m * 55);
    for (uint64_t i = 0; i < m; i++) { memcpy(t + i * 55, lr + i * 54, 54); t[i * 55 + 54] = '\n'; }
    *text = t; *text_len = m * 55;
  }
  uint64_t nd = 0;
  for (uint64_t i = 0; i < m; i++) if (i == 0 || memcmp(lr + i * 54, lr + (i - 1) * 54, 54)) nd++;
  out->n_total = m; out->n_distinct = nd;
  out->key_hi = (uint64_t *)malloc(nd * 8); out->key_lo = (uint64_t *)malloc(nd * 8);
  out->count = (uint64_t *)malloc(nd * 8);
  uint64_t o = 0;
  for (uint64_t i = 0; i < m;) {
    uint64_t e = i + 1;
    while (e < m && !memcmp(lr + e * 54, lr + i * 54, 54)) e++;
    pack_str(lr + i * 54, 54, &out->key_hi[o], &out->key_lo[o]);
    out->count[o] = e - i;
    o++; i = e;
  }
  free(lr);
  return ORC_OK;
}

/* ------------------------------------------------------------------ contiguous, definitional */
/* PARITY UNPINNED (kmc_oracle.h).  Rules, from SURVEY.md §8c "proposed rules":
 *   - windows are k consecutive bytes of ONE record (never span records, as main.rs:58-81
 *     handles one record at a time);
 *   - a window is skipped iff any of its k bytes is not in {A,C,G,T,a,c,g,t};
 *   - lower case is folded to upper;
 *   - canonical: key = lexicographic min(window, reverse complement of window);
 *   - table ascending by key string; counts = multiplicity.                               */
typedef struct { unsigned __int128 v; } u128box;
static int cmp_u128(const void *a, const void *b) {
  unsigned __int128 x = ((const u128box *)a)->v, y = ((const u128box *)b)->v;
  return x < y ? -1 : x > y;
}
int orc_contiguous_def(const uint8_t *bases, const uint64_t *rec_off, uint64_t n_recs, uint32_t k,
                       int canonical, orc_table *out) {
  if (k < 1 || k > 64) return ORC_E_ARG;
  uint64_t cap = 0;
  for (uint64_t r = 0; r < n_recs; r++) {
    uint64_t len = rec_off[r + 1] - rec_off[r];
    if (len >= k) cap += len - k + 1;
  }
  u128box *keys = (u128box *)malloc((cap ? cap : 1) * sizeof(u128box));
  uint64_t n = 0;
  char fw[65], rv[65];
  for (uint64_t r = 0; r < n_recs; r++) {
    const uint8_t *s = bases + rec_off[r];
    uint64_t len = rec_off[r + 1] - rec_off[r];
    for (uint64_t w = 0; w + k <= len; w++) {
      int ok = 1;
      for (uint32_t i = 0; i < k; i++) {
        uint8_t c = (uint8_t)(s[w + i] & 0xDF);
        if (orc_code_strict(c) < 0) { ok = 0; break; }
        fw[i] = (char)c;
      }
      if (!ok) continue;
      for (uint32_t i = 0; i < k; i++) {
        char c = fw[k - 1 - i];
        rv[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
      }
      const char *pick = (canonical && memcmp(rv, fw, k) < 0) ? rv : fw;
      uint64_t hi, lo;
      pack_str(pick, k, &hi, &lo);
      keys[n++].v = (((unsigned __int128)hi) << 64) | lo;
    }
  }
  qsort(keys, n, sizeof(u128box), cmp_u128);
  uint64_t nd = 0;
  for (uint64_t i = 0; i < n; i++) if (i == 0 || keys[i].v != keys[i - 1].v) nd++;
  memset(out, 0, sizeof *out);
  out->n_total = n; out->n_distinct = nd;
  out->key_hi = (uint64_t *)malloc((nd ? nd : 1) * 8); out->key_lo = (uint64_t *)malloc((nd ? nd : 1) * 8);
  out->count = (uint64_t *)malloc((nd ? nd : 1) * 8);
  uint64_t o = 0;
  for (uint64_t i = 0; i < n;) {
    uint64_t e = i + 1;
    while (e < n && keys[e].v == keys[i].v) e++;
    out->key_hi[o] = (uint64_t)(keys[i].v >> 64); out->key_lo[o] = (uint64_t)keys[i].v; out->count[o] = e - i;
    o++; i = e;
  }
  free(keys);
  return ORC_OK;
}

/* ------------------------------------------------------------------ packed, multi-threaded - */
#define KEY_T uint64_t
#define SUF 64
#include "orc_mt_impl.inc"
#undef KEY_T
#undef SUF
#define KEY_T unsigned __int128
#define SUF 128
#include "orc_mt_impl.inc"
#undef KEY_T
#undef SUF

int orc_contiguous_mt(const uint8_t *bases, const uint64_t *rec_off, uint64_t n_recs, uint32_t k,
                      int canonical, int n_threads, orc_table *out) {
  if (k < 1 || k > 64) return ORC_E_ARG;
  if (k <= 32) return orc_run_mt_64(bases, rec_off, n_recs, 0, k, canonical, 0, 0, 0, 0, 2 * k, n_threads, out);
  return orc_run_mt_128(bases, rec_off, n_recs, 0, k, canonical, 0, 0, 0, 0, 2 * k, n_threads, out);
}

int orc_gapped_mt(const uint8_t *bases, const uint64_t *rec_off, uint64_t n_recs, uint32_t l_len,
                  uint32_t r_len, uint32_t d_min, uint32_t d_max, int n_threads, orc_table *out) {
  if (l_len < 1 || r_len < 1 || l_len > 32 || r_len > 32 || d_min < l_len + r_len || d_max < d_min) return ORC_E_ARG;
  uint32_t kb = 2 * (l_len + r_len);
  if (kb <= 64) return orc_run_mt_64(bases, rec_off, n_recs, 1, 0, 0, l_len, r_len, d_min, d_max, kb, n_threads, out);
  return orc_run_mt_128(bases, rec_off, n_recs, 1, 0, 0, l_len, r_len, d_min, d_max, kb, n_threads, out);
}

/* ------------------------------------------------------------------ emit ------------------ */
static void unpack_key(uint64_t hi, uint64_t lo, uint32_t nb, char *s) {
  unsigned __int128 v = (((unsigned __int128)hi) << 64) | lo;
  for (uint32_t i = 0; i < nb; i++) { s[nb - 1 - i] = "ACGT"[(unsigned)(v & 3)]; v >>= 2; }
}
int64_t orc_emit_expanded(const orc_table *t, uint32_t nb, const char *path) {
  FILE *f = path ? fopen(path, "wb") : stdout;
  if (!f) return ORC_E_IO;
  char line[130];
  int64_t bytes = 0;
  for (uint64_t i = 0; i < t->n_distinct; i++) {
    unpack_key(t->key_hi[i], t->key_lo[i], nb, line);
    line[nb] = '\n';
    for (uint64_t c = 0; c < t->count[i]; c++) { fwrite(line, 1, nb + 1, f); bytes += nb + 1; }
  }
  if (path) fclose(f); else fflush(f);
  return bytes;
}
int64_t orc_emit_counts(const orc_table *t, uint32_t nb, const char *path) {
  FILE *f = path ? fopen(path, "wb") : stdout;
  if (!f) return ORC_E_IO;
  char line[160];
  int64_t bytes = 0;
  for (uint64_t i = 0; i < t->n_distinct; i++) {
    unpack_key(t->key_hi[i], t->key_lo[i], nb, line);
    int m = snprintf(line + nb, sizeof line - nb, "\t%llu\n", (unsigned long long)t->count[i]);
    fwrite(line, 1, nb + m, f); bytes += nb + m;
  }
  if (path) fclose(f); else fflush(f);
  return bytes;
}
