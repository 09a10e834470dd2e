/*
 * kmc.h — C ABI of libkmc.so, the B200 (sm_100a) k-mer counting engine.
 *
 * The reference (jaxonwang/k-mer-count) has no plugin / operator / FFI interface: its only boundary
 * is the process (k-mer-count/src/main.rs:43-44 in, :88-90 out).  This header is the thin C ABI the
 * task's north_star asks for between host code and the CUDA kernels, following the proposal in
 * SURVEY.md §8b.  Every entry point cites the reference lines whose work it replaces.  Plain C
 * types only — no torch / C++ types cross this boundary — so a Rust `-sys` crate, cgo or ctypes can
 * bind it unchanged (INTEGRATION.md shows the bindings).
 *
 * Call chain (one ctx per GPU, one host thread per ctx; a ctx is not thread-safe):
 *     kmc_create → { kmc_staging → fill → kmc_submit }*  → kmc_finish → kmc_read* → kmc_destroy
 * or, when keys + table exceed HBM, in key-range passes over the resident input:
 *     kmc_submit* → { kmc_finish_part(p, P) → kmc_read* }  for p = 0..P-1
 * or, with inputs already resident in HBM:
 *     kmc_create → kmc_submit_device* → kmc_finish → kmc_table_device / kmc_read
 * Multi-GPU (one process per GPU):
 *     kmc_submit* → kmc_route_to_peers(n_parts)   [keys stored into the owners' buffers over NVLink]
 *     (pipelined: kmc_dist_hist → kmc_owner_begin → { kmc_route_to_peers_part(c) → hand-over → kmc_owner_feed* }* → kmc_finish)
 *                 → [counts exchanged, ranks synchronised] → kmc_ingest_keys* → kmc_finish
 *   or, with the exchange done by the host (NCCL all-to-all):
 *     kmc_submit* → kmc_route(n_parts) → [all-to-all of the routed keys] → kmc_ingest_keys* → kmc_finish
 *
 * There is no CPU fallback anywhere behind this ABI: without a CUDA device kmc_create fails with
 * KMC_E_NO_DEVICE.
 */
#ifndef KMC_H
#define KMC_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KMC_ABI_VERSION 1

/* ---- error codes (all entry points return 0 or one of these; nothing throws or aborts) ------- */
enum {
  KMC_OK = 0,
  KMC_E_ARG = -1,            /* bad argument / bad call order                                      */
  KMC_E_NO_DEVICE = -2,      /* no CUDA device / wrong architecture                                */
  KMC_E_CUDA = -3,           /* a CUDA runtime call failed; kmc_last_error has the text            */
  KMC_E_NOMEM = -4,          /* host or device allocation failed                                   */
  KMC_E_BADBASE = -5,        /* lr-gapped mode: a byte other than A,C,G,T inside an emitted chunk:
                                main.rs:23 panic!("Unexpected charactor ..") → exit 101            */
  KMC_E_EMPTY = -6,          /* lr-gapped mode: no chunk at all: main.rs:35 source[0] panics       */
  KMC_E_COUNT_OVERFLOW = -7, /* a multiplicity does not fit the table's 32-bit count               */
  KMC_E_CAPACITY = -8,       /* more input than the ctx was sized for and it could not grow        */
  KMC_E_BADBASE_OFFSET0 = -9,/* lr-gapped mode: non-ACGT byte only at chunk offset 0, which
                                main.rs:36 never inspects and would print verbatim; a 2-bit key
                                cannot hold it, so this build refuses (documented divergence)      */
  KMC_E_FORMAT = -10         /* FASTA text does not start with '>': main.rs:59 `reader.read(..).unwrap()`
                                ("Expected > at record start.")                                      */
};

/* ---- configuration ---------------------------------------------------------------------------- */
enum { KMC_MODE_CONTIGUOUS = 0, /* ordinary k-mers, k <= 32 → 64-bit keys, k <= 64 → 128-bit keys   */
       KMC_MODE_LR_GAPPED = 1 };/* the reference: L‖R gapped pairs, main.rs:48-49,63-80            */
enum { KMC_STRATEGY_AUTO = 0, KMC_STRATEGY_HASH = 1, KMC_STRATEGY_SORT = 2,
       KMC_STRATEGY_SORT_BASELINE = 3 /* plain extract → LSD radix sort → run-length encode        */ };

typedef struct {
  uint32_t abi_version;    /* KMC_ABI_VERSION                                                     */
  uint32_t mode;           /* KMC_MODE_*                                                          */
  uint32_t k;              /* contiguous: 1..64; ignored in lr-gapped mode                        */
  uint32_t canonical;      /* contiguous: key = min(k-mer, reverse complement); must be 0 for lr  */
  uint32_t strategy;       /* KMC_STRATEGY_*                                                      */
  int32_t  device;         /* CUDA ordinal, -1 = current device                                   */
  uint32_t l_len, r_len;   /* lr-gapped: 0,0 → 27,27  (main.rs:48-49); each 1..32                 */
  uint32_t d_min, d_max;   /* lr-gapped: chunk sizes, 0,0 → 80,140 (main.rs:63 `80..141`)         */
  uint64_t expected_bases; /* sizing hint for device buffers; 0 = grow on demand                  */
  uint32_t reserved[8];    /* must be zero                                                        */
} kmc_config;

typedef struct kmc_ctx kmc_ctx; /* opaque; owns all device memory, streams and events             */

/* Replaces main.rs:43-50 (process set-up: open input, allocate `lr_chunk`).                        */
int kmc_create(kmc_ctx **out, const kmc_config *cfg);
void kmc_destroy(kmc_ctx *ctx);
/* Text of the last failure on this ctx (or of the last failed kmc_create when ctx is NULL);
 * library-owned, valid until the next call.  Replaces the panic message on stderr.                */
const char *kmc_last_error(const kmc_ctx *ctx);
const char *kmc_strerror(int code);

/* Optional: run everything on the caller's CUDA stream (a cudaStream_t / CUstream as void*).      */
int kmc_set_stream(kmc_ctx *ctx, void *cuda_stream);
/* Forget submitted input and results but keep device buffers (next job on the same ctx).          */
int kmc_reset(kmc_ctx *ctx);

/* ---- input: what main.rs:58-62 + :73 (`reader.read`, `record.seq()`) hand to the hot loop ------
 * Sequence bytes are raw ASCII, records concatenated; rec_off has n_recs+1 entries, rec_off[0]=0,
 * rec_off[n_recs]=n_bases (offsets relative to THIS submit).  Windows never span records.  The N /
 * lower-case policy lives in the kernels, not in the parser.                                      */

/* Library-owned PINNED host buffers for the host to fill (double-buffered: the pointers change
 * between calls; a call may block until the previous copy out of the returned buffer finished).   */
int kmc_staging(kmc_ctx *ctx, size_t want_bases, size_t want_recs, uint8_t **bases, uint64_t **rec_off,
                size_t *cap_bases, size_t *cap_recs);
/* Asynchronous H2D of the last staging buffer + append to the device-resident input.              */
int kmc_submit(kmc_ctx *ctx, size_t n_bases, size_t n_recs);
/* Same, from caller-owned HOST memory (pageable or pinned); copies synchronously if pageable.     */
int kmc_submit_host(kmc_ctx *ctx, const uint8_t *bases, const uint64_t *rec_off, size_t n_bases, size_t n_recs);
/* Raw FASTA text (the bytes of the file, HOST memory): parsed ON THE DEVICE into sequence bytes + record
 * offsets with the rules of bio's fasta::Reader as main.rs:45,59-62 uses it (header lines start with '>',
 * sequence lines are appended with trailing whitespace trimmed, the first all-empty record ends the input),
 * then submitted.  Replaces the host-side line parser in front of kmc_submit.  n_bases/n_recs (optional)
 * receive what was found.                                                                           */
int kmc_submit_fasta(kmc_ctx *ctx, const uint8_t *fasta_text, size_t n_bytes, uint64_t *n_bases, uint64_t *n_recs);
/* Same, but the input is already in HBM (device pointers, `bases` 16-byte aligned).  The buffers
 * are referenced, not copied, and must stay valid and unchanged until kmc_finish returns.         */
int kmc_submit_device(kmc_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_rec_off, size_t n_bases, size_t n_recs);

/* ---- the hot path: main.rs:63-81 (window extraction) + :84-87 (ordering / grouping) ------------
 * Extract every key, count, and leave the table — distinct keys ascending by (key_hi,key_lo), with
 * multiplicities — resident in HBM.  n_total = number of key occurrences (lines the reference would
 * print), n_distinct = rows.                                                                       */
int kmc_finish(kmc_ctx *ctx, uint64_t *n_distinct, uint64_t *n_total);
/* The same for inputs whose keys or table do not fit in HBM at once (BASELINE.json configs 3-4 on ONE GPU:
 * 1e10 k-mers are 80-160 GB of keys): count in n_parts passes over the resident input.  The key space is cut
 * into n_parts consecutive key ranges of about equal population (from a histogram of the submitted input, so the
 * same input always gives the same ranges); this call counts range `part` only and leaves ITS table — n_total =
 * the key occurrences inside the range.  Ranges ascend with `part`: reading the tables of part 0, 1, ...,
 * n_parts-1 in turn yields exactly the rows kmc_finish would have produced, in the same order (main.rs:87-90),
 * and the kmc_digest values add up (mod 2^64).  The input stays submitted between calls; the previous part's
 * table is dropped.  Errors that concern the whole input (main.rs:23,35) are raised by every part.           */
int kmc_finish_part(kmc_ctx *ctx, uint32_t part, uint32_t n_parts, uint64_t *n_distinct, uint64_t *n_total);

/* ---- output: what main.rs:88-90 prints, as arrays --------------------------------------------
 * Rows [first, first+n) into caller-owned HOST arrays (any of them may be NULL).  key_hi is zero for
 * keys of <= 64 bits.  Key encoding: A=0,C=1,G=2,T=3, first base most significant, so ascending
 * integer order is the bytewise String order of main.rs:87.                                        */
int kmc_read(kmc_ctx *ctx, uint64_t first, uint64_t n, uint64_t *key_lo, uint64_t *key_hi, uint64_t *count);
/* Rows [first, first+n) as text, formatted ON THE DEVICE into a library-owned pinned host buffer (*text, *len
 * bytes; valid until the next call on this ctx).  expanded != 0: each key as letters + '\n', repeated `count`
 * times — exactly what main.rs:88-90 prints; expanded == 0: "<kmer>\t<count>\n".  Returns KMC_E_CAPACITY when
 * the text of these rows exceeds max_bytes (ask for fewer rows).                                    */
int kmc_format(kmc_ctx *ctx, uint64_t first, uint64_t n, int expanded, size_t max_bytes, const char **text, size_t *len);
/* Device pointers of the table (valid until reset/destroy).  d_key_hi is NULL for 64-bit keys.    */
int kmc_table_device(kmc_ctx *ctx, const uint64_t **d_key_lo, const uint64_t **d_key_hi, const uint32_t **d_count);
/* Order-independent 64-bit digest of the table, computed on the device:
 * sum over rows of mix(key_hi,key_lo,count) mod 2^64 (SURVEY.md §8d full-scale parity check).     */
int kmc_digest(kmc_ctx *ctx, uint64_t *digest);
/* Number of bases per key: k, or l_len+r_len in lr-gapped mode.                                   */
uint32_t kmc_key_bases(const kmc_ctx *ctx);

/* ---- multi-GPU: hash-prefix routing -------------------------------------------
 * kmc_route: extract this rank's keys from the submitted input and group them by owner part
 * (kmc_owner_of, a hash prefix) in one device buffer; part p's keys are the part_count[p] keys starting
 * at key index part_begin[p] of *d_keys (both caller-owned host arrays of n_parts entries; the parts need
 * not be adjacent).  Keys are 8 bytes (k<=32) or 16 bytes (lo,hi) each; key_bytes tells which.      */
int kmc_route(kmc_ctx *ctx, uint32_t n_parts, uint64_t *part_begin, uint64_t *part_count, const void **d_keys,
              uint32_t *key_bytes);
/* Fused route + exchange over NVLink peer memory: like kmc_route, but part p's keys are stored by the
 * routing kernel straight into d_part_ptr[p] — normally this rank's region of rank p's receive buffer,
 * mapped with kmc_ipc_open (or any device pointer, e.g. torch symmetric memory).  Each region holds at
 * most part_cap_keys keys; part_count[p] (host) receives how many keys part p has.  A count above
 * part_cap_keys means that region overflowed (nothing of the job is usable): the ranks must agree on a
 * larger capacity — the counts are exact — and every rank calls kmc_route_to_peers again.  The caller exchanges
 * the counts and synchronises the ranks before the owners call kmc_ingest_keys on what they received.   */
int kmc_route_to_peers(kmc_ctx *ctx, uint32_t n_parts, void *const *d_part_ptr, uint64_t part_cap_keys,
                       uint64_t *part_count);
/* The same in chunks, so that the exchange overlaps the owners' counting (streaming owner, below): chunk `chunk` of
 * `n_chunks` equal slices of the input; part_count receives the CUMULATIVE counts after this chunk (region p holds
 * keys [0, part_count[p]) of this rank; the previous call's counts say where the new ones start).  max_ctas > 0:
 * the routing kernel takes at most that many SMs and leaves the rest to the owner's kernels running beside it.     */
int kmc_route_to_peers_part(kmc_ctx *ctx, uint32_t n_parts, void *const *d_part_ptr, uint64_t part_cap_keys,
                            uint64_t *part_count, uint32_t chunk, uint32_t n_chunks, uint32_t max_ctas);
/* Streaming owner: count the keys the ranks route here while they are still routing.
 *   kmc_owner_begin(global_hist, n_owners)  plan the partitioned count for this owner's share — 1 / n_owners of every
 *       bin of global_hist, the sum of all ranks' kmc_dist_hist histograms (the owner function is a hash).
 *       *streaming = 0: declined (small job, strategy, keys that do not suit the path): route as usual, then
 *       kmc_ingest_keys + kmc_finish.
 *   kmc_owner_feed(d_keys, n)  after every chunk's hand-over: the keys that have just arrived in one sender's region.
 *       Asynchronous (a second stream, beside the next chunk's routing kernel); the memory is referenced until
 *       kmc_finish returns.
 *   kmc_finish  the rest of the count.  Falls back to a recount of everything fed if a bucket overflowed.          */
int kmc_owner_begin(kmc_ctx *ctx, const uint64_t global_hist[4096], uint32_t n_owners, uint32_t *streaming);
int kmc_owner_feed(kmc_ctx *ctx, const void *d_keys, uint64_t n_keys);
/* ---- multi-GPU, range partition: the level-1 scatter of the counting pipeline done by the SENDERS --------
 * For high-cardinality input the hash route above costs an extra pass: owners re-scatter what they received.
 * Here the ranks agree on one plan for the whole key space; every sender runs the ordinary level-1 scatter into a
 * local staging array laid out owner by owner, and each owner's slab crosses NVLink as ONE device-to-device copy
 * into the owner's receive buffer; owners run only the second scatter and the per-bucket sort.  Owners hold
 * consecutive key ranges of about equal population, so the ranks' tables in rank order are the globally sorted
 * table.  The input goes in n_chunks equal chunks: chunk c + 1 is scattered while chunk c is on the links and
 * the owners work on chunk c - 1.  Per job, on every rank:
 *   kmc_submit* → kmc_dist_hist → [all-gather the histograms] → kmc_dist_plan_chunks → kmc_recv_buffer(need) (+ IPC
 *   exchange when a buffer moved) → for every chunk c: kmc_dist_scatter_part(c) [asynchronous], then
 *   kmc_dist_scatter_wait(c) → [rank barrier: everybody's chunk c has landed] → kmc_dist_owner_part(c);
 *   → kmc_dist_scatter_end → [all-reduce of `overflow`] → kmc_finish
 * kmc_dist_hist: upper-estimate histogram of this rank's keys over the top 12 key bits (4096 bins; sampled), and
 *   whether its input looks low-cardinality (then the hash route + hash table is the better path).
 * kmc_dist_plan_chunks: all_hist = the `world` histograms in rank order (identical on every rank, so every rank
 *   derives the same plan).  need_bytes[r] = receive-buffer size rank r must provide (kmc_recv_buffer); all zero
 *   when the job does not suit the plan (tiny, or keys sharing long prefixes) — use the hash route then.
 *   kmc_dist_plan = one chunk.
 * kmc_dist_scatter_part: d_peer_buf[r] = rank r's receive buffer as mapped in this process; chunks in order.
 * kmc_dist_scatter_wait: returns when this rank's copies of that chunk have landed in the owners' buffers.
 * kmc_dist_owner_part: this rank's level-2 scatter over a chunk every sender has delivered (optional: what has not
 *   been handed over chunk by chunk is done by kmc_finish).
 * kmc_dist_scatter_end: *overflow != 0: a bucket exceeded its planned capacity; all ranks must then recount through
 *   the hash route (input is still resident).
 * kmc_dist_scatter: all chunks, then kmc_dist_scatter_end (no overlap with the owners; one-call form).
 * kmc_finish afterwards counts what the senders stored in this rank's receive buffer.                         */
int kmc_dist_hist(kmc_ctx *ctx, uint64_t hist[4096], uint32_t *low_cardinality);
int kmc_dist_plan(kmc_ctx *ctx, uint32_t world, uint32_t rank, const uint64_t *all_hist, uint64_t *need_bytes);
int kmc_dist_plan_chunks(kmc_ctx *ctx, uint32_t world, uint32_t rank, const uint64_t *all_hist, uint32_t n_chunks,
                         uint64_t *need_bytes);
int kmc_dist_scatter_part(kmc_ctx *ctx, void *const *d_peer_buf, uint32_t chunk);
int kmc_dist_scatter_wait(kmc_ctx *ctx, uint32_t chunk);
int kmc_dist_owner_part(kmc_ctx *ctx, uint32_t chunk);
int kmc_dist_scatter_end(kmc_ctx *ctx, uint32_t *overflow);
int kmc_dist_scatter(kmc_ctx *ctx, void *const *d_peer_buf, uint32_t *overflow);
/* Library-owned device buffer for received keys (grow-only; 16 bytes per key in 128-bit mode).      */
int kmc_recv_buffer(kmc_ctx *ctx, uint64_t n_keys, void **d_ptr);
/* CUDA IPC plumbing for one-process-per-GPU: export a device allocation of this process as a 64-byte
 * handle; map a peer's handle into this process; unmap it.                                         */
int kmc_ipc_export(kmc_ctx *ctx, const void *d_ptr, unsigned char handle[64]);
int kmc_ipc_open(kmc_ctx *ctx, const unsigned char handle[64], void **d_peer_ptr);
int kmc_ipc_close(kmc_ctx *ctx, void *d_peer_ptr);
/* Hand the ctx keys it owns (device pointer, same layout as kmc_route's output; referenced until
 * kmc_finish).  kmc_finish then counts the ingested keys instead of extracting from the input.    */
int kmc_ingest_keys(kmc_ctx *ctx, const void *d_keys, uint64_t n_keys);
/* ---- multi-GPU, low-cardinality input: count locally, exchange rows ("(key,count) pairs after local
 * combine when cardinality is low") -------------------------------------------------------------------------------
 * When every rank's input has few distinct keys (BASELINE.json config 5: reads from a small repetitive genome), routing
 * every key occurrence to its owner moves gigabytes to merge what fits in megabytes.  Instead each rank counts its own
 * shard (kmc_finish), groups the rows of its table by owner with kmc_table_route, the rows are exchanged (one small
 * all-to-all of keys and one of counts), and every owner merges what it received:
 *   kmc_submit* → kmc_finish → kmc_table_route(n_parts) → [all-to-all] → kmc_reset → kmc_ingest_pairs* → kmc_finish
 * kmc_table_route: part p's rows are the part_count[p] entries from index part_begin[p] of *d_keys (keys) and
 *   *d_counts (counts, 64-bit): library-owned device arrays, valid until the next kmc_table_route or kmc_destroy.
 * kmc_ingest_pairs: rows this ctx owns (device pointers, referenced until kmc_finish).  kmc_finish then leaves the
 *   sorted table in which equal keys' counts have been added up; n_total = the sum of all counts.  Cannot be mixed
 *   with submitted input or kmc_ingest_keys.  Both calls: keys of <= 64 bits only.                                */
int kmc_table_route(kmc_ctx *ctx, uint32_t n_parts, uint64_t *part_begin, uint64_t *part_count, const uint64_t **d_keys,
                    const uint64_t **d_counts);
int kmc_ingest_pairs(kmc_ctx *ctx, const uint64_t *d_keys, const uint64_t *d_counts, uint64_t n_rows);
/* ---- multi-GPU output stage: one ascending table from the owners' tables (main.rs:87-90 on N ranks) -----
 * After a hash-partitioned count every rank holds an ascending table of keys no other rank holds; what the
 * reference prints is their merge.  kmc_merge_tables makes that merge this ctx's table (as after kmc_finish:
 * kmc_read / kmc_format / kmc_digest / kmc_table_device work on it): n_runs ascending runs of rows in DEVICE
 * memory, columns laid out like kmc_table_device's (d_key_hi may be NULL for keys of <= 64 bits), n_rows[r]
 * rows each; the runs are not referenced after the call returns.  The runs' key sets must be pairwise disjoint
 * (KMC_E_ARG otherwise).  The ctx must hold no input (fresh, or after kmc_reset), and no run may be this ctx's
 * own table (merge into a second ctx).                                                                    */
int kmc_merge_tables(kmc_ctx *ctx, uint32_t n_runs, const uint64_t *const *d_key_lo, const uint64_t *const *d_key_hi,
                     const uint32_t *const *d_count, const uint64_t *n_rows, uint64_t *n_distinct, uint64_t *n_total);
/* owner part of a key, host-side (the same function the device uses).                            */
uint32_t kmc_owner_of(uint64_t key_hi, uint64_t key_lo, uint32_t n_parts);

/* ---- synthetic input on the device (random_fasta_generator.py:5-15) ----------
 * The reference's generator prints 200 unseeded 400-base records and takes no arguments.  These are its
 * seeded, scalable stand-in: every byte is a pure function of (seed, stream, index) through Philox4x32-10
 * (csrc/kmc_gen.cuh), so any window of a stream can be produced in any order, and the host twin
 * (k-mer-count_b200/gen.py, tools/gen_fasta.py) produces the same bytes.  All pointers are device memory.
 * kmc_gen_bases: upper-case ACGT, bases [first, first + n) of stream `seed` → d_out[0..n).
 * kmc_gen_nruns: lays that seed's N runs (1e-4 starts per base, geometric length of mean 50) over
 *   d_bases[0..n) = bases [first, first + n).
 * kmc_gen_reads: reads [first_read, first_read + n_reads) of read_len bases, each from a uniform position
 *   and strand of d_genome[0..genome_len) → d_out[0 .. n_reads * read_len).                           */
int kmc_gen_bases(kmc_ctx *ctx, uint64_t seed, uint64_t first, uint64_t n, uint8_t *d_out);
int kmc_gen_nruns(kmc_ctx *ctx, uint64_t seed, uint64_t first, uint64_t n, uint8_t *d_bases);
int kmc_gen_reads(kmc_ctx *ctx, uint64_t seed, const uint8_t *d_genome, uint64_t genome_len, uint32_t read_len,
                  uint64_t first_read, uint64_t n_reads, uint8_t *d_out);

/* ---- introspection ------------------------------------------------------------------------------
 * JSON object: per-phase device times (CUDA events on the ctx stream), kernel-launch count, chosen
 * strategy, sizes.  Returns bytes needed (incl. NUL); writes at most cap.                          */
size_t kmc_stats_json(kmc_ctx *ctx, char *buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* KMC_H */
