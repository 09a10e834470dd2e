// kmc_api.cu — the C ABI of include/kmc.h: context, staging, pipeline orchestration.
// All device work is hand-written CUDA for sm_100a (kmc_extract.cuh, kmc_sort.cuh, kmc_fast.cuh,
// kmc_hash.cuh).  There is no CPU fallback: every path below launches kernels or fails.
#include "../../include/kmc.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: phases show up as ranges in Nsight Systems / ncu --nvtx

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kmc_common.cuh"
#include "kmc_extract.cuh"
#include "kmc_sort.cuh"
#include "kmc_fast.cuh"
#include "kmc_hash.cuh"
#include "kmc_hash128.cuh"
#include "kmc_fasta.cuh"
#include "kmc_format.cuh"
#include "kmc_gen.cuh"
#include "kmc_merge.cuh"
#include <cmath>

using namespace kmc;

namespace {

thread_local std::string g_create_err;

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};

struct Segment {            // one piece of input, resident in HBM (a submit, or a chunk of a large pinned submit)
  DevBuf own_bases, own_off, brk;
  const uint8_t *bases = nullptr;   // points into own_bases or at caller memory (submit_device)
  const uint64_t *rec_off = nullptr;
  uint64_t n_bases = 0, n_recs = 0;
  uint64_t off_shift = 0;           // added to rec_off[] values to make them relative to this segment (chunks share one array)
  const uint8_t *host_alias = nullptr; // the same bases in pinned host memory (readable by kernels): sampling passes use
                                       // it so that they need not wait for the copy
  cudaEvent_t ready = nullptr;      // recorded on the copy stream after this segment's H2D
  bool wait_ready = false;
};

struct Phase {
  std::string name;
  cudaEvent_t a = nullptr, b = nullptr;
  float ms = 0.f;
  double host_begin = 0, host_end = 0; // host clock (ms) when the phase was opened / closed: finds host-side stalls
};
inline double host_now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
struct KernelStat {
  std::string name;
  uint64_t launches = 0;
  double ms = 0;
};

} // namespace

// kmc_dist_plan's result (see there)
struct DistOwner {            // the owner's half of a range-partitioned count in progress
  FastPlan pl{};              // pseudo-bucket x = (chunk, bucket, sender)
  bool key32 = false, split64 = false;
  uint32_t nb_max = 1, n_xc = 0;  // n_xc = pseudo-buckets per chunk
  uint64_t t_max = 1, n_fine = 0;
  unsigned int *ticket = nullptr;
  unsigned long long *d_total = nullptr, *status = nullptr;
};
struct DistPlan {
  bool valid = false, scattered = false, owner_ready = false;
  uint32_t world = 0, rank = 0, b1 = 0, n_all = 0; // n_all = 2^b1 level-1 buckets over the whole key space
  uint32_t n_chunks = 1, chunks_sent = 0, chunks_owned = 0;
  uint32_t fine_cap = 0;                           // capacity of a fine bucket's element buffer in fast_finish for this plan
  std::vector<uint32_t> own_lo;                    // [world+1] first level-1 bucket of every owner
  // sender view: capacity of MY region of bucket b (per chunk) and its offset inside my slab at b's owner; per owner:
  // where my slab starts inside a chunk of its array, the slab's length, the length of one chunk of its array, and
  // where the slab sits in my staging array
  std::vector<double> cum_frac;                    // [n_chunks+1] chunk c = the input's CTA tiles [cum_frac[c], cum_frac[c+1]) (first and last chunk are half-size: short fill and drain)
  std::vector<uint64_t> s_cap, s_in;               // [n_chunks x n_all]
  std::vector<uint64_t> slab_pre, slab_len, chunk_off, stage_off; // [n_chunks x world]; chunk_off = start of chunk c in that owner's array
  uint64_t stage_len = 0;                          // keys of one staging half (the largest chunk's slabs for all other owners)
  std::vector<uint8_t> l1e;                        // [n_all]
  std::vector<uint64_t> fine_hist;                 // [ncoarse] global upper estimate (fine-bucket capacities)
  std::vector<uint64_t> x_cap, x_off;              // [n_chunks x my level-1 buckets x world] owner view: capacity / offset in my array of every sender's region
  uint64_t l1_keys = 0;                            // keys my level-1 array must hold (all chunks)
  DistOwner owner;
  cudaEvent_t ev_ready = nullptr, ev_scattered[16] = {}, ev_copied[16] = {};
};

struct kmc_ctx {
  kmc_config cfg{};
  int device = 0;
  uint32_t n_sms = 148;   // multiProcessorCount of the device (kmc_create); grids are sized in multiples of it
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  unsigned char *mailbox = nullptr, *mailbox_dev = nullptr; // pinned + mapped: small results written by kernels
  unsigned char *upbox = nullptr, *upbox_dev = nullptr;     // pinned + mapped: small uploads pulled by a kernel
  size_t upbox_cap = 0;
  cudaStream_t copy_stream = nullptr; // chunked H2D of large pinned submits, overlapped with the level-1 scatter
  uint32_t key_bits = 0, key_bases = 0;
  bool wide = false; // 128-bit keys

  // pinned staging, double-buffered
  uint8_t *h_bases[2] = {nullptr, nullptr};
  uint64_t *h_off[2] = {nullptr, nullptr};
  size_t cap_bases = 0, cap_recs = 0;
  int cur = 0;
  bool staged = false;
  cudaEvent_t copy_done[2] = {nullptr, nullptr};
  bool copy_pending[2] = {false, false};

  // input
  std::vector<Segment> segs;
  size_t n_segs = 0; // live segments (segs beyond are pooled buffers)
  uint64_t total_bases = 0, total_recs = 0;

  // ingested keys (multi-GPU)
  std::vector<std::pair<const void *, uint64_t>> ingested;
  // ingested (key, count) rows (multi-GPU, locally combined counts): keys, counts, rows
  struct PairArray { const uint64_t *keys; const uint64_t *counts; uint64_t n; };
  std::vector<PairArray> ingested_pairs;
  DevBuf pair_rows, pair_state; // kmc_table_route: rows grouped by owner; per-part population / cursors

  // work buffers (grow-only, reused across kmc_reset)
  DevBuf keys_a, keys_b, block_hist, offsets, sums, scalars, route_keys;
  DevBuf gap_l, gap_r, gap_f;
  DevBuf t_lo, t_hi, t_cnt;
  DevBuf fast_l1, fast_l2, fast_state, fast_tables, fast_fdesc, recv_keys;
  DevBuf route_state, route_tables;     // routing pass (sender role): part cursors and table
  DevBuf hash_slots, hash_scalars, hash_hot;
  DevBuf fa_raw, fa_tiles, fa_flags;
  DevBuf fmt_len, fmt_off, fmt_text;
  DevBuf merge_lo, merge_hi, merge_cnt; // kmc_merge_tables: the other half of the ping-pong
  char *fmt_host = nullptr;
  size_t fmt_host_cap = 0;
  uint64_t probe_distinct = 0;
  uint32_t n_hot = 0;
  uint32_t hash_aborts = 0;
  std::vector<unsigned char> fast_host; // plan tables staged for upload
  uint32_t fast_fallbacks = 0;          // times the partitioned path overflowed and the job was recounted
  const char *fast_variant = "";        // level-2 element form of the last partitioned count: u32 / split64 / u64 / u128

  // partial count (kmc_finish_part): the coarse-bin range [range_lo, range_lo + range_n) being counted, and the
  // coarse histogram of the whole input it is cut from (computed once per input)
  bool range_on = false;
  uint32_t range_lo = 0, range_n = 0;
  std::vector<uint64_t> part_hist; // raw (sampled) counts per coarse bin
  // ... and, for contiguous input, ALL keys scattered once by their top bits (kmc_finish_part, first call): every part is
  // then counted from its slice of this array instead of extracting the whole input again
  bool kept_valid = false, kept_tried = false;
  uint32_t kept_b1 = 0;
  std::vector<uint64_t> kept_start, kept_count; // per top-bits bucket: key index in kept_keys, keys
  DevBuf kept_keys, kept_tables, kept_state;
  uint32_t part_hist_step = 0;     // sampling step of part_hist; 0 = not computed yet

  // multi-GPU range partition (kmc_dist_*): plan shared by all ranks, this rank's views of it
  DistPlan dist;
  void *job_box = nullptr;                // the partitioned count in progress (a FastJob, defined with the fast path)
  // streaming owner (kmc_owner_*): keys fed to the partitioned count while the other ranks are still routing
  cudaStream_t owner_stream = nullptr;
  bool owner_on = false;
  std::vector<std::pair<const void *, uint64_t>> owner_fed; // what was fed, for a recount should the count not suit the path
  DevBuf dist_tables, dist_stage, dist_cursors;
  cudaStream_t peer_stream = nullptr;     // range partition: the slab copies to the owners (beside the next chunk's scatter)
  cudaStream_t peer_lane[8] = {};         // ... further streams for copies to different peers at the same time
  cudaEvent_t peer_lane_ev[8] = {};

  // results
  bool finished = false;
  uint64_t n_total = 0, n_distinct = 0;
  uint32_t strategy_used = 0;

  // stats
  std::vector<Phase> phases;
  std::vector<cudaEvent_t> event_pool;
  size_t events_used = 0;
  uint64_t launches = 0, launches_total = 0;
  uint64_t h2d_bytes = 0;
  std::string stats;
  // optional per-kernel timing (env KMC_KERNEL_TIMING=1): one event pair per launch
  bool ktiming = false;
  std::vector<Phase> klaunches;
  std::vector<KernelStat> kstats;
  std::vector<std::pair<const char *, double>> host_marks; // KMC_HOST_PROF: host clock at points between phases
};

namespace {

int fail(kmc_ctx *c, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_err = buf;
  return code;
}

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return fail(c, e_ == cudaErrorMemoryAllocation ? KMC_E_NOMEM : KMC_E_CUDA, "%s: %s (%s:%d)", #call, \
                  cudaGetErrorString(e_), __FILE__, __LINE__);                                           \
  } while (0)

int ktime_begin(kmc_ctx *c, const char *name);
int ktime_end(kmc_ctx *c);

#define LAUNCH(kern, grid, block, smem, ...)                                                             \
  do {                                                                                                   \
    if (c->ktiming) { int r_ = ktime_begin(c, #kern); if (r_) return r_; }                               \
    kern<<<(grid), (block), (smem), c->stream>>>(__VA_ARGS__);                                           \
    c->launches++;                                                                                       \
    CK(cudaGetLastError());                                                                              \
    if (c->ktiming) { int r_ = ktime_end(c); if (r_) return r_; }                                        \
  } while (0)

int ensure(kmc_ctx *c, DevBuf &b, size_t bytes) {
  if (bytes <= b.cap && b.p) return KMC_OK;
  if (b.p) { CK(cudaStreamSynchronize(c->stream)); CK(cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
  size_t want = std::max<size_t>(bytes, 256);
  // sizes derived from sampled estimates wobble by a fraction of a percent from job to job: leave headroom so that
  // a slightly larger next job does not free and reallocate gigabytes (cudaFree/cudaMalloc stall the host for tens of ms)
  // (small buffers too: plan tables of a few MB grow by single descriptors from job to job)
  want += want / 8 + 4096;
  want = (want + 255) & ~size_t(255);
  CK(cudaMalloc(&b.p, want));
  b.cap = want;
  return KMC_OK;
}
void release(DevBuf &b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr; b.cap = 0;
}

int phase_begin(kmc_ctx *c, const char *name) {
  Phase ph;
  ph.name = name;
  for (cudaEvent_t *e : {&ph.a, &ph.b}) {
    if (c->events_used == c->event_pool.size()) {
      cudaEvent_t ev;
      CK(cudaEventCreate(&ev));
      c->event_pool.push_back(ev);
    }
    *e = c->event_pool[c->events_used++];
  }
  CK(cudaEventRecord(ph.a, c->stream));
  nvtxRangePushA(name);
  ph.host_begin = host_now_ms();
  c->phases.push_back(ph);
  return KMC_OK;
}
int phase_end(kmc_ctx *c) {
  CK(cudaEventRecord(c->phases.back().b, c->stream));
  nvtxRangePop();
  c->phases.back().host_end = host_now_ms();
  return KMC_OK;
}
int ktime_begin(kmc_ctx *c, const char *name) {
  Phase ph;
  ph.name = name;
  for (cudaEvent_t *e : {&ph.a, &ph.b}) {
    if (c->events_used == c->event_pool.size()) {
      cudaEvent_t ev;
      CK(cudaEventCreate(&ev));
      c->event_pool.push_back(ev);
    }
    *e = c->event_pool[c->events_used++];
  }
  CK(cudaEventRecord(ph.a, c->stream));
  c->klaunches.push_back(ph);
  return KMC_OK;
}
int ktime_end(kmc_ctx *c) {
  CK(cudaEventRecord(c->klaunches.back().b, c->stream));
  return KMC_OK;
}
#define HOST_MARK(name) c->host_marks.emplace_back(name, host_now_ms())
#define PHASE_BEGIN(name) do { int r_ = phase_begin(c, name); if (r_) return r_; } while (0)
#define PHASE_END() do { int r_ = phase_end(c); if (r_) return r_; } while (0)
#define TRY(...) do { int r_ = (__VA_ARGS__); if (r_) return r_; } while (0)

inline uint32_t grid_for(uint64_t n, uint32_t per_block) { return (uint32_t)std::max<uint64_t>(1, (n + per_block - 1) / per_block); }

// scalars buffer layout (device): [0] cursor u64, [1] digest u64, [2] err flags u32 (in a u64 slot), [3] spare
unsigned long long *d_cursor(kmc_ctx *c) { return (unsigned long long *)c->scalars.p; }
unsigned long long *d_digest(kmc_ctx *c) { return (unsigned long long *)c->scalars.p + 1; }
uint32_t *d_err(kmc_ctx *c) { return (uint32_t *)((unsigned long long *)c->scalars.p + 2); }
unsigned long long *d_total_all(kmc_ctx *c) { return (unsigned long long *)c->scalars.p + 3; }
uint32_t *d_route_err(kmc_ctx *c) { return (uint32_t *)((unsigned long long *)c->scalars.p + 4); } // flags of a routing pass

// Small device→host reads go through a pinned, device-mapped mailbox written by a kernel, not through the copy
// engines: a cudaMemcpy D2H queues behind the 64 MB H2D chunks of a large submit and would stall the compute
// stream until every chunk has landed (measured: +6 ms per job).  Synchronises the compute stream.
__global__ void mailbox_copy_kernel(const unsigned char *__restrict__ src, unsigned char *__restrict__ dst, uint32_t bytes) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < bytes; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
constexpr size_t kMailboxBytes = 64 << 10;
int d2h_small(kmc_ctx *c, void *dst, const void *d_src, size_t bytes, size_t mailbox_off = 0) {
  if (!c->mailbox) {
    CK(cudaHostAlloc((void **)&c->mailbox, kMailboxBytes, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void **)&c->mailbox_dev, c->mailbox, 0));
  }
  if (mailbox_off + bytes > kMailboxBytes) return fail(c, KMC_E_ARG, "d2h_small: %zu bytes do not fit the mailbox", bytes);
  LAUNCH(mailbox_copy_kernel, std::max<uint32_t>(1, (uint32_t)std::min<size_t>(bytes / 256, 32)), 256, 0, (const unsigned char *)d_src,
         c->mailbox_dev + mailbox_off, (uint32_t)bytes);
  c->launches--; // plumbing, not part of the hot path's launch count
  CK(cudaStreamSynchronize(c->stream));
  memcpy(dst, c->mailbox + mailbox_off, bytes);
  return KMC_OK;
}

// ... and small host→device uploads (plan tables) likewise: staged in pinned mapped memory and pulled by a kernel,
// because a cudaMemcpy H2D would queue behind the chunk copies on the same engine.
__global__ void upload_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, uint64_t n16) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
int h2d_small(kmc_ctx *c, void *d_dst, const void *src, size_t bytes) {
  size_t need = (bytes + 15) & ~size_t(15);
  if (need > c->upbox_cap) {
    CK(cudaStreamSynchronize(c->stream));
    if (c->upbox) CK(cudaFreeHost(c->upbox));
    c->upbox = nullptr; c->upbox_cap = 0;
    size_t cap = std::max<size_t>(need * 2, 1 << 20);
    CK(cudaHostAlloc((void **)&c->upbox, cap, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void **)&c->upbox_dev, c->upbox, 0));
    c->upbox_cap = cap;
  } else {
    CK(cudaStreamSynchronize(c->stream)); // the previous upload kernel may still be reading the box
  }
  memcpy(c->upbox, src, bytes);
  LAUNCH(upload_kernel, (uint32_t)std::min<size_t>(std::max<size_t>(1, need / 16 / 256), (size_t)c->n_sms * 4), 256, 0,
         (const uint4 *)c->upbox_dev, (uint4 *)d_dst, (uint64_t)(need / 16));
  c->launches--;
  return KMC_OK;
}

int zero_scalars(kmc_ctx *c) {
  TRY(ensure(c, c->scalars, 64));
  CK(cudaMemsetAsync(c->scalars.p, 0, 64, c->stream));
  return KMC_OK;
}
int read_scalars(kmc_ctx *c, uint64_t *cursor, uint32_t *err, uint64_t *total_all = nullptr) {
  unsigned long long h[4];
  TRY(d2h_small(c, h, c->scalars.p, sizeof h));
  if (cursor) *cursor = h[0];
  if (err) *err = (uint32_t)h[2];
  if (total_all) *total_all = h[3];
  return KMC_OK;
}

// coarse bins (kmc_fast.cuh): the top min(12, key_bits) bits of a key
uint32_t coarse_bits(const kmc_ctx *c) { return std::min<uint32_t>(kCoarseBitsMax, c->key_bits); }

// kernel parameters of one segment; host_alias: read the pinned host copy (sampling passes before the H2D has landed)
ExtractParams seg_params(const kmc_ctx *c, const Segment &s, bool host_alias = false) {
  ExtractParams P{host_alias ? s.host_alias : s.bases, (const uint32_t *)s.brk.p, s.n_bases, c->cfg.k, c->cfg.canonical, 0, 0, 0, 0};
  if (c->range_on) { P.range_on = 1; P.range_shift = c->key_bits - coarse_bits(c); P.range_lo = c->range_lo; P.range_n = c->range_n; }
  return P;
}
GapParams gap_params(const kmc_ctx *c, const Segment &s) {
  const kmc_config &f = c->cfg;
  GapParams P{s.bases, (const uint32_t *)s.brk.p, s.n_bases, f.l_len, f.r_len, f.d_min, f.d_max, 0, 0, 0, 0};
  if (c->range_on) { P.range_on = 1; P.range_shift = c->key_bits - coarse_bits(c); P.range_lo = c->range_lo; P.range_n = c->range_n; }
  return P;
}

// ---- input segments ------------------------------------------------------------------------------------
int new_segment(kmc_ctx *c, Segment **out) {
  if (c->n_segs == c->segs.size()) c->segs.emplace_back();
  Segment &s = c->segs[c->n_segs++];
  s.bases = nullptr; s.rec_off = nullptr; s.n_bases = s.n_recs = 0;
  s.off_shift = 0; s.host_alias = nullptr; s.wait_ready = false;
  c->part_hist_step = 0; // new input: the partial-count histogram is stale
  c->kept_valid = false; c->kept_tried = false;
  *out = &s;
  return KMC_OK;
}

// after bases / rec_off are in place (device): build the record-start mask
int segment_mark(kmc_ctx *c, Segment &s) {
  size_t words = (s.n_bases + 31) / 32 + 4;
  TRY(ensure(c, s.brk, words * 4));
  CK(cudaMemsetAsync(s.brk.p, 0, words * 4, c->stream));
  if (s.n_recs)
    LAUNCH(mark_breaks_kernel, grid_for(s.n_recs, 256), 256, 0, s.rec_off, s.n_recs, (unsigned long long)s.off_shift, (uint32_t *)s.brk.p);
  return KMC_OK;
}

// kernels that read a segment's bases from HBM must not start before its copy has landed
int seg_wait(kmc_ctx *c, Segment &s) {
  if (s.wait_ready) { CK(cudaStreamWaitEvent(c->stream, s.ready, 0)); s.wait_ready = false; }
  return KMC_OK;
}

// A large submit from PINNED host memory: cut it at record boundaries into ~64 MB chunks, one segment each, copied
// on a separate stream.  kmc_finish's sampling passes read the pinned memory directly (1/16 of it), and the level-1
// scatter of chunk i runs while chunk i+1 is still on the bus.  The caller's buffer must stay unchanged until
// kmc_finish returns (library staging buffers: until the next kmc_staging hands them out again).
constexpr size_t kChunkBases = (size_t)64 << 20;
int submit_chunked(kmc_ctx *c, const uint8_t *bases, const uint8_t *bases_dev_alias, const uint64_t *rec_off, size_t n_bases,
                   size_t n_recs, cudaEvent_t done) {
  if (!c->copy_stream) CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  c->segs.reserve(c->segs.size() + 2 * (n_bases / kChunkBases) + 4); // Segment pointers below must stay valid
  // everything queued so far on the compute stream (e.g. a reset's tail) precedes the copies
  Segment *first;
  TRY(new_segment(c, &first));
  TRY(ensure(c, first->own_off, (n_recs + 1) * 8));
  if (!first->ready) CK(cudaEventCreateWithFlags(&first->ready, cudaEventDisableTiming));
  CK(cudaEventRecord(first->ready, c->stream));
  CK(cudaStreamWaitEvent(c->copy_stream, first->ready, 0));
  {
    int r_ = phase_begin(c, "h2d"); // events of this phase live on the copy stream
    if (r_) return r_;
    CK(cudaEventRecord(c->phases.back().a, c->copy_stream));
  }
  const size_t h2d_phase = c->phases.size() - 1;
  CK(cudaMemcpyAsync(first->own_off.p, rec_off, (n_recs + 1) * 8, cudaMemcpyHostToDevice, c->copy_stream));
  const uint64_t *d_off = (const uint64_t *)first->own_off.p;
  size_t r0 = 0;
  Segment *s = first;
  bool is_first = true;
  while (r0 < n_recs || is_first) {
    // records [r0, r1): as many as fit the chunk (at least one)
    size_t lo = r0 + 1, hi = n_recs;
    const uint64_t limit = rec_off[r0] + kChunkBases;
    while (lo < hi) { size_t mid = (lo + hi + 1) / 2; if (rec_off[mid] <= limit) lo = mid; else hi = mid - 1; }
    size_t r1 = n_recs ? std::min(std::max(lo, r0 + 1), n_recs) : 0;
    if (!is_first) TRY(new_segment(c, &s));
    const uint64_t b0 = n_recs ? rec_off[r0] : 0, b1 = n_recs ? rec_off[r1] : 0;
    TRY(ensure(c, s->own_bases, (b1 - b0) + 64));
    if (!s->ready) CK(cudaEventCreateWithFlags(&s->ready, cudaEventDisableTiming));
    if (b1 > b0) CK(cudaMemcpyAsync(s->own_bases.p, bases + b0, b1 - b0, cudaMemcpyHostToDevice, c->copy_stream));
    CK(cudaEventRecord(s->ready, c->copy_stream));
    s->wait_ready = true;
    s->bases = (const uint8_t *)s->own_bases.p;
    s->host_alias = bases_dev_alias + b0;
    s->rec_off = d_off + r0;
    s->off_shift = (uint64_t)0 - b0;
    s->n_bases = b1 - b0; s->n_recs = r1 - r0;
    is_first = false;
    r0 = r1;
    if (!n_recs) break;
  }
  if (done) CK(cudaEventRecord(done, c->copy_stream));
  CK(cudaEventRecord(c->phases[h2d_phase].b, c->copy_stream));
  c->h2d_bytes += n_bases + (n_recs + 1) * 8;
  c->total_bases += n_bases; c->total_recs += n_recs;
  // record-start masks need only the offsets: wait for that copy (the first event on the copy stream after it)
  PHASE_BEGIN("mark");
  for (size_t i = 0; i < c->n_segs; i++) {
    Segment &g = c->segs[i];
    if (g.rec_off < d_off || g.rec_off > d_off + n_recs) continue; // a segment of an earlier submit
    // the offsets copy precedes this segment's bases copy on the copy stream, so its ready event covers both
    CK(cudaStreamWaitEvent(c->stream, first->ready, 0));
    TRY(segment_mark(c, g));
  }
  PHASE_END();
  return KMC_OK;
}

int submit_from_host(kmc_ctx *c, const uint8_t *bases, const uint64_t *rec_off, size_t n_bases, size_t n_recs,
                     cudaEvent_t done) {
  if (c->finished) return fail(c, KMC_E_ARG, "kmc_submit after kmc_finish (call kmc_reset first)");
  if (n_bases >= 4 * kChunkBases && n_recs >= 2 && !getenv("KMC_NO_CHUNKED_H2D")) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, bases) == cudaSuccess && at.type == cudaMemoryTypeHost) {
      cudaPointerAttributes at2;
      if (cudaPointerGetAttributes(&at2, rec_off) == cudaSuccess && at2.type == cudaMemoryTypeHost)
        return submit_chunked(c, bases, (const uint8_t *)at.devicePointer, rec_off, n_bases, n_recs, done);
    }
    cudaGetLastError();
  }
  Segment *s;
  TRY(new_segment(c, &s));
  TRY(ensure(c, s->own_bases, n_bases + 64));
  TRY(ensure(c, s->own_off, (n_recs + 1) * 8));
  PHASE_BEGIN("h2d");
  if (n_bases) CK(cudaMemcpyAsync(s->own_bases.p, bases, n_bases, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(s->own_off.p, rec_off, (n_recs + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  if (done) CK(cudaEventRecord(done, c->stream));
  PHASE_END();
  c->h2d_bytes += n_bases + (n_recs + 1) * 8;
  s->bases = (const uint8_t *)s->own_bases.p;
  s->rec_off = (const uint64_t *)s->own_off.p;
  s->n_bases = n_bases; s->n_recs = n_recs;
  c->total_bases += n_bases; c->total_recs += n_recs;
  PHASE_BEGIN("mark");
  TRY(segment_mark(c, *s));
  PHASE_END();
  return KMC_OK;
}

// ---- generic radix sort + RLE ---------------------------------------------------------------------------
// one stable counting-sort pass of n keys from src to dst by digit(key) < 256
template <typename KeyT, typename DigitFn>
int radix_pass(kmc_ctx *c, const KeyT *src, KeyT *dst, uint64_t n, DigitFn digit) {
  uint32_t n_blocks = grid_for(n, kRsTile);
  uint64_t m = (uint64_t)n_blocks * kRadix;
  TRY(ensure(c, c->block_hist, m * 4));
  TRY(ensure(c, c->offsets, m * 8));
  uint32_t scan_blocks = grid_for(m, kScanTile);
  TRY(ensure(c, c->sums, (size_t)std::max<uint64_t>(scan_blocks, grid_for(n, kRleTile)) * 8 + 64));
  auto rs_hist = rs_hist_kernel<KeyT, DigitFn>;
  auto rs_scatter = rs_scatter_kernel<KeyT, DigitFn>;
  LAUNCH(rs_hist, n_blocks, kRsThreads, 0, src, n, digit, (uint32_t *)c->block_hist.p, n_blocks);
  LAUNCH(scan_reduce_kernel, scan_blocks, kScanThreads, 0, (const uint32_t *)c->block_hist.p, m, (uint64_t *)c->sums.p);
  LAUNCH(scan_spine_kernel, 1, 1024, 0, (uint64_t *)c->sums.p, (uint64_t)scan_blocks);
  LAUNCH(scan_apply_kernel, scan_blocks, kScanThreads, 0, (const uint32_t *)c->block_hist.p, m,
         (const uint64_t *)c->sums.p, (uint64_t *)c->offsets.p);
  LAUNCH(rs_scatter, n_blocks, kRsThreads, 0, src, dst, n, digit, (const uint64_t *)c->offsets.p, n_blocks);
  return KMC_OK;
}

// sorts n keys in buffer `a` (scratch `b`) on bits [0,bits); *sorted receives the buffer with the result
template <typename KeyT>
int radix_sort(kmc_ctx *c, KeyT *a, KeyT *b, uint64_t n, uint32_t bits, KeyT **sorted) {
  *sorted = a;
  if (n <= 1 || bits == 0) return KMC_OK;
  KeyT *src = a, *dst = b;
  for (uint32_t shift = 0; shift < bits; shift += 8) {
    uint32_t nb = std::min<uint32_t>(8, bits - shift);
    TRY(radix_pass<KeyT, BitsDigit>(c, src, dst, n, BitsDigit{shift, nb}));
    std::swap(src, dst);
  }
  *sorted = src;
  return KMC_OK;
}

// sorted keys → table (t_lo, t_hi, t_cnt); `scratch` must hold (n+1) u64
template <typename KeyT>
int rle_to_table(kmc_ctx *c, const KeyT *sorted, uint64_t n, uint64_t *scratch) {
  c->n_total = n;
  c->n_distinct = 0;
  if (n == 0) return KMC_OK;
  uint32_t blocks = grid_for(n, kRleTile);
  TRY(ensure(c, c->sums, (size_t)blocks * 8 + 64));
  LAUNCH(rle_count_kernel<KeyT>, blocks, kRleThreads, 0, sorted, n, (uint64_t *)c->sums.p);
  // total = last block sum + its exclusive prefix: fetch both sides of the spine scan
  uint64_t last_cnt = 0, last_ex = 0;
  CK(cudaMemcpyAsync(&last_cnt, (uint64_t *)c->sums.p + (blocks - 1), 8, cudaMemcpyDeviceToHost, c->stream));
  LAUNCH(scan_spine_kernel, 1, 1024, 0, (uint64_t *)c->sums.p, (uint64_t)blocks);
  CK(cudaMemcpyAsync(&last_ex, (uint64_t *)c->sums.p + (blocks - 1), 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  uint64_t d = last_cnt + last_ex;
  c->n_distinct = d;
  TRY(ensure(c, c->t_lo, d * 8));
  if (sizeof(KeyT) == 16) TRY(ensure(c, c->t_hi, d * 8));
  TRY(ensure(c, c->t_cnt, d * 4));
  LAUNCH(rle_write_kernel<KeyT>, blocks, kRleThreads, 0, sorted, n, (const uint64_t *)c->sums.p, (uint64_t *)c->t_lo.p,
         sizeof(KeyT) == 16 ? (uint64_t *)c->t_hi.p : (uint64_t *)nullptr, scratch);
  LAUNCH(rle_diff_kernel, grid_for(d, 256), 256, 0, (const uint64_t *)scratch, d, n, (uint32_t *)c->t_cnt.p, d_err(c));
  return KMC_OK;
}

// ---- key production ----------------------------------------------------------------------------------------
// contiguous mode: all valid keys of all segments, compacted into keys_a; *n_keys on the host
template <typename KeyT>
int extract_all(kmc_ctx *c, uint64_t *n_keys) {
  uint64_t cap = std::max<uint64_t>(c->total_bases, 1) + 2;
  TRY(ensure(c, c->keys_a, cap * sizeof(KeyT)));
  TRY(zero_scalars(c));
  PHASE_BEGIN("extract");
  for (size_t i = 0; i < c->n_segs; i++) {
    Segment &s = c->segs[i];
    if (!s.n_bases) continue;
    TRY(seg_wait(c, s));
    ExtractParams P = seg_params(c, s);
    uint64_t tiles = num_warp_tiles(s.n_bases, win_lanes<KeyT>());
    uint32_t grid = (uint32_t)std::min<uint64_t>((tiles + 7) / 8, (uint64_t)c->n_sms * 16);
    auto extract_compact = extract_compact_kernel<KeyT, true>;
    LAUNCH(extract_compact, grid, 256, 0, P, tiles, (KeyT *)c->keys_a.p, d_cursor(c));
  }
  PHASE_END();
  TRY(read_scalars(c, n_keys, nullptr));
  return KMC_OK;
}

// lr-gapped mode: count + validate, then fill keys_a
template <typename KeyT>
int gapped_all(kmc_ctx *c, uint64_t *n_keys) {
  const kmc_config &f = c->cfg;
  TRY(zero_scalars(c));
  PHASE_BEGIN("extract");
  uint64_t mx = 1;
  for (size_t i = 0; i < c->n_segs; i++) mx = std::max<uint64_t>(mx, c->segs[i].n_bases);
  TRY(ensure(c, c->gap_l, mx * 8));
  TRY(ensure(c, c->gap_r, mx * 8));
  TRY(ensure(c, c->gap_f, mx));
  // pass 1: count and validate (per segment; the per-position arrays are reused)
  for (size_t i = 0; i < c->n_segs; i++) {
    Segment &s = c->segs[i];
    if (!s.n_bases) continue;
    TRY(seg_wait(c, s));
    GapParams P = gap_params(c, s);
    uint32_t g = grid_for(s.n_bases, 256);
    LAUNCH(gap_mers_kernel, g, 256, 0, P, (uint64_t *)c->gap_l.p, (uint64_t *)c->gap_r.p, (uint8_t *)c->gap_f.p);
    auto gap_pairs_count = gap_pairs_kernel<KeyT, false>;
    LAUNCH(gap_pairs_count, g, 256, 0, P, (const uint64_t *)c->gap_l.p, (const uint64_t *)c->gap_r.p,
           (const uint8_t *)c->gap_f.p, (KeyT *)nullptr, d_cursor(c), d_total_all(c), d_err(c));
  }
  uint64_t total = 0, total_all = 0; // keys of this part / of the whole input
  uint32_t err = 0;
  TRY(read_scalars(c, &total, &err, &total_all));
  if (err & 1) return fail(c, KMC_E_BADBASE, "Unexpected character (not one of A,C,G,T) inside an L/R chunk");
  if (err & 2) return fail(c, KMC_E_BADBASE_OFFSET0, "non-ACGT byte at offset 0 of an L/R chunk cannot be encoded");
  if (total_all == 0) return fail(c, KMC_E_EMPTY, "no L/R chunk in the input (every record shorter than %u bases)", f.d_min);
  TRY(ensure(c, c->keys_a, (total + 2) * sizeof(KeyT)));
  TRY(zero_scalars(c));
  for (size_t i = 0; i < c->n_segs; i++) {
    Segment &s = c->segs[i];
    if (!s.n_bases) continue;
    GapParams P = gap_params(c, s);
    uint32_t g = grid_for(s.n_bases, 256);
    if (c->n_segs > 1)
      LAUNCH(gap_mers_kernel, g, 256, 0, P, (uint64_t *)c->gap_l.p, (uint64_t *)c->gap_r.p, (uint8_t *)c->gap_f.p);
    auto gap_pairs_fill = gap_pairs_kernel<KeyT, true>;
    LAUNCH(gap_pairs_fill, g, 256, 0, P, (const uint64_t *)c->gap_l.p, (const uint64_t *)c->gap_r.p,
           (const uint8_t *)c->gap_f.p, (KeyT *)c->keys_a.p, d_cursor(c), d_total_all(c), d_err(c));
  }
  PHASE_END();
  *n_keys = total;
  return KMC_OK;
}

template <typename KeyT>
int produce_keys(kmc_ctx *c, uint64_t *n_keys) {
  if (!c->ingested.empty()) {
    uint64_t n = 0;
    for (auto &e : c->ingested) n += e.second;
    TRY(ensure(c, c->keys_a, (n + 2) * sizeof(KeyT)));
    PHASE_BEGIN("gather");
    uint64_t o = 0;
    for (auto &e : c->ingested) {
      if (e.second)
        CK(cudaMemcpyAsync((KeyT *)c->keys_a.p + o, e.first, e.second * sizeof(KeyT), cudaMemcpyDeviceToDevice, c->stream));
      o += e.second;
    }
    PHASE_END();
    *n_keys = n;
    return KMC_OK;
  }
  if (c->cfg.mode == KMC_MODE_LR_GAPPED) return gapped_all<KeyT>(c, n_keys);
  return extract_all<KeyT>(c, n_keys);
}

// ---- multi-GPU routing: group this rank's keys by owner part ---------------------------------
template <typename KeyT>
int route_impl(kmc_ctx *c, uint32_t n_parts, uint64_t *part_off) {
  uint64_t n = 0;
  if (c->cfg.mode == KMC_MODE_LR_GAPPED) TRY(gapped_all<KeyT>(c, &n));
  else TRY(extract_all<KeyT>(c, &n));
  TRY(ensure(c, c->route_keys, (n + 2) * sizeof(KeyT)));
  for (uint32_t p = 0; p <= n_parts; p++) part_off[p] = (p == n_parts) ? n : 0;
  if (n == 0) return KMC_OK;
  PHASE_BEGIN("route");
  TRY(radix_pass<KeyT, OwnerDigit>(c, (const KeyT *)c->keys_a.p, (KeyT *)c->route_keys.p, n, OwnerDigit{n_parts}));
  PHASE_END();
  // start of part p = scanned offset of (digit p, block 0)
  uint32_t n_blocks = grid_for(n, kRsTile);
  CK(cudaMemcpy2DAsync(part_off, 8, c->offsets.p, (size_t)n_blocks * 8, 8, n_parts, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  part_off[n_parts] = n;
  return KMC_OK;
}

// fast routing: the level-1 scatter of kmc_fast.cuh with bucket = owner part.  Each part gets
// a region sized from the upper bound (one key per base) with slack; *done = false → use the generic route.
// part_ptr == nullptr: the parts are regions of c->route_keys (kmc_route).  Otherwise part p is stored at
// part_ptr[p] — a peer's memory over NVLink (kmc_route_to_peers); positions are then absolute addresses / 8.
// chunk / n_chunks: route only that slice of the input's CTA tiles (the cursors carry on from chunk to chunk, the
// counts returned are cumulative): the owners work on chunk c while chunk c + 1 is on the links.  The routing pass has
// its own cursor / table buffers and its own flag word — an owner's count may be running in this ctx beside it.
// max_ctas: at most that many CTAs (one per SM), leaving the other SMs to the owner's kernels; 0 = all.
template <typename KeyT>
int route_fast(kmc_ctx *c, uint32_t n_parts, uint64_t *part_begin, uint64_t *part_count, bool *done,
               void *const *part_ptr = nullptr, uint64_t peer_cap = 0, uint32_t chunk = 0, uint32_t n_chunks = 1, uint32_t max_ctas = 0) {
  *done = false;
  if (!part_ptr && (c->total_bases < (1u << 18) || n_parts > (uint32_t)kMaxL1)) return KMC_OK;
  const uint64_t cap = part_ptr ? peer_cap : (((uint64_t)((double)c->total_bases / n_parts * 1.03) + 65536 + 15) & ~15ull);
  const uint64_t total = part_ptr ? 0 : cap * n_parts;
  TRY(ensure(c, c->route_keys, (total + 2 * kMaxTile) * sizeof(KeyT)));
  const size_t o_start = 0, o_cap = o_start + ((size_t)(n_parts + 1) * 8 + 15) / 16 * 16, tab_bytes = o_cap + (size_t)n_parts * 8;
  TRY(ensure(c, c->route_tables, tab_bytes));
  TRY(ensure(c, c->route_state, kMaxL1 * 8 + 64));
  if (chunk == 0) {
    std::vector<unsigned char> host(tab_bytes, 0);
    uint64_t *l1s = (uint64_t *)(host.data() + o_start), *l1c = (uint64_t *)(host.data() + o_cap);
    for (uint32_t p = 0; p <= n_parts; p++) l1s[p] = part_ptr ? (p < n_parts ? (uint64_t)(uintptr_t)part_ptr[p] / sizeof(KeyT) : 0) : cap * p;
    for (uint32_t p = 0; p < n_parts; p++) l1c[p] = cap;
    CK(cudaMemsetAsync(c->route_state.p, 0, kMaxL1 * 8, c->stream));
    CK(cudaMemsetAsync(d_route_err(c), 0, 8, c->stream));
    TRY(h2d_small(c, c->route_tables.p, host.data(), tab_bytes));
  }
  FastPlan pl{};
  pl.kb = c->key_bits; pl.b1 = 0; pl.n_l1 = n_parts; pl.n_fine = 0;
  pl.l1_trash = part_ptr ? (uint64_t)(uintptr_t)c->route_keys.p / sizeof(KeyT) : total;
  KeyT *dst = part_ptr ? (KeyT *)nullptr : (KeyT *)c->route_keys.p;
  pl.l1_start = (const uint64_t *)((unsigned char *)c->route_tables.p + o_start);
  pl.l1_cap = (const uint64_t *)((unsigned char *)c->route_tables.p + o_cap);
  pl.l1_cursor = (unsigned long long *)c->route_state.p;
  // this chunk's share of the CTA tiles of all segments, in segment order
  uint64_t all_ct = 0;
  for (size_t i = 0; i < c->n_segs; i++)
    if (c->segs[i].n_bases) all_ct += (num_warp_tiles(c->segs[i].n_bases, win_lanes<KeyT>()) + kFastWarps - 1) / kFastWarps;
  const uint64_t g0 = all_ct * chunk / n_chunks, g1 = all_ct * (chunk + 1) / n_chunks;
  PHASE_BEGIN("route");
  {
    size_t smem = L1Smem<KeyT>::bytes(part1_stage<KeyT>(), n_parts);
    auto fast_route = fast_part1_kernel<KeyT, true, OwnerBucket>;
    CK(cudaFuncSetAttribute(fast_route, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const OwnerBucket bucket{n_parts};
    uint64_t seg0 = 0;
    for (size_t i = 0; i < c->n_segs; i++) {
      Segment &s = c->segs[i];
      if (!s.n_bases) continue;
      const uint64_t tiles = num_warp_tiles(s.n_bases, win_lanes<KeyT>()), n_ct = (tiles + kFastWarps - 1) / kFastWarps;
      const uint64_t lo = std::max(g0, seg0), hi = std::min(g1, seg0 + n_ct);
      seg0 += n_ct;
      if (hi <= lo) continue;
      TRY(seg_wait(c, s));
      ExtractParams P = seg_params(c, s);
      uint32_t grid = (uint32_t)std::min<uint64_t>(hi - lo, (uint64_t)(max_ctas ? std::min<uint32_t>(max_ctas, c->n_sms) : c->n_sms));
      LAUNCH(fast_route, grid, kFastThreads, smem, P, tiles, pl, bucket, dst, d_route_err(c), lo - (seg0 - n_ct), hi - (seg0 - n_ct));
    }
  }
  PHASE_END();
  std::vector<unsigned long long> cur(n_parts);
  unsigned long long rerr = 0;
  TRY(d2h_small(c, cur.data(), pl.l1_cursor, (size_t)n_parts * 8, 0));
  TRY(d2h_small(c, &rerr, d_route_err(c), 8, 8192));
  if ((uint32_t)rerr & kFlagOverflow) {
    CK(cudaMemsetAsync(d_route_err(c), 0, 8, c->stream));
    // peers: not an error here — the counts tell the caller (some exceed part_cap_keys), who must agree with the other
    // ranks on a larger capacity and route again
    if (part_ptr) { for (uint32_t p = 0; p < n_parts; p++) part_count[p] = cur[p]; *done = true; }
    return KMC_OK;
  }
  for (uint32_t p = 0; p < n_parts; p++) { if (part_begin) part_begin[p] = cap * p; part_count[p] = cur[p]; }
  *done = true;
  return KMC_OK;
}

// ---- strategies --------------------------------------------------------------------------------------------
template <typename KeyT>
int finish_baseline(kmc_ctx *c) {
  uint64_t n = 0;
  TRY(produce_keys<KeyT>(c, &n));
  c->strategy_used = KMC_STRATEGY_SORT_BASELINE;
  TRY(ensure(c, c->keys_b, (n + 2) * sizeof(KeyT)));
  KeyT *sorted = nullptr;
  PHASE_BEGIN("sort");
  TRY(radix_sort<KeyT>(c, (KeyT *)c->keys_a.p, (KeyT *)c->keys_b.p, n, c->key_bits, &sorted));
  PHASE_END();
  PHASE_BEGIN("rle");
  uint64_t *scratch = (uint64_t *)(sorted == (KeyT *)c->keys_a.p ? c->keys_b.p : c->keys_a.p);
  TRY(rle_to_table<KeyT>(c, sorted, n, scratch));
  PHASE_END();
  return KMC_OK;
}


// 64-bit key sources of a job: either the submitted segments (keys are extracted on the fly) or key arrays
// (ingested keys are used in place; lr-gapped keys are materialised first).
struct KeyArrays {
  bool from_array = false;
  uint64_t n = 0;
  std::vector<std::pair<const void *, uint64_t>> arrays;
};
template <typename KeyT>
int key_sources(kmc_ctx *c, KeyArrays *ka) {
  ka->from_array = !c->ingested.empty() || c->cfg.mode == KMC_MODE_LR_GAPPED;
  ka->n = 0; ka->arrays.clear();
  if (!ka->from_array) return KMC_OK;
  if (!c->ingested.empty()) {
    for (auto &e : c->ingested) if (e.second) { ka->arrays.emplace_back(e.first, e.second); ka->n += e.second; }
  } else {
    TRY(produce_keys<KeyT>(c, &ka->n));
    if (ka->n) ka->arrays.emplace_back((const void *)c->keys_a.p, ka->n);
  }
  return KMC_OK;
}

#include "kmc_api_hash.cuh"   // hash strategies, cardinality probe
#include "kmc_api_fast.cuh"   // partitioned path, kept key array
#include "kmc_api_dist.cuh"   // range partition, streaming owner

template <typename KeyT>
int finish_impl(kmc_ctx *c) {
  // keys already sit in my level-1 array — unless the job went on through the hash route afterwards (a peer's scatter
  // overflowed: every rank recounts, and the hash route's peer stores have overwritten the receive buffer)
  if (c->ingested.empty() && c->dist.valid && c->dist.scattered) return finish_dist<KeyT>(c);
  const uint32_t strat = c->cfg.strategy;
  if (strat != KMC_STRATEGY_SORT_BASELINE) {
    bool used = false;
    if constexpr (sizeof(KeyT) == 8) {
      if (strat == KMC_STRATEGY_HASH) {
        // forced: size the table for the worst case (every key distinct), at most 2^32 slots
        uint64_t n_in = 0;
        for (auto &e : c->ingested) n_in += e.second;
        if (c->ingested.empty()) n_in = c->cfg.mode == KMC_MODE_LR_GAPPED ? c->total_bases * (c->cfg.d_max - c->cfg.d_min + 1) : c->total_bases;
        bool low = false;
        TRY(hash_probe(c, &low)); // also finds the hot keys
        uint32_t lg = 20;
        while (lg < 32 && (1ull << lg) < 2 * n_in) lg++;
        TRY(finish_hash(c, lg, (1ull << lg) / 10 * 7, &used));
        if (used) return KMC_OK;
      } else if (strat == KMC_STRATEGY_AUTO) {
        bool low = false;
        TRY(hash_probe(c, &low));
        if (low) {
          // a table sized from the sample stays in L2 (a 2^25-slot one does not: 512 MB of randomly touched lines);
          // if the full input has more distinct keys than that, retry once with the big table
          uint32_t lg = 20;
          while (lg < 25 && (1ull << lg) < 4 * c->probe_distinct) lg++;
          TRY(finish_hash(c, lg, (1ull << lg) / 2, &used));
          if (used) return KMC_OK;
          if (lg < 25) {
            TRY(finish_hash(c, 25, 1ull << 24, &used));
            if (used) return KMC_OK;
          }
        }
      }
    }
    if constexpr (sizeof(KeyT) == 16) {
      if (strat == KMC_STRATEGY_HASH) {
        // forced: size the table for the worst case (every key distinct), at most 2^31 slots of 32 bytes
        uint64_t n_in = 0;
        for (auto &e : c->ingested) n_in += e.second;
        if (c->ingested.empty()) n_in = c->cfg.mode == KMC_MODE_LR_GAPPED ? c->total_bases * (c->cfg.d_max - c->cfg.d_min + 1) : c->total_bases;
        c->probe_distinct = 0;
        uint32_t lg = 20;
        while (lg < 31 && (1ull << lg) < 2 * n_in) lg++;
        TRY(finish_hash128(c, lg, (1ull << lg) / 10 * 7, &used));
        if (used) return KMC_OK;
      } else if (strat == KMC_STRATEGY_AUTO) {
        bool low = false;
        TRY(hash128_probe(c, &low));
        if (low) { // a table sized from the sample, as for 64-bit keys; once more with a big one if it fills
          uint32_t lg = 20;
          while (lg < 24 && (1ull << lg) < 4 * c->probe_distinct) lg++;
          TRY(finish_hash128(c, lg, (1ull << lg) / 2, &used));
          if (used) return KMC_OK;
          if (lg < 24) {
            TRY(finish_hash128(c, 24, 1ull << 23, &used));
            if (used) return KMC_OK;
          }
        }
      }
    }
    const uint32_t fallbacks = c->fast_fallbacks;
    TRY(finish_fast<KeyT>(c, &used));
    if (used) return KMC_OK;
    if (c->fast_fallbacks != fallbacks) { // a bucket overflowed: once more with half-full buckets of full capacity
      TRY(finish_fast<KeyT>(c, &used, 1));
      if (used) return KMC_OK;
    }
  }
  return finish_baseline<KeyT>(c);
}

// exclusive scan of m u32 counts into u64 offsets (three kernels, uses c->sums)
int scan_u32(kmc_ctx *c, const uint32_t *in, uint64_t m, uint64_t *out) {
  uint32_t blocks = grid_for(m, kScanTile);
  TRY(ensure(c, c->sums, (size_t)blocks * 8 + 64));
  LAUNCH(scan_reduce_kernel, blocks, kScanThreads, 0, in, m, (uint64_t *)c->sums.p);
  LAUNCH(scan_spine_kernel, 1, 1024, 0, (uint64_t *)c->sums.p, (uint64_t)blocks);
  LAUNCH(scan_apply_kernel, blocks, kScanThreads, 0, in, m, (const uint64_t *)c->sums.p, out);
  return KMC_OK;
}

void build_stats(kmc_ctx *c) {
  std::string s = "{";
  char buf[1024];
  snprintf(buf, sizeof buf,
           "\"n_bases\": %llu, \"n_records\": %llu, \"n_total\": %llu, \"n_distinct\": %llu, \"key_bits\": %u, "
           "\"strategy_used\": %u, \"fast_variant\": \"%s\", \"fast_fallbacks\": %u, \"hash_aborts\": %u, \"hot_keys\": %u, \"kernel_launches\": %llu, \"kernel_launches_total\": %llu, \"h2d_bytes\": %llu, \"phases_ms\": {",
           (unsigned long long)c->total_bases, (unsigned long long)c->total_recs, (unsigned long long)c->n_total,
           (unsigned long long)c->n_distinct, c->key_bits, c->strategy_used, c->fast_variant, c->fast_fallbacks, c->hash_aborts, c->n_hot, (unsigned long long)c->launches,
           (unsigned long long)c->launches_total, (unsigned long long)c->h2d_bytes);
  s += buf;
  // sum phases of the same name
  std::vector<std::pair<std::string, float>> agg;
  for (auto &p : c->phases) {
    bool found = false;
    for (auto &a : agg) if (a.first == p.name) { a.second += p.ms; found = true; }
    if (!found) agg.emplace_back(p.name, p.ms);
  }
  for (size_t i = 0; i < agg.size(); i++) {
    snprintf(buf, sizeof buf, "%s\"%s\": %.4f", i ? ", " : "", agg[i].first.c_str(), agg[i].second);
    s += buf;
  }
  s += "}, \"kernels\": {";
  for (size_t i = 0; i < c->kstats.size(); i++) {
    snprintf(buf, sizeof buf, "%s\"%s\": {\"launches\": %llu, \"ms\": %.4f}", i ? ", " : "", c->kstats[i].name.c_str(),
             (unsigned long long)c->kstats[i].launches, c->kstats[i].ms);
    s += buf;
  }
  s += "}}";
  c->stats = s;
}

} // namespace

// ================================================================================================ C ABI
// a range-partitioned count in progress is given up (the job goes on through another route, or the ctx is reset): its
// owner-side kernels may still be reading the receive buffer
static void dist_abandon(kmc_ctx *c) {
  if (c->dist.owner_ready && c->owner_stream) cudaStreamSynchronize(c->owner_stream);
  if (c->dist.chunks_sent && c->peer_stream) cudaStreamSynchronize(c->peer_stream);
  c->dist.valid = false; c->dist.scattered = false; c->dist.owner_ready = false; c->dist.chunks_sent = 0; c->dist.chunks_owned = 0;
}

extern "C" {

const char *kmc_strerror(int code) {
  switch (code) {
    case KMC_OK: return "ok";
    case KMC_E_ARG: return "bad argument or call order";
    case KMC_E_NO_DEVICE: return "no usable CUDA device";
    case KMC_E_CUDA: return "CUDA runtime error";
    case KMC_E_NOMEM: return "out of memory";
    case KMC_E_BADBASE: return "unexpected character in an L/R chunk";
    case KMC_E_EMPTY: return "no L/R chunk in the input";
    case KMC_E_COUNT_OVERFLOW: return "a count exceeds 32 bits";
    case KMC_E_CAPACITY: return "capacity exceeded";
    case KMC_E_BADBASE_OFFSET0: return "non-ACGT byte at offset 0 of an L/R chunk";
    case KMC_E_FORMAT: return "FASTA text does not start with '>'";
    default: return "unknown error";
  }
}

const char *kmc_last_error(const kmc_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int kmc_create(kmc_ctx **out, const kmc_config *cfg) {
  kmc_ctx *c = nullptr; // for CK/fail: errors go to g_create_err until the ctx exists
  if (!out || !cfg) return fail(c, KMC_E_ARG, "null argument");
  *out = nullptr;
  if (cfg->abi_version != KMC_ABI_VERSION) return fail(c, KMC_E_ARG, "abi_version %u != %u", cfg->abi_version, KMC_ABI_VERSION);
  kmc_config f = *cfg;
  for (uint32_t r : f.reserved) if (r) return fail(c, KMC_E_ARG, "reserved fields must be zero");
  if (f.mode == KMC_MODE_CONTIGUOUS) {
    if (f.k < 1 || f.k > 64) return fail(c, KMC_E_ARG, "k must be 1..64 (got %u)", f.k);
  } else if (f.mode == KMC_MODE_LR_GAPPED) {
    if (!f.l_len && !f.r_len) { f.l_len = 27; f.r_len = 27; }       // main.rs:48-49
    if (!f.d_min && !f.d_max) { f.d_min = 80; f.d_max = 140; }      // main.rs:63
    if (f.canonical) return fail(c, KMC_E_ARG, "lr-gapped mode is forward-strand only (main.rs:76-78)");
    if (f.l_len < 1 || f.l_len > 32 || f.r_len < 1 || f.r_len > 32) return fail(c, KMC_E_ARG, "l_len and r_len must be 1..32");
    if (f.d_min < f.l_len + f.r_len || f.d_max < f.d_min) return fail(c, KMC_E_ARG, "need l_len+r_len <= d_min <= d_max");
  } else {
    return fail(c, KMC_E_ARG, "unknown mode %u", f.mode);
  }
  if (f.strategy > KMC_STRATEGY_SORT_BASELINE) return fail(c, KMC_E_ARG, "unknown strategy %u", f.strategy);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(c, KMC_E_NO_DEVICE, "no CUDA device: libkmc has no CPU fallback");
  }
  int dev = f.device;
  if (dev < 0) CK(cudaGetDevice(&dev));
  if (dev >= ndev) return fail(c, KMC_E_NO_DEVICE, "device %d of %d", dev, ndev);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(c, KMC_E_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
  CK(cudaSetDevice(dev));
  kmc_ctx *x = new (std::nothrow) kmc_ctx();
  if (!x) return fail(c, KMC_E_NOMEM, "host allocation failed");
  x->cfg = f;
  x->device = dev;
  x->n_sms = (uint32_t)prop.multiProcessorCount;
  x->key_bases = f.mode == KMC_MODE_CONTIGUOUS ? f.k : f.l_len + f.r_len;
  x->key_bits = 2 * x->key_bases;
  x->wide = x->key_bits > 64;
  cudaError_t e = cudaStreamCreateWithFlags(&x->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete x; return fail(c, KMC_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
  x->own_stream = true;
  { const char *kt = getenv("KMC_KERNEL_TIMING"); x->ktiming = kt && kt[0] == '1'; }
  for (int i = 0; i < 2; i++) cudaEventCreateWithFlags(&x->copy_done[i], cudaEventDisableTiming);
  *out = x;
  return KMC_OK;
}

void kmc_destroy(kmc_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (int i = 0; i < 2; i++) {
    if (c->h_bases[i]) cudaFreeHost(c->h_bases[i]);
    if (c->h_off[i]) cudaFreeHost(c->h_off[i]);
    if (c->copy_done[i]) cudaEventDestroy(c->copy_done[i]);
  }
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  if (c->mailbox) cudaFreeHost(c->mailbox);
  if (c->upbox) cudaFreeHost(c->upbox);
  if (c->fmt_host) cudaFreeHost(c->fmt_host);
  for (auto &s : c->segs) { release(s.own_bases); release(s.own_off); release(s.brk); if (s.ready) cudaEventDestroy(s.ready); }
  for (DevBuf *b : {&c->keys_a, &c->keys_b, &c->block_hist, &c->offsets, &c->sums, &c->scalars, &c->route_keys, &c->gap_l,
                    &c->gap_r, &c->gap_f, &c->t_lo, &c->t_hi, &c->t_cnt, &c->fast_l1, &c->fast_l2, &c->fast_state, &c->fast_tables, &c->fast_fdesc, &c->recv_keys, &c->hash_slots, &c->hash_scalars, &c->hash_hot, &c->fa_raw, &c->fa_tiles, &c->fa_flags, &c->fmt_len, &c->fmt_off, &c->fmt_text, &c->merge_lo, &c->merge_hi, &c->merge_cnt, &c->pair_rows, &c->pair_state,
                    &c->kept_keys, &c->kept_tables, &c->kept_state, &c->dist_tables, &c->dist_stage, &c->dist_cursors, &c->route_state, &c->route_tables})
    release(*b);
  for (auto ev : c->event_pool) cudaEventDestroy(ev);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  if (c->owner_stream) cudaStreamDestroy(c->owner_stream);
  if (c->peer_stream) cudaStreamDestroy(c->peer_stream);
  for (int i = 0; i < 8; i++) { if (c->peer_lane[i]) cudaStreamDestroy(c->peer_lane[i]); if (c->peer_lane_ev[i]) cudaEventDestroy(c->peer_lane_ev[i]); }
  if (c->dist.ev_ready) cudaEventDestroy(c->dist.ev_ready);
  for (int i = 0; i < 16; i++) { if (c->dist.ev_scattered[i]) cudaEventDestroy(c->dist.ev_scattered[i]); if (c->dist.ev_copied[i]) cudaEventDestroy(c->dist.ev_copied[i]); }
  delete static_cast<FastJob *>(c->job_box);
  delete c;
}

int kmc_set_stream(kmc_ctx *c, void *s) {
  if (!c) return KMC_E_ARG;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  if (c->own_stream) { cudaStreamDestroy(c->stream); c->own_stream = false; }
  c->stream = (cudaStream_t)s;
  return KMC_OK;
}

int kmc_reset(kmc_ctx *c) {
  if (!c) return KMC_E_ARG;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  if (c->copy_stream) CK(cudaStreamSynchronize(c->copy_stream));
  if (c->owner_stream) CK(cudaStreamSynchronize(c->owner_stream));
  c->owner_on = false; c->owner_fed.clear();
  c->n_segs = 0; c->total_bases = c->total_recs = 0;
  c->ingested.clear();
  c->ingested_pairs.clear();
  c->finished = false; c->n_total = c->n_distinct = 0;
  c->range_on = false; c->part_hist_step = 0;
  c->kept_valid = false; c->kept_tried = false;
  if (c->kept_keys.cap > ((size_t)1 << 30)) release(c->kept_keys); // the one work buffer as large as the input's keys: not kept across jobs
  dist_abandon(c);
  c->phases.clear(); c->klaunches.clear(); c->events_used = 0;
  c->launches_total += c->launches; c->launches = 0; c->h2d_bytes = 0;
  c->staged = false;
  return KMC_OK;
}

uint32_t kmc_key_bases(const kmc_ctx *c) { return c ? c->key_bases : 0; }

int kmc_staging(kmc_ctx *c, size_t want_bases, size_t want_recs, uint8_t **bases, uint64_t **rec_off, size_t *cap_bases,
                size_t *cap_recs) {
  if (!c || !bases || !rec_off) return KMC_E_ARG;
  CK(cudaSetDevice(c->device));
  want_bases = std::max<size_t>(want_bases, 1 << 20);
  want_recs = std::max<size_t>(want_recs, 1 << 12);
  if (want_bases > c->cap_bases || want_recs > c->cap_recs) {
    CK(cudaStreamSynchronize(c->stream));
    size_t nb = std::max(want_bases, c->cap_bases), nr = std::max(want_recs, c->cap_recs);
    for (int i = 0; i < 2; i++) {
      if (c->h_bases[i]) CK(cudaFreeHost(c->h_bases[i]));
      if (c->h_off[i]) CK(cudaFreeHost(c->h_off[i]));
      c->h_bases[i] = nullptr; c->h_off[i] = nullptr;
      CK(cudaHostAlloc((void **)&c->h_bases[i], nb, cudaHostAllocDefault));
      CK(cudaHostAlloc((void **)&c->h_off[i], (nr + 1) * 8, cudaHostAllocDefault));
      c->copy_pending[i] = false;
    }
    c->cap_bases = nb; c->cap_recs = nr;
  }
  int i = c->cur;
  if (c->copy_pending[i]) { CK(cudaEventSynchronize(c->copy_done[i])); c->copy_pending[i] = false; }
  *bases = c->h_bases[i]; *rec_off = c->h_off[i];
  if (cap_bases) *cap_bases = c->cap_bases;
  if (cap_recs) *cap_recs = c->cap_recs;
  c->staged = true;
  return KMC_OK;
}

int kmc_submit(kmc_ctx *c, size_t n_bases, size_t n_recs) {
  if (!c) return KMC_E_ARG;
  if (!c->staged) return fail(c, KMC_E_ARG, "kmc_submit without kmc_staging");
  if (n_bases > c->cap_bases || n_recs > c->cap_recs) return fail(c, KMC_E_CAPACITY, "submit larger than the staging buffer");
  CK(cudaSetDevice(c->device));
  int i = c->cur;
  if (c->h_off[i][0] != 0 || c->h_off[i][n_recs] != n_bases) return fail(c, KMC_E_ARG, "rec_off[0] must be 0 and rec_off[n_recs] == n_bases");
  TRY(submit_from_host(c, c->h_bases[i], c->h_off[i], n_bases, n_recs, c->copy_done[i]));
  c->copy_pending[i] = true;
  c->cur ^= 1;
  c->staged = false;
  return KMC_OK;
}

int kmc_submit_host(kmc_ctx *c, const uint8_t *bases, const uint64_t *rec_off, size_t n_bases, size_t n_recs) {
  if (!c || !rec_off || (!bases && n_bases)) return KMC_E_ARG;
  if (rec_off[0] != 0 || rec_off[n_recs] != n_bases) return fail(c, KMC_E_ARG, "rec_off[0] must be 0 and rec_off[n_recs] == n_bases");
  CK(cudaSetDevice(c->device));
  const size_t segs_before = c->n_segs;
  TRY(submit_from_host(c, bases, rec_off, n_bases, n_recs, nullptr));
  // pageable memory was copied synchronously enough to be reused; a chunked (pinned) submit stays asynchronous and
  // the caller's buffer must stay unchanged until kmc_finish returns
  if (!(c->n_segs > segs_before && c->segs[segs_before].wait_ready)) CK(cudaStreamSynchronize(c->stream));
  return KMC_OK;
}

int kmc_submit_fasta(kmc_ctx *c, const uint8_t *text, size_t n, uint64_t *n_bases_out, uint64_t *n_recs_out) {
  if (!c || (!text && n)) return KMC_E_ARG;
  if (c->finished) return fail(c, KMC_E_ARG, "kmc_submit_fasta after kmc_finish (call kmc_reset first)");
  if (n && text[0] != '>') return fail(c, KMC_E_FORMAT, "Expected > at record start.");
  CK(cudaSetDevice(c->device));
  Segment *s;
  TRY(new_segment(c, &s));
  uint64_t n_seq = 0, n_hdr = 0;
  if (n) {
    const uint64_t tiles = (n + kFaTile - 1) / kFaTile;
    // tile arrays: prev_nl i64 | seq_off u64 | hdr_off u64 | tile_seq u32 | tile_hdr u32
    TRY(ensure(c, c->fa_raw, n + 64));
    TRY(ensure(c, c->fa_tiles, tiles * 32 + 256));
    long long *prev_nl = (long long *)c->fa_tiles.p;
    uint64_t *seq_off = (uint64_t *)(prev_nl + tiles), *hdr_off = seq_off + tiles;
    uint32_t *tile_seq = (uint32_t *)(hdr_off + tiles), *tile_hdr = tile_seq + tiles;
    PHASE_BEGIN("h2d");
    CK(cudaMemcpyAsync(c->fa_raw.p, text, n, cudaMemcpyHostToDevice, c->stream));
    PHASE_END();
    c->h2d_bytes += n;
    PHASE_BEGIN("fasta_parse");
    const uint8_t *raw = (const uint8_t *)c->fa_raw.p;
    LAUNCH(fasta_lastnl_kernel, (uint32_t)tiles, kFaThreads, 0, raw, (uint64_t)n, prev_nl);
    LAUNCH(fasta_scan_nl_kernel, 1, 1024, 0, prev_nl, tiles);
    LAUNCH(fasta_count_kernel, (uint32_t)tiles, kFaThreads, 0, raw, (uint64_t)n, (const long long *)prev_nl, tile_seq, tile_hdr);
    TRY(scan_u32(c, tile_seq, tiles, seq_off));
    TRY(scan_u32(c, tile_hdr, tiles, hdr_off));
    uint64_t last_off[2];
    uint32_t last_cnt[2];
    TRY(d2h_small(c, &last_off[0], seq_off + (tiles - 1), 8, 0));
    TRY(d2h_small(c, &last_off[1], hdr_off + (tiles - 1), 8, 64));
    TRY(d2h_small(c, &last_cnt[0], tile_seq + (tiles - 1), 4, 128));
    TRY(d2h_small(c, &last_cnt[1], tile_hdr + (tiles - 1), 4, 192));
    n_seq = last_off[0] + last_cnt[0];
    n_hdr = last_off[1] + last_cnt[1];
    TRY(ensure(c, s->own_bases, n_seq + 64));
    TRY(ensure(c, s->own_off, (n_hdr + 2) * 8));
    TRY(ensure(c, c->fa_flags, n_hdr + 64));
    LAUNCH(fasta_write_kernel, (uint32_t)tiles, kFaThreads, 0, raw, (uint64_t)n, (const long long *)prev_nl, (const uint64_t *)seq_off,
           (const uint64_t *)hdr_off, (uint8_t *)s->own_bases.p, (uint64_t *)s->own_off.p, (uint8_t *)c->fa_flags.p);
    CK(cudaMemcpyAsync((uint64_t *)s->own_off.p + n_hdr, &n_seq, 8, cudaMemcpyHostToDevice, c->stream));
    // main.rs:60-62: stop at the first record that is entirely empty
    TRY(zero_scalars(c));
    CK(cudaMemsetAsync(d_cursor(c), 0xFF, 8, c->stream));
    if (n_hdr)
      LAUNCH(fasta_first_empty_kernel, grid_for(n_hdr, 256), 256, 0, (const uint64_t *)s->own_off.p, (const uint8_t *)c->fa_flags.p, n_hdr, d_cursor(c));
    uint64_t first_empty = ~0ull;
    TRY(read_scalars(c, &first_empty, nullptr));
    if (first_empty < n_hdr) {
      n_hdr = first_empty;
      TRY(d2h_small(c, &n_seq, (uint64_t *)s->own_off.p + n_hdr, 8));
    }
    PHASE_END();
  } else {
    TRY(ensure(c, s->own_off, 64));
    CK(cudaMemsetAsync(s->own_off.p, 0, 8, c->stream));
    TRY(ensure(c, s->own_bases, 64));
  }
  s->bases = (const uint8_t *)s->own_bases.p;
  s->rec_off = (const uint64_t *)s->own_off.p;
  s->n_bases = n_seq; s->n_recs = n_hdr;
  c->total_bases += n_seq; c->total_recs += n_hdr;
  PHASE_BEGIN("mark");
  TRY(segment_mark(c, *s));
  PHASE_END();
  CK(cudaStreamSynchronize(c->stream)); // the caller's text may be pageable / reused right away
  if (n_bases_out) *n_bases_out = n_seq;
  if (n_recs_out) *n_recs_out = n_hdr;
  return KMC_OK;
}

int kmc_submit_device(kmc_ctx *c, const uint8_t *d_bases, const uint64_t *d_rec_off, size_t n_bases, size_t n_recs) {
  if (!c || !d_rec_off || (!d_bases && n_bases)) return KMC_E_ARG;
  if (((uintptr_t)d_bases & 15) != 0) return fail(c, KMC_E_ARG, "d_bases must be 16-byte aligned");
  if (c->finished) return fail(c, KMC_E_ARG, "kmc_submit_device after kmc_finish (call kmc_reset first)");
  CK(cudaSetDevice(c->device));
  Segment *s;
  TRY(new_segment(c, &s));
  s->bases = d_bases; s->rec_off = d_rec_off; s->n_bases = n_bases; s->n_recs = n_recs;
  c->total_bases += n_bases; c->total_recs += n_recs;
  PHASE_BEGIN("mark");
  TRY(segment_mark(c, *s));
  PHASE_END();
  return KMC_OK;
}

int kmc_ingest_keys(kmc_ctx *c, const void *d_keys, uint64_t n_keys) {
  if (!c || (!d_keys && n_keys)) return KMC_E_ARG;
  if (c->finished) return fail(c, KMC_E_ARG, "kmc_ingest_keys after kmc_finish");
  dist_abandon(c); // a range-partition scatter of this job, if any, is abandoned
  c->ingested.emplace_back(d_keys, n_keys);
  return KMC_OK;
}

int kmc_ingest_pairs(kmc_ctx *c, const uint64_t *d_keys, const uint64_t *d_counts, uint64_t n_rows) {
  if (!c || ((!d_keys || !d_counts) && n_rows)) return KMC_E_ARG;
  if (c->finished) return fail(c, KMC_E_ARG, "kmc_ingest_pairs after kmc_finish (call kmc_reset first)");
  if (c->wide) return fail(c, KMC_E_ARG, "kmc_ingest_pairs: 64-bit keys only");
  if (c->n_segs || !c->ingested.empty()) return fail(c, KMC_E_ARG, "kmc_ingest_pairs cannot be mixed with submitted input or ingested keys");
  c->ingested_pairs.push_back({d_keys, d_counts, n_rows});
  return KMC_OK;
}

int kmc_table_route(kmc_ctx *c, uint32_t n_parts, uint64_t *part_begin, uint64_t *part_count, const uint64_t **d_keys,
                    const uint64_t **d_counts) {
  if (!c || !part_begin || !part_count || !d_keys || !d_counts) return KMC_E_ARG;
  if (!c->finished) return fail(c, KMC_E_ARG, "kmc_table_route before kmc_finish");
  if (c->wide) return fail(c, KMC_E_ARG, "kmc_table_route: 64-bit keys only");
  if (n_parts < 1 || n_parts > 1024) return fail(c, KMC_E_ARG, "kmc_table_route: n_parts must be 1..1024");
  CK(cudaSetDevice(c->device));
  const uint64_t rows = c->n_distinct;
  TRY(ensure(c, c->pair_rows, 2 * (rows + 2) * 8));
  TRY(ensure(c, c->pair_state, 1024 * 8));
  uint64_t *out_keys = (uint64_t *)c->pair_rows.p, *out_counts = out_keys + (rows + 2);
  unsigned long long *state = (unsigned long long *)c->pair_state.p;
  std::vector<unsigned long long> h(n_parts, 0);
  if (rows) {
    CK(cudaMemsetAsync(state, 0, (size_t)n_parts * 8, c->stream));
    const uint32_t grid = std::min<uint32_t>(grid_for(rows, 1024), c->n_sms * 8);
    LAUNCH(table_owner_hist_kernel, grid, 256, 0, (const uint64_t *)c->t_lo.p, rows, n_parts, state);
    TRY(d2h_small(c, h.data(), state, (size_t)n_parts * 8));
    std::vector<unsigned long long> begin(n_parts, 0);
    for (uint32_t p = 1; p < n_parts; p++) begin[p] = begin[p - 1] + h[p - 1];
    TRY(h2d_small(c, state, begin.data(), (size_t)n_parts * 8));
    LAUNCH(table_owner_scatter_kernel, grid, 256, 0, (const uint64_t *)c->t_lo.p, (const uint32_t *)c->t_cnt.p, rows, n_parts, state,
           out_keys, out_counts);
    CK(cudaStreamSynchronize(c->stream));
  }
  uint64_t b = 0;
  for (uint32_t p = 0; p < n_parts; p++) { part_begin[p] = b; part_count[p] = h[p]; b += h[p]; }
  *d_keys = out_keys; *d_counts = out_counts;
  return KMC_OK;
}

static int finish_common(kmc_ctx *c, uint64_t *n_distinct, uint64_t *n_total) {
  const double host_t0 = host_now_ms();
  const size_t first_phase = c->phases.size();
  c->host_marks.clear();
  if (!c->ingested_pairs.empty() && (c->n_segs || !c->ingested.empty()))
    return fail(c, KMC_E_ARG, "kmc_finish: (key, count) rows cannot be mixed with submitted input or ingested keys");
  int rc;
  if (c->owner_on) {
    rc = c->wide ? owner_finish<U128>(c) : owner_finish<uint64_t>(c);
  } else if (c->ingested.empty() && c->dist.valid && c->dist.scattered) {
    // range partition: the owner's level-2 scatter may already be running; the flag word is live
    rc = c->wide ? finish_dist<U128>(c) : finish_dist<uint64_t>(c);
  } else {
    TRY(zero_scalars(c));
    rc = !c->ingested_pairs.empty() ? finish_pairs(c) : c->wide ? finish_impl<U128>(c) : finish_impl<uint64_t>(c);
  }
  if (rc) return rc;
  uint32_t err = 0;
  TRY(read_scalars(c, nullptr, &err));
  if (err & 4) return fail(c, KMC_E_COUNT_OVERFLOW, "a k-mer occurs more than 2^32-1 times");
  for (auto &p : c->phases) cudaEventElapsedTime(&p.ms, p.a, p.b);
  c->kstats.clear();
  for (auto &p : c->klaunches) {
    cudaEventElapsedTime(&p.ms, p.a, p.b);
    KernelStat *ks = nullptr;
    for (auto &k : c->kstats) if (k.name == p.name) ks = &k;
    if (!ks) { c->kstats.emplace_back(); ks = &c->kstats.back(); ks->name = p.name; }
    ks->launches++; ks->ms += p.ms;
  }
  // development aid: KMC_HOST_PROF=<ms> prints the host-side timeline of a kmc_finish that took longer than that
  static const double host_prof = getenv("KMC_HOST_PROF") ? atof(getenv("KMC_HOST_PROF")) : -1.0;
  if (host_prof >= 0 && host_now_ms() - host_t0 > host_prof) {
    fprintf(stderr, "[kmc host] finish %.2f ms:", host_now_ms() - host_t0);
    for (size_t i = first_phase; i < c->phases.size(); i++)
      fprintf(stderr, " %s %.2f..%.2f (gpu %.2f)", c->phases[i].name.c_str(), c->phases[i].host_begin - host_t0,
              c->phases[i].host_end - host_t0, c->phases[i].ms);
    for (auto &m : c->host_marks) fprintf(stderr, " [%s %.2f]", m.first, m.second - host_t0);
    fprintf(stderr, "\n");
  }
  c->finished = true;
  if (n_distinct) *n_distinct = c->n_distinct;
  if (n_total) *n_total = c->n_total;
  return KMC_OK;
}

int kmc_finish(kmc_ctx *c, uint64_t *n_distinct, uint64_t *n_total) {
  if (!c) return KMC_E_ARG;
  if (c->finished) return fail(c, KMC_E_ARG, "kmc_finish called twice");
  CK(cudaSetDevice(c->device));
  return finish_common(c, n_distinct, n_total);
}

int kmc_finish_part(kmc_ctx *c, uint32_t part, uint32_t n_parts, uint64_t *n_distinct, uint64_t *n_total) {
  if (!c) return KMC_E_ARG;
  const uint32_t ncoarse = 1u << coarse_bits(c);
  if (n_parts == 0 || part >= n_parts) return fail(c, KMC_E_ARG, "kmc_finish_part: need part < n_parts");
  if (n_parts > ncoarse) return fail(c, KMC_E_ARG, "kmc_finish_part: at most %u parts for %u-bit keys", ncoarse, c->key_bits);
  if (!c->ingested.empty()) return fail(c, KMC_E_ARG, "kmc_finish_part: ingested keys are one part of a multi-GPU job already");
  CK(cudaSetDevice(c->device));
  if (c->finished) { // the previous part's table goes, the submitted input stays
    CK(cudaStreamSynchronize(c->stream));
    c->finished = false; c->n_total = c->n_distinct = 0;
    c->phases.clear(); c->klaunches.clear(); c->events_used = 0;
  }
  if (n_parts == 1) return finish_common(c, n_distinct, n_total);
  if (!c->part_hist_step) { // once per input: where the keys are
    c->range_on = false;
    if (c->cfg.mode == KMC_MODE_LR_GAPPED) {
      TRY(c->wide ? coarse_hist_gapped<U128>(c, c->part_hist) : coarse_hist_gapped<uint64_t>(c, c->part_hist));
      c->part_hist_step = 1;
    } else {
      KeyArrays ka;
      uint32_t step = 1;
      TRY(c->wide ? coarse_hist<U128>(c, ka, c->part_hist, &step) : coarse_hist<uint64_t>(c, ka, c->part_hist, &step));
      c->part_hist_step = step;
    }
  }
  // contiguous input: scatter all keys once (first call), then every part is a slice of that array
  if (!c->kept_tried && c->cfg.mode == KMC_MODE_CONTIGUOUS && c->cfg.strategy != KMC_STRATEGY_SORT_BASELINE &&
      c->total_bases >= (uint64_t)env_int("KMC_KEPT_MIN_BASES", 1 << 22)) {
    c->kept_tried = true;
    TRY(c->wide ? kept_scatter<U128>(c) : kept_scatter<uint64_t>(c));
  }
  if (c->kept_valid) {
    // part p = top-bits buckets [B(p), B(p+1)), B(p) = the first bucket with at least p/n_parts of the keys before it
    const uint32_t n_l1 = 1u << c->kept_b1;
    if (n_parts > n_l1) return fail(c, KMC_E_ARG, "kmc_finish_part: at most %u parts for this input", n_l1);
    unsigned __int128 all = 0;
    for (uint64_t v : c->kept_count) all += v;
    auto bound = [&](uint32_t p) -> uint32_t {
      if (p == 0) return 0;
      if (p >= n_parts) return n_l1;
      const unsigned __int128 want = (all * p + n_parts - 1) / n_parts;
      unsigned __int128 before = 0;
      uint32_t idx = 0;
      while (idx < n_l1 && before < want) before += c->kept_count[idx++];
      return idx;
    };
    const uint32_t lo = bound(part), hi = bound(part + 1);
    const size_t kw = c->wide ? 16 : 8;
    c->ingested.clear();
    for (uint32_t b = lo; b < hi; b++)
      if (c->kept_count[b]) c->ingested.emplace_back((const unsigned char *)c->kept_keys.p + c->kept_start[b] * kw, c->kept_count[b]);
    int rc;
    if (c->ingested.empty()) { // a range without keys: the empty table
      c->n_total = c->n_distinct = 0; c->finished = true;
      if (n_distinct) *n_distinct = 0;
      if (n_total) *n_total = 0;
      rc = KMC_OK;
    } else {
      // the part's keys share their top bits: plan as for a key range (level-1 bits may go below the coarse prefix —
      // a dense range of canonical k-mers needs them), from the whole input's histogram
      const uint32_t sh = coarse_bits(c) - c->kept_b1;
      c->range_on = true; c->range_lo = lo << sh; c->range_n = (hi - lo) << sh;
      rc = finish_common(c, n_distinct, n_total);
      c->range_on = false;
    }
    c->ingested.clear();
    return rc;
  }
  // part p = coarse bins [B(p), B(p+1)), B(p) = the first bin with at least p/n_parts of the keys before it
  unsigned __int128 total = 0;
  for (uint64_t v : c->part_hist) total += v;
  auto boundary = [&](uint32_t p) -> uint32_t {
    if (p == 0) return 0;
    if (p >= n_parts) return ncoarse;
    const unsigned __int128 want = (total * p + n_parts - 1) / n_parts;
    unsigned __int128 before = 0;
    uint32_t idx = 0;
    while (idx < ncoarse && before < want) before += c->part_hist[idx++];
    return idx;
  };
  const uint32_t lo = boundary(part), hi = boundary(part + 1);
  c->range_on = true; c->range_lo = lo; c->range_n = hi - lo;
  int rc = finish_common(c, n_distinct, n_total);
  c->range_on = false;
  return rc;
}

int kmc_read(kmc_ctx *c, uint64_t first, uint64_t n, uint64_t *key_lo, uint64_t *key_hi, uint64_t *count) {
  if (!c) return KMC_E_ARG;
  if (!c->finished) return fail(c, KMC_E_ARG, "kmc_read before kmc_finish");
  if (first > c->n_distinct || n > c->n_distinct - first) return fail(c, KMC_E_ARG, "row range out of bounds");
  if (!n) return KMC_OK;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  if (key_lo) CK(cudaMemcpy(key_lo, (uint64_t *)c->t_lo.p + first, n * 8, cudaMemcpyDeviceToHost));
  if (key_hi) {
    if (c->wide) CK(cudaMemcpy(key_hi, (uint64_t *)c->t_hi.p + first, n * 8, cudaMemcpyDeviceToHost));
    else memset(key_hi, 0, n * 8);
  }
  if (count) {
    // widen u32 → u64 in place, back to front
    uint32_t *tmp = (uint32_t *)count;
    CK(cudaMemcpy(tmp, (uint32_t *)c->t_cnt.p + first, n * 4, cudaMemcpyDeviceToHost));
    for (uint64_t i = n; i-- > 0;) count[i] = tmp[i];
  }
  return KMC_OK;
}

int kmc_format(kmc_ctx *c, uint64_t first, uint64_t n, int expanded, size_t max_bytes, const char **text, size_t *len) {
  if (!c || !text || !len) return KMC_E_ARG;
  if (!c->finished) return fail(c, KMC_E_ARG, "kmc_format before kmc_finish");
  if (first > c->n_distinct || n > c->n_distinct - first) return fail(c, KMC_E_ARG, "row range out of bounds");
  *text = ""; *len = 0;
  if (!n) return KMC_OK;
  CK(cudaSetDevice(c->device));
  const uint32_t nb = c->key_bases;
  TRY(ensure(c, c->fmt_len, n * 4));
  TRY(ensure(c, c->fmt_off, n * 8));
  const uint32_t *cnt = (const uint32_t *)c->t_cnt.p + first;
  LAUNCH(fmt_len_kernel, grid_for(n, 256), 256, 0, cnt, n, nb, expanded, (uint32_t *)c->fmt_len.p);
  TRY(scan_u32(c, (const uint32_t *)c->fmt_len.p, n, (uint64_t *)c->fmt_off.p));
  uint64_t last_off = 0;
  uint32_t last_len = 0;
  TRY(d2h_small(c, &last_off, (uint64_t *)c->fmt_off.p + (n - 1), 8, 0));
  TRY(d2h_small(c, &last_len, (uint32_t *)c->fmt_len.p + (n - 1), 4, 64));
  const uint64_t bytes = expanded ? (last_off + last_len) * (uint64_t)(nb + 1) : last_off + last_len;
  if (bytes > max_bytes) return fail(c, KMC_E_CAPACITY, "the text of %llu rows is %llu bytes (> %zu)", (unsigned long long)n,
                                     (unsigned long long)bytes, max_bytes);
  TRY(ensure(c, c->fmt_text, bytes + 64));
  if (bytes > c->fmt_host_cap) {
    if (c->fmt_host) CK(cudaFreeHost(c->fmt_host));
    c->fmt_host = nullptr; c->fmt_host_cap = 0;
    size_t cap = std::max<size_t>(bytes + bytes / 8, 1 << 20);
    CK(cudaHostAlloc((void **)&c->fmt_host, cap, cudaHostAllocDefault));
    c->fmt_host_cap = cap;
  }
  LAUNCH(fmt_write_kernel, std::min<uint32_t>(grid_for(n, 8), c->n_sms * 16), 256, 0, (const uint64_t *)c->t_lo.p + first,
         c->wide ? (const uint64_t *)c->t_hi.p + first : (const uint64_t *)nullptr, cnt, (const uint64_t *)c->fmt_off.p, n, nb, expanded,
         (char *)c->fmt_text.p);
  CK(cudaMemcpyAsync(c->fmt_host, c->fmt_text.p, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *text = c->fmt_host; *len = bytes;
  return KMC_OK;
}

int kmc_table_device(kmc_ctx *c, const uint64_t **d_key_lo, const uint64_t **d_key_hi, const uint32_t **d_count) {
  if (!c) return KMC_E_ARG;
  if (!c->finished) return fail(c, KMC_E_ARG, "kmc_table_device before kmc_finish");
  if (d_key_lo) *d_key_lo = c->n_distinct ? (const uint64_t *)c->t_lo.p : nullptr;
  if (d_key_hi) *d_key_hi = (c->wide && c->n_distinct) ? (const uint64_t *)c->t_hi.p : nullptr;
  if (d_count) *d_count = c->n_distinct ? (const uint32_t *)c->t_cnt.p : nullptr;
  return KMC_OK;
}

int kmc_digest(kmc_ctx *c, uint64_t *digest) {
  if (!c || !digest) return KMC_E_ARG;
  if (!c->finished) return fail(c, KMC_E_ARG, "kmc_digest before kmc_finish");
  CK(cudaSetDevice(c->device));
  CK(cudaMemsetAsync(d_digest(c), 0, 8, c->stream));
  if (c->n_distinct)
    LAUNCH(digest_kernel, std::min<uint32_t>(grid_for(c->n_distinct, 256), c->n_sms * 8), 256, 0, (const uint64_t *)c->t_lo.p,
           c->wide ? (const uint64_t *)c->t_hi.p : (const uint64_t *)nullptr, (const uint32_t *)c->t_cnt.p, c->n_distinct,
           d_digest(c));
  unsigned long long h = 0;
  TRY(d2h_small(c, &h, d_digest(c), 8));
  *digest = h;
  return KMC_OK;
}

// Multi-GPU output stage (kmc.h): merge ascending runs of rows with disjoint keys into this ctx's table.
int kmc_merge_tables(kmc_ctx *c, uint32_t n_runs, const uint64_t *const *d_key_lo, const uint64_t *const *d_key_hi,
                     const uint32_t *const *d_count, const uint64_t *n_rows, uint64_t *n_distinct, uint64_t *n_total) {
  if (!c || !n_rows || (n_runs && (!d_key_lo || !d_count))) return KMC_E_ARG;
  if (c->wide && n_runs && !d_key_hi) return fail(c, KMC_E_ARG, "kmc_merge_tables: 128-bit keys need d_key_hi");
  if (c->n_segs || !c->ingested.empty() || !c->ingested_pairs.empty())
    return fail(c, KMC_E_ARG, "kmc_merge_tables: the ctx holds input (call kmc_reset first)");
  CK(cudaSetDevice(c->device));
  const bool wide = c->wide;
  std::vector<RunView> cur;
  uint64_t total = 0;
  for (uint32_t r = 0; r < n_runs; r++) {
    if (!n_rows[r]) continue;
    if (!d_key_lo[r] || !d_count[r] || (wide && !d_key_hi[r])) return fail(c, KMC_E_ARG, "kmc_merge_tables: run %u has null columns", r);
    cur.push_back(RunView{d_key_lo[r], wide ? d_key_hi[r] : nullptr, d_count[r], n_rows[r]});
    total += n_rows[r];
  }
  c->phases.clear(); c->klaunches.clear(); c->events_used = 0; c->launches = 0;
  TRY(zero_scalars(c));
  uint32_t rounds = 0;
  for (size_t m = cur.size(); m > 1; m = (m + 1) / 2) rounds++;
  TRY(ensure(c, c->t_lo, total * 8 + 64));
  if (wide) TRY(ensure(c, c->t_hi, total * 8 + 64));
  TRY(ensure(c, c->t_cnt, total * 4 + 64));
  if (rounds > 1) {
    TRY(ensure(c, c->merge_lo, total * 8 + 64));
    if (wide) TRY(ensure(c, c->merge_hi, total * 8 + 64));
    TRY(ensure(c, c->merge_cnt, total * 4 + 64));
  }
  PHASE_BEGIN("merge");
  if (cur.size() == 1) { // one run: the table is a copy of it
    LAUNCH(merge_copy_kernel, std::min<uint32_t>(grid_for(total, 256), c->n_sms * 8), 256, 0, cur[0], (uint64_t *)c->t_lo.p,
           (uint64_t *)c->t_hi.p, (uint32_t *)c->t_cnt.p);
  }
  for (uint32_t round = 1; cur.size() > 1; round++) {
    const bool to_table = ((rounds - round) & 1u) == 0; // the last round writes the table; the ones before alternate
    uint64_t *o_lo = (uint64_t *)(to_table ? c->t_lo.p : c->merge_lo.p), *o_hi = (uint64_t *)(to_table ? c->t_hi.p : c->merge_hi.p);
    uint32_t *o_cnt = (uint32_t *)(to_table ? c->t_cnt.p : c->merge_cnt.p);
    std::vector<RunView> next;
    uint64_t off = 0;
    for (size_t i = 0; i < cur.size(); i += 2) {
      const uint64_t n = cur[i].n + (i + 1 < cur.size() ? cur[i + 1].n : 0);
      if (i + 1 < cur.size()) {
        const uint32_t grid = grid_for(n, 256 * kMergeRows);
        if (wide) LAUNCH(merge_pair_kernel<true>, grid, 256, 0, cur[i], cur[i + 1], o_lo + off, o_hi + off, o_cnt + off, d_err(c));
        else LAUNCH(merge_pair_kernel<false>, grid, 256, 0, cur[i], cur[i + 1], o_lo + off, o_hi, o_cnt + off, d_err(c));
      } else {
        LAUNCH(merge_copy_kernel, std::min<uint32_t>(grid_for(n, 256), c->n_sms * 8), 256, 0, cur[i], o_lo + off, wide ? o_hi + off : o_hi, o_cnt + off);
      }
      next.push_back(RunView{o_lo + off, wide ? o_hi + off : nullptr, o_cnt + off, n});
      off += n;
    }
    cur.swap(next);
  }
  if (total) LAUNCH(count_sum_kernel, std::min<uint32_t>(grid_for(total, 1024), c->n_sms * 8), 256, 0, (const uint32_t *)c->t_cnt.p, total, d_total_all(c));
  PHASE_END();
  uint32_t err = 0;
  uint64_t sum = 0;
  TRY(read_scalars(c, nullptr, &err, &sum));
  if (err & kFlagSharedKey) return fail(c, KMC_E_ARG, "kmc_merge_tables: two runs hold the same key (owners must hold disjoint key sets)");
  for (auto &p : c->phases) cudaEventElapsedTime(&p.ms, p.a, p.b);
  c->n_distinct = total; c->n_total = sum;
  c->strategy_used = KMC_STRATEGY_SORT;
  c->finished = true;
  if (n_distinct) *n_distinct = total;
  if (n_total) *n_total = sum;
  return KMC_OK;
}

uint32_t kmc_owner_of(uint64_t key_hi, uint64_t key_lo, uint32_t n_parts) { return owner_of(key_hi, key_lo, n_parts); }

int kmc_route(kmc_ctx *c, uint32_t n_parts, uint64_t *part_begin, uint64_t *part_count, const void **d_keys,
              uint32_t *key_bytes) {
  if (!c || !part_begin || !part_count || !d_keys) return KMC_E_ARG;
  if (n_parts < 1 || n_parts > kRadix) return fail(c, KMC_E_ARG, "n_parts must be 1..%d", kRadix);
  if (c->finished) return fail(c, KMC_E_ARG, "kmc_route after kmc_finish");
  dist_abandon(c);
  CK(cudaSetDevice(c->device));
  TRY(zero_scalars(c));
  bool done = false;
  if (c->cfg.mode == KMC_MODE_CONTIGUOUS) {
    if (c->wide) TRY(route_fast<U128>(c, n_parts, part_begin, part_count, &done));
    else TRY(route_fast<uint64_t>(c, n_parts, part_begin, part_count, &done));
  }
  if (!done) {
    std::vector<uint64_t> off(n_parts + 1);
    int rc = c->wide ? route_impl<U128>(c, n_parts, off.data()) : route_impl<uint64_t>(c, n_parts, off.data());
    if (rc) return rc;
    for (uint32_t p = 0; p < n_parts; p++) { part_begin[p] = off[p]; part_count[p] = off[p + 1] - off[p]; }
  }
  *d_keys = c->route_keys.p;
  if (key_bytes) *key_bytes = c->wide ? 16 : 8;
  return KMC_OK;
}

int kmc_route_to_peers(kmc_ctx *c, uint32_t n_parts, void *const *d_part_ptr, uint64_t part_cap_keys, uint64_t *part_count) {
  if (!c || !d_part_ptr || !part_count) return KMC_E_ARG;
  if (n_parts < 1 || n_parts > kRadix) return fail(c, KMC_E_ARG, "n_parts must be 1..%d", kRadix);
  if (c->cfg.mode != KMC_MODE_CONTIGUOUS)
    return fail(c, KMC_E_ARG, "kmc_route_to_peers handles contiguous mode; use kmc_route + an all-to-all for lr-gapped keys");
  if (c->finished) return fail(c, KMC_E_ARG, "kmc_route_to_peers after kmc_finish");
  dist_abandon(c);
  for (uint32_t p = 0; p < n_parts; p++)
    if (!d_part_ptr[p] || ((uintptr_t)d_part_ptr[p] & 127)) return fail(c, KMC_E_ARG, "part pointers must be 128-byte aligned device pointers");
  CK(cudaSetDevice(c->device));
  TRY(zero_scalars(c));
  bool done = false;
  if (c->wide) TRY(route_fast<U128>(c, n_parts, nullptr, part_count, &done, d_part_ptr, part_cap_keys));
  else TRY(route_fast<uint64_t>(c, n_parts, nullptr, part_count, &done, d_part_ptr, part_cap_keys));
  return done ? KMC_OK : fail(c, KMC_E_ARG, "kmc_route_to_peers: nothing routed");
}

int kmc_route_to_peers_part(kmc_ctx *c, uint32_t n_parts, void *const *d_part_ptr, uint64_t part_cap_keys, uint64_t *part_count,
                            uint32_t chunk, uint32_t n_chunks, uint32_t max_ctas) {
  if (!c || !d_part_ptr || !part_count) return KMC_E_ARG;
  if (n_parts < 1 || n_parts > kRadix) return fail(c, KMC_E_ARG, "n_parts must be 1..%d", kRadix);
  if (n_chunks < 1 || chunk >= n_chunks) return fail(c, KMC_E_ARG, "kmc_route_to_peers_part: need chunk < n_chunks");
  if (c->cfg.mode != KMC_MODE_CONTIGUOUS)
    return fail(c, KMC_E_ARG, "kmc_route_to_peers handles contiguous mode; use kmc_route + an all-to-all for lr-gapped keys");
  if (c->finished) return fail(c, KMC_E_ARG, "kmc_route_to_peers after kmc_finish");
  dist_abandon(c);
  for (uint32_t p = 0; p < n_parts; p++)
    if (!d_part_ptr[p] || ((uintptr_t)d_part_ptr[p] & 127)) return fail(c, KMC_E_ARG, "part pointers must be 128-byte aligned device pointers");
  CK(cudaSetDevice(c->device));
  if (!c->owner_on && chunk == 0) TRY(zero_scalars(c));
  else TRY(ensure(c, c->scalars, 64));
  bool done = false;
  if (c->wide) TRY(route_fast<U128>(c, n_parts, nullptr, part_count, &done, d_part_ptr, part_cap_keys, chunk, n_chunks, max_ctas));
  else TRY(route_fast<uint64_t>(c, n_parts, nullptr, part_count, &done, d_part_ptr, part_cap_keys, chunk, n_chunks, max_ctas));
  return done ? KMC_OK : fail(c, KMC_E_ARG, "kmc_route_to_peers: nothing routed");
}

int kmc_owner_begin(kmc_ctx *c, const uint64_t global_hist[4096], uint32_t n_owners, uint32_t *streaming) {
  if (!c || !global_hist || !streaming || !n_owners) return KMC_E_ARG;
  *streaming = 0;
  if (c->finished) return fail(c, KMC_E_ARG, "kmc_owner_begin after kmc_finish");
  if (c->cfg.mode != KMC_MODE_CONTIGUOUS || !c->ingested.empty() || !c->ingested_pairs.empty())
    return fail(c, KMC_E_ARG, "kmc_owner_begin: contiguous mode, before any kmc_ingest_*");
  if (c->cfg.strategy != KMC_STRATEGY_AUTO && c->cfg.strategy != KMC_STRATEGY_SORT) return KMC_OK; // not the partitioned path's job
  CK(cudaSetDevice(c->device));
  return c->wide ? owner_begin_impl<U128>(c, global_hist, n_owners, streaming) : owner_begin_impl<uint64_t>(c, global_hist, n_owners, streaming);
}

int kmc_owner_feed(kmc_ctx *c, const void *d_keys, uint64_t n_keys) {
  if (!c || (!d_keys && n_keys)) return KMC_E_ARG;
  if (!c->owner_on) return fail(c, KMC_E_ARG, "kmc_owner_feed without a successful kmc_owner_begin");
  CK(cudaSetDevice(c->device));
  return c->wide ? owner_feed_impl<U128>(c, d_keys, n_keys) : owner_feed_impl<uint64_t>(c, d_keys, n_keys);
}

int kmc_dist_hist(kmc_ctx *c, uint64_t hist[4096], uint32_t *low_cardinality) {
  if (!c || !hist) return KMC_E_ARG;
  if (c->finished) return fail(c, KMC_E_ARG, "kmc_dist_hist after kmc_finish");
  if (c->cfg.mode != KMC_MODE_CONTIGUOUS || !c->ingested.empty())
    return fail(c, KMC_E_ARG, "kmc_dist_hist: contiguous mode with submitted reads only");
  CK(cudaSetDevice(c->device));
  return c->wide ? dist_hist_impl<U128>(c, hist, low_cardinality) : dist_hist_impl<uint64_t>(c, hist, low_cardinality);
}

int kmc_dist_plan_chunks(kmc_ctx *c, uint32_t world, uint32_t rank, const uint64_t *all_hist, uint32_t n_chunks, uint64_t *need_bytes) {
  if (!c || !all_hist || !need_bytes) return KMC_E_ARG;
  if (world < 1 || world > kDistMaxWorld || rank >= world) return fail(c, KMC_E_ARG, "kmc_dist_plan: 1 <= world <= %u, rank < world", kDistMaxWorld);
  if (n_chunks < 1 || n_chunks > kDistMaxChunks) return fail(c, KMC_E_ARG, "kmc_dist_plan: 1 <= n_chunks <= %u", kDistMaxChunks);
  if (c->cfg.mode != KMC_MODE_CONTIGUOUS) return fail(c, KMC_E_ARG, "kmc_dist_plan: contiguous mode only");
  if (c->owner_stream) CK(cudaStreamSynchronize(c->owner_stream));
  return c->wide ? dist_plan_impl<U128>(c, world, rank, all_hist, n_chunks, need_bytes) : dist_plan_impl<uint64_t>(c, world, rank, all_hist, n_chunks, need_bytes);
}
int kmc_dist_plan(kmc_ctx *c, uint32_t world, uint32_t rank, const uint64_t *all_hist, uint64_t *need_bytes) {
  return kmc_dist_plan_chunks(c, world, rank, all_hist, 1, need_bytes);
}

int kmc_dist_scatter_part(kmc_ctx *c, void *const *d_peer_buf, uint32_t chunk) {
  if (!c || !d_peer_buf) return KMC_E_ARG;
  if (c->finished) return fail(c, KMC_E_ARG, "kmc_dist_scatter after kmc_finish");
  if (!c->dist.valid) return fail(c, KMC_E_ARG, "kmc_dist_scatter: no plan (kmc_dist_plan returned need_bytes = 0?)");
  CK(cudaSetDevice(c->device));
  return c->wide ? dist_scatter_part_impl<U128>(c, d_peer_buf, chunk) : dist_scatter_part_impl<uint64_t>(c, d_peer_buf, chunk);
}
int kmc_dist_scatter_wait(kmc_ctx *c, uint32_t chunk) {
  if (!c) return KMC_E_ARG;
  if (!c->dist.valid || chunk >= c->dist.chunks_sent) return fail(c, KMC_E_ARG, "kmc_dist_scatter_wait: chunk %u has not been scattered", chunk);
  CK(cudaSetDevice(c->device));
  CK(cudaEventSynchronize(c->dist.ev_copied[chunk]));
  return KMC_OK;
}
int kmc_dist_owner_part(kmc_ctx *c, uint32_t chunk) {
  if (!c) return KMC_E_ARG;
  if (!c->dist.valid) return fail(c, KMC_E_ARG, "kmc_dist_owner_part: no plan");
  CK(cudaSetDevice(c->device));
  return c->wide ? dist_owner_part_impl<U128>(c, chunk) : dist_owner_part_impl<uint64_t>(c, chunk);
}
int kmc_dist_scatter_end(kmc_ctx *c, uint32_t *overflow) {
  if (!c || !overflow) return KMC_E_ARG;
  if (!c->dist.valid) return fail(c, KMC_E_ARG, "kmc_dist_scatter_end: no plan");
  CK(cudaSetDevice(c->device));
  return dist_scatter_end_impl(c, overflow);
}
int kmc_dist_scatter(kmc_ctx *c, void *const *d_peer_buf, uint32_t *overflow) {
  if (!c || !d_peer_buf || !overflow) return KMC_E_ARG;
  if (!c->dist.valid) return fail(c, KMC_E_ARG, "kmc_dist_scatter: no plan (kmc_dist_plan returned need_bytes = 0?)");
  for (uint32_t ch = 0; ch < c->dist.n_chunks; ch++) TRY(kmc_dist_scatter_part(c, d_peer_buf, ch));
  return kmc_dist_scatter_end(c, overflow);
}

int kmc_recv_buffer(kmc_ctx *c, uint64_t n_keys, void **d_ptr) {
  if (!c || !d_ptr) return KMC_E_ARG;
  CK(cudaSetDevice(c->device));
  TRY(ensure(c, c->recv_keys, (n_keys + 16) * (c->wide ? 16 : 8)));
  *d_ptr = c->recv_keys.p;
  return KMC_OK;
}

int kmc_ipc_export(kmc_ctx *c, const void *d_ptr, unsigned char handle[64]) {
  if (!c || !d_ptr || !handle) return KMC_E_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CK(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
  memcpy(handle, &h, 64);
  return KMC_OK;
}

int kmc_ipc_open(kmc_ctx *c, const unsigned char handle[64], void **d_peer_ptr) {
  if (!c || !handle || !d_peer_ptr) return KMC_E_ARG;
  CK(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CK(cudaIpcOpenMemHandle(d_peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return KMC_OK;
}

int kmc_ipc_close(kmc_ctx *c, void *d_peer_ptr) {
  if (!c || !d_peer_ptr) return KMC_E_ARG;
  CK(cudaSetDevice(c->device));
  CK(cudaIpcCloseMemHandle(d_peer_ptr));
  return KMC_OK;
}

int kmc_gen_bases(kmc_ctx *c, uint64_t seed, uint64_t first, uint64_t n, uint8_t *d_out) {
  if (!c || (!d_out && n)) return KMC_E_ARG;
  if (!n) return KMC_OK;
  CK(cudaSetDevice(c->device));
  LAUNCH(gen_bases_kernel, (uint32_t)std::min<uint64_t>(grid_for(n / 16 + 2, 256), (uint64_t)c->n_sms * 16), 256, 0, seed, first, n, d_out);
  return KMC_OK;
}

int kmc_gen_nruns(kmc_ctx *c, uint64_t seed, uint64_t first, uint64_t n, uint8_t *d_bases) {
  if (!c || (!d_bases && n)) return KMC_E_ARG;
  if (!n) return KMC_OK;
  CK(cudaSetDevice(c->device));
  LAUNCH(gen_nruns_kernel, (uint32_t)std::min<uint64_t>(grid_for(n / kGenNBlock + 2, 256), (uint64_t)c->n_sms * 16), 256, 0, seed, first, n, d_bases);
  return KMC_OK;
}

int kmc_gen_reads(kmc_ctx *c, uint64_t seed, const uint8_t *d_genome, uint64_t genome_len, uint32_t read_len, uint64_t first_read,
                  uint64_t n_reads, uint8_t *d_out) {
  if (!c || !d_genome || (!d_out && n_reads)) return KMC_E_ARG;
  if (read_len < 1 || genome_len < read_len) return fail(c, KMC_E_ARG, "kmc_gen_reads: need 1 <= read_len <= genome_len");
  if (!n_reads) return KMC_OK;
  CK(cudaSetDevice(c->device));
  LAUNCH(gen_reads_kernel, (uint32_t)std::min<uint64_t>(grid_for(n_reads * read_len, 1024), (uint64_t)c->n_sms * 16), 256, 0, seed, d_genome,
         genome_len, read_len, first_read, n_reads, d_out);
  return KMC_OK;
}

size_t kmc_stats_json(kmc_ctx *c, char *buf, size_t cap) {
  if (!c) return 0;
  build_stats(c);
  size_t need = c->stats.size() + 1;
  if (buf && cap) {
    size_t m = std::min(cap - 1, c->stats.size());
    memcpy(buf, c->stats.data(), m);
    buf[m] = 0;
  }
  return need;
}

} // extern "C"
