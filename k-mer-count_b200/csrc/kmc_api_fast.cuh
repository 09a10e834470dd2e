// kmc_api_fast.cuh — a section of kmc_api.cu (included there, inside its anonymous namespace, after the ctx and its
// helpers; not a stand-alone header): the partitioned counting path — sampled histogram, plan, level-1 / level-2 scatter, bucket sort — and the kept key array of kmc_finish_part.

// Sampled histogram of the top coarse_bits() key bits of the job's keys (raw counts; *step_out = sampling step).
// Uses the head of c->fast_state.
template <typename KeyT>
int coarse_hist(kmc_ctx *c, const KeyArrays &ka, std::vector<uint64_t> &hist, uint32_t *step_out) {
  const uint32_t kb = c->key_bits, cb = coarse_bits(c), ncoarse = 1u << cb;
  TRY(ensure(c, c->fast_state, 4096 * 8 + 64));
  CK(cudaMemsetAsync(c->fast_state.p, 0, 4096 * 8, c->stream));
  unsigned long long *ghist = (unsigned long long *)c->fast_state.p;
  // sample so that ~64M keys are looked at (all of them for small inputs)
  const uint64_t n_in = ka.from_array ? ka.n : c->total_bases;
  uint32_t step = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(kHistSampleMax, n_in >> 26));
  // chunks still on their way are sampled straight from pinned host memory: read 4x less of it over the bus
  if (!ka.from_array) for (size_t i = 0; i < c->n_segs; i++) if (c->segs[i].wait_ready && c->segs[i].host_alias && step > 1) { step = std::min<uint32_t>(64, step * 4); break; }
  PHASE_BEGIN("fast_hist");
  if (ka.from_array) {
    for (auto &a : ka.arrays) {
      uint32_t grid = (uint32_t)std::min<uint64_t>(grid_for(a.second, 1024 * step), (uint64_t)c->n_sms * 8);
      auto fast_hist_array = fast_hist_array_kernel<KeyT>;
      LAUNCH(fast_hist_array, grid, 256, ncoarse * 4, (const KeyT *)a.first, a.second, step, kb - cb, ncoarse, ghist);
    }
  } else {
    for (size_t i = 0; i < c->n_segs; i++) {
      Segment &s = c->segs[i];
      if (!s.n_bases) continue;
      const bool sample_host = step > 1 && s.host_alias && s.wait_ready;
      if (!sample_host) TRY(seg_wait(c, s));
      ExtractParams P = seg_params(c, s, sample_host);
      uint64_t tiles = num_warp_tiles(s.n_bases, win_lanes<KeyT>());
      uint32_t grid = (uint32_t)std::min<uint64_t>((tiles / step + 8) / 8, (uint64_t)c->n_sms * 8);
      auto fast_hist = fast_hist_kernel<KeyT, true>;
      LAUNCH(fast_hist, grid, 256, ncoarse * 4, P, tiles, step, kb - cb, ncoarse, ghist);
    }
  }
  PHASE_END();
  hist.assign(ncoarse, 0);
  TRY(d2h_small(c, hist.data(), ghist, ncoarse * 8));
  *step_out = step;
  return KMC_OK;
}

// lr-gapped keys are never materialised as a whole in a partial count: exact histogram straight from the L/R-mers
template <typename KeyT>
int coarse_hist_gapped(kmc_ctx *c, std::vector<uint64_t> &hist) {
  const uint32_t kb = c->key_bits, cb = coarse_bits(c), ncoarse = 1u << cb;
  TRY(ensure(c, c->fast_state, 4096 * 8 + 64));
  CK(cudaMemsetAsync(c->fast_state.p, 0, 4096 * 8, c->stream));
  uint64_t mx = 1;
  for (size_t i = 0; i < c->n_segs; i++) mx = std::max<uint64_t>(mx, c->segs[i].n_bases);
  TRY(ensure(c, c->gap_l, mx * 8));
  TRY(ensure(c, c->gap_r, mx * 8));
  TRY(ensure(c, c->gap_f, mx));
  PHASE_BEGIN("fast_hist");
  for (size_t i = 0; i < c->n_segs; i++) {
    Segment &s = c->segs[i];
    if (!s.n_bases) continue;
    TRY(seg_wait(c, s));
    GapParams P = gap_params(c, s);
    uint32_t g = grid_for(s.n_bases, 256);
    LAUNCH(gap_mers_kernel, g, 256, 0, P, (uint64_t *)c->gap_l.p, (uint64_t *)c->gap_r.p, (uint8_t *)c->gap_f.p);
    auto gap_hist = gap_hist_kernel<KeyT>;
    LAUNCH(gap_hist, g, 256, ncoarse * 4, P, (const uint64_t *)c->gap_l.p, (const uint64_t *)c->gap_r.p, kb - cb, ncoarse,
           (unsigned long long *)c->fast_state.p);
  }
  PHASE_END();
  hist.assign(ncoarse, 0);
  TRY(d2h_small(c, hist.data(), c->fast_state.p, ncoarse * 8));
  return KMC_OK;
}

// Shape of the two-level partition for an (upper-estimate) coarse histogram: how finely every coarse bin is split
// (2^e[ci] fine buckets of <= target keys), the level-1 width b1, and per level-1 bucket the number of key bits
// (below the b1 prefix) that select its fine bucket.  Level 1 = the top b1 key bits.  Only the level-1 buckets that
// meet the coarse range [c_lo, c_hi) exist, numbered from l1_base (all 2^b1 of them unless this is a partial count —
// which may therefore use more level-1 bits: what is bounded is the number of buckets the scatter kernel ranks in
// shared memory, kMaxL1).  false: the input does not suit the partitioned path.
// Capacity of the fine buckets of a coarse bin whose fine buckets expect `avg` keys each: 10 % + 6 sigma of slack,
// a multiple of kFineAlign (bucket starts are sums of capacities: fast_finish's loads then start on a 128 B line).
inline uint32_t fine_cap_for(double avg, uint32_t cap_max) {
  uint32_t cp = (uint32_t)(avg * 1.10 + 6.0 * std::sqrt(avg) + 64.0);
  return std::min<uint32_t>((cp + (kFineAlign - 1)) & ~(uint32_t)(kFineAlign - 1), cap_max);
}
// development knobs (tools/ab.py): KMC_FINE_TARGET_RT = keys aimed at per 64-bit-key fine bucket, KMC_B1 = level-1 bits
inline int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// 64-bit keys whose fine buckets leave more than 32 key bits are sorted fastest as Split64 (kmc_fast.cuh), which needs
// every bucket to leave at most 32 + kFinishBits bits.  Sparse coarse bins (canonical k-mers thin out towards the top
// of the key space) would be split less than that: the number of extra splits that brings them within reach, or 0
// when the input is too small for it to pay (buckets of a few hundred keys).
inline uint32_t split64_min_e(uint32_t kb, uint64_t n_est, bool wide) {
  const uint32_t cb = std::min<uint32_t>(kCoarseBitsMax, kb);
  if (wide || kb <= cb + 32 + (uint32_t)kFinishBits) return 0;
  const uint32_t me = kb - cb - (32 + (uint32_t)kFinishBits);
  if (me > 12 || (n_est >> (cb + me)) < 512) return 0;
  return me;
}

struct PlanShape {
  uint32_t b1 = 0, l1_base = 0, n_l1 = 0;
  std::vector<uint32_t> e;   // [ncoarse]
  std::vector<uint8_t> l1e;  // [n_l1]
  uint64_t n_fine = 0;
};
// min_e: split every coarse bin at least 2^min_e ways (split64_min_e: keeps sparse bins within Split64's reach)
bool plan_shape(const std::vector<uint64_t> &hist, uint32_t kb, uint32_t c_lo, uint32_t c_hi, bool ranged, int target,
                PlanShape &P, uint32_t b1_min = 0, uint32_t min_e = 0) {
  const uint32_t cb = std::min<uint32_t>(kCoarseBitsMax, kb), ncoarse = 1u << cb;
  P.e.assign(ncoarse, 0);
  for (uint32_t ci = 0; ci < ncoarse; ci++) {
    uint32_t ee = hist[ci] ? min_e : 0;
    while (((hist[ci] + ((1ull << ee) - 1)) >> ee) > (uint64_t)target) ee++;
    if (ee > kb - cb) return false; // cannot split far enough: too many keys share a prefix (duplicates)
    P.e[ci] = ee;
  }
  uint32_t b1_lo = cb > 6 ? cb - 6 : 0, b1_hi = ranged ? cb : std::min<uint32_t>(cb, 10);
  uint64_t nf_guess = 0;
  for (uint32_t ci = c_lo; ci < c_hi; ci++) nf_guess += 1ull << P.e[ci];
  uint32_t b1 = (uint32_t)std::lround(std::log2(std::sqrt((double)nf_guess) * (double)ncoarse / (double)(c_hi - c_lo)));
  b1 = std::max(std::max(b1_lo, std::min(b1_min, b1_hi)), std::min(b1, b1_hi));
  if (const int forced = env_int("KMC_B1", 0); forced > 0) b1 = std::max(b1_lo, std::min<uint32_t>((uint32_t)forced, b1_hi));
  auto l1_span = [&](uint32_t bits, uint32_t *base) { // level-1 buckets met by the coarse range at `bits` level-1 bits
    *base = c_lo >> (cb - bits);
    return ((c_hi - 1) >> (cb - bits)) + 1 - *base;
  };
  uint32_t l1_base = 0, n_l1 = 0;
  while (b1 > b1_lo && l1_span(b1, &l1_base) > (uint32_t)kMaxL1) b1--;
  for (;; b1++) {
    if (b1 > b1_hi) return false;
    n_l1 = l1_span(b1, &l1_base);
    if (n_l1 > (uint32_t)kMaxL1) return false; // too many keys for two levels of this size
    P.l1e.assign(n_l1, 0);
    uint32_t mx = 0;
    for (uint32_t rb = 0; rb < n_l1; rb++) {
      const uint32_t b = l1_base + rb;
      uint32_t em = 0;
      for (uint32_t ci = b << (cb - b1); ci < ((b + 1) << (cb - b1)); ci++) em = std::max(em, P.e[ci]);
      P.l1e[rb] = (uint8_t)(cb - b1 + em);
      mx = std::max<uint32_t>(mx, P.l1e[rb]);
    }
    if ((1ull << mx) <= (uint64_t)kMaxFinePerL1) break;
  }
  P.b1 = b1; P.l1_base = l1_base; P.n_l1 = n_l1;
  P.n_fine = 0;
  for (uint32_t b = 0; b < n_l1; b++) P.n_fine += 1ull << P.l1e[b];
  return P.n_fine <= (1ull << 28);
}

// level-2 scatter of the keys in [l1_done, l1_cursor) of every level-1 bucket (all of them if done == nullptr);
// t_max = tiles per bucket the grid provides (a CTA takes several if there are more); nb_max = most fine buckets
// under one level-1 bucket (sizes the shared memory).
template <typename KeyT, typename L2T>
int launch_part2_as(kmc_ctx *c, const FastPlan &pl, const KeyT *l1, uint32_t nb_max, uint64_t t_max, unsigned long long *done, bool flush) {
  const size_t smem = PartSmem<KeyT>::bytes(p2_tile<KeyT>(), nb_max);
  auto fast_part2 = fast_part2_kernel<KeyT, L2T>;
  CK(cudaFuncSetAttribute(fast_part2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 grid((uint32_t)std::min<uint64_t>(std::max<uint64_t>(t_max, 1), 1u << 20), pl.n_l1);
  LAUNCH(fast_part2, grid, kFastThreads, smem, pl, l1, (L2T *)c->fast_l2.p, d_err(c), (const unsigned long long *)done, flush ? 1u : 0u, nb_max);
  if (done) {
    LAUNCH(l2_done_kernel, grid_for(pl.n_l1, 256), 256, 0, pl, done, (uint32_t)p2_tile<KeyT>(), flush ? 1u : 0u, grid.x);
    c->launches--; // plumbing
  }
  return KMC_OK;
}
template <typename KeyT>
int launch_part2(kmc_ctx *c, const FastPlan &pl, const KeyT *l1, bool key32, uint32_t nb_max, uint64_t t_max, unsigned long long *done = nullptr,
                 bool flush = true) {
  if constexpr (sizeof(KeyT) == 16) return launch_part2_as<U128, U128>(c, pl, l1, nb_max, t_max, done, flush);
  else if (key32) return launch_part2_as<uint64_t, uint32_t>(c, pl, l1, nb_max, t_max, done, flush);
  else return launch_part2_as<uint64_t, uint64_t>(c, pl, l1, nb_max, t_max, done, flush);
}

constexpr uint64_t kFastMinKeys = 1u << 18, kFastMinKeysGapped = 1u << 23;

struct FastJob {   // one partitioned count in progress: the plan, and what the later stages need of it
  bool active = false;
  FastPlan pl{};
  bool key32 = false, split64 = false, ranged = false;
  uint32_t nb_max = 1, c_lo = 0, c_hi = 0;
  uint64_t t_max = 1, n_fine = 0, fed = 0;
  unsigned int *ticket = nullptr;
  unsigned long long *d_total = nullptr, *status = nullptr, *l1_done = nullptr;
};
FastJob &job_of(kmc_ctx *c) {
  if (!c->job_box) c->job_box = new FastJob();
  return *static_cast<FastJob *>(c->job_box);
}

// ---- partitioned fast path (kmc_fast.cuh) --------------------------------------------------------------------
// Three stages, so that keys can be fed while they arrive (chunks of a pinned submit; chunks routed by the other ranks):
//   fast_begin  plan from an upper-estimate coarse histogram, buffers, tables            → *ok
//   fast_feed_* level-1 scatter of a segment / a key array (+ the level-2 scatter of what has come in so far)
//   fast_end    (rest of the) level-2 scatter, bucket sort, totals                      → *used
// *ok / *used = false: the input does not suit the path (tiny, duplicate-heavy, or a bucket overflowed); nothing is left
// behind and the caller counts with the baseline path.
// relax = 1: aim at half-full buckets (retry after an overflow: input whose keys come in many copies spreads
// less evenly than the plan's Poisson slack assumes).
template <typename KeyT>
int fast_begin(kmc_ctx *c, std::vector<uint64_t> &hist, uint64_t n_est, int relax, bool *ok) {
  constexpr bool kWide = sizeof(KeyT) == 16;
  int kCap = kWide ? 4096 : kFineCap;
  int kTarget = (kWide ? 3200 : std::min(env_int("KMC_FINE_TARGET_RT", kFineTarget), kFineTarget)) >> relax;
  *ok = false;
  FastJob &J = job_of(c);
  J.active = false;
  const uint32_t kb = c->key_bits;
  const uint32_t cb = std::min<uint32_t>(kCoarseBitsMax, kb);
  const uint32_t ncoarse = 1u << cb;
  const bool ranged = c->range_on;
  const uint32_t c_lo = ranged ? c->range_lo : 0u, c_hi = ranged ? c->range_lo + c->range_n : ncoarse;
  // fast_state layout: ghist[4096] u64 | ticket u32 (+pad) | d_total u64 | l1_cursor[kMaxL1] u64 | l1_done[kMaxL1] u64 | fine_cursor[nf] u32 | status[nf] u64
  const size_t off_ticket = 4096 * 8, off_dtotal = off_ticket + 8, off_l1cur = off_dtotal + 8, off_l1done = off_l1cur + kMaxL1 * 8,
               off_fine = off_l1done + kMaxL1 * 8;
  // ---- plan
  PlanShape shape;
  const uint32_t min_e = getenv("KMC_NO_SPLIT64") ? 0u : split64_min_e(kb, n_est, kWide);
  if (!plan_shape(hist, kb, c_lo, c_hi, ranged, kTarget, shape, 0, min_e)) return KMC_OK;
  if (!kWide && (kFineCap64 != kFineCap || kFineTarget64 != kFineTarget)) {
    // buckets that leave more than 32 key bits are sorted as 64-bit elements, whose bucket shape is smaller: plan again
    bool wide_elems = false;
    for (uint32_t b = 0; b < shape.n_l1; b++) if (kb - shape.b1 - shape.l1e[b] > 32) wide_elems = true;
    if (wide_elems) {
      kCap = kFineCap64;
      kTarget = std::min(kTarget, kFineTarget64 >> relax);
      if (!plan_shape(hist, kb, c_lo, c_hi, ranged, kTarget, shape, 0, min_e)) return KMC_OK;
    }
  }
  const uint32_t b1 = shape.b1, l1_base = shape.l1_base, n_l1 = shape.n_l1;
  const std::vector<uint8_t> &l1e = shape.l1e;
  const uint64_t n_fine = shape.n_fine;
  // Host tables (a few tens of KB): per level-1 bucket l1_start | l1_cap | l1_tile0 | l1_fine0 | l1_e, and per coarse
  // bin the first fine bucket, its level-2 start and the capacity of its fine buckets.  The per-fine-bucket
  // descriptors (n_fine x 32 B, megabytes) are expanded from these on the device (plan_expand_kernel): filling and
  // uploading them from the host cost ~1.2 ms of idle GPU per job.
  auto al16 = [](size_t x) { return (x + 15) & ~size_t(15); };
  const uint32_t cshift = cb - b1, n_cb = n_l1 << cshift; // coarse bins under the existing level-1 buckets
  const size_t o_l1s = 0, o_cap = o_l1s + al16((size_t)(n_l1 + 1) * 8),
               o_t0 = o_cap + al16((size_t)n_l1 * 8), o_f0 = o_t0 + al16((size_t)(n_l1 + 1) * 4),
               o_e = o_f0 + al16((size_t)(n_l1 + 1) * 4), o_cs = o_e + al16(n_l1), o_cf = o_cs + al16((size_t)n_cb * 8),
               o_cc = o_cf + al16((size_t)n_cb * 4), tab_bytes = o_cc + al16((size_t)n_cb * 2);
  c->fast_host.assign(tab_bytes, 0);
  uint64_t *cstart = (uint64_t *)(c->fast_host.data() + o_cs);
  uint32_t *cfine0 = (uint32_t *)(c->fast_host.data() + o_cf);
  uint16_t *ccap = (uint16_t *)(c->fast_host.data() + o_cc);
  uint64_t *l1s = (uint64_t *)(c->fast_host.data() + o_l1s), *l1cap = (uint64_t *)(c->fast_host.data() + o_cap);
  uint32_t *t0 = (uint32_t *)(c->fast_host.data() + o_t0), *f0 = (uint32_t *)(c->fast_host.data() + o_f0);
  uint8_t *l1ep = c->fast_host.data() + o_e;
  uint64_t l1_keys = 0, l2_keys = 0, tiles2 = 0, t_max = 1;
  uint32_t fb = 0, nb_max = 1;
  bool key32 = !kWide; // every bucket leaves <= 32 key bits: level 2 stores 32-bit suffixes
  bool split64 = !kWide && !getenv("KMC_NO_SPLIT64"); // ... at most 32 + kFinishBits: sorted as 32-bit suffixes + sub-bin ids (Split64)
  for (uint32_t b = 0; b < n_l1; b++) {
    if (kb - b1 - l1e[b] > 32) key32 = false;
    if (kb - b1 - l1e[b] > 32 + (uint32_t)kFinishBits) split64 = false;
  }
  if (key32) split64 = false;
  for (uint32_t b = 0; b < n_l1; b++) { // b: index among the existing level-1 buckets; b_abs: its key prefix
    const uint32_t b_abs = l1_base + b;
    uint64_t nb = 0;
    const uint32_t sub_bits = l1e[b] - (cb - b1); // fine buckets per coarse bin of this level-1 bucket = 2^sub_bits
    f0[b] = fb;
    for (uint32_t ci = b_abs << (cb - b1); ci < ((b_abs + 1) << (cb - b1)); ci++) {
      nb += hist[ci];
      double avg = (double)hist[ci] / (double)(1ull << sub_bits);
      const uint32_t cp = relax ? (uint32_t)kCap : fine_cap_for(avg, kCap);
      const uint32_t ci_rel = ci - (l1_base << cshift);
      cstart[ci_rel] = l2_keys; cfine0[ci_rel] = fb; ccap[ci_rel] = (uint16_t)cp;
      l2_keys += (uint64_t)cp << sub_bits;
      fb += 1u << sub_bits;
    }
    uint64_t cap1 = ((uint64_t)((double)nb * 1.03) + 8192 + 15) & ~15ull;
    l1s[b] = l1_keys; l1cap[b] = cap1; t0[b] = (uint32_t)tiles2; l1ep[b] = l1e[b];
    l1_keys += cap1;
    tiles2 += (cap1 + p2_tile<KeyT>() - 1) / p2_tile<KeyT>();
    t_max = std::max<uint64_t>(t_max, (cap1 + p2_tile<KeyT>() - 1) / p2_tile<KeyT>());
    nb_max = std::max<uint32_t>(nb_max, 1u << l1e[b]);
  }
  l1s[n_l1] = l1_keys; t0[n_l1] = (uint32_t)tiles2; f0[n_l1] = fb;
  if (tiles2 > 0x7FFFFFFFull) return KMC_OK;
  HOST_MARK("planned");

  // ---- buffers (each array ends with a trash area of one tile + slack for runs that spill over a bucket end)
  const uint64_t slack = 2 * kMaxTile;
  TRY(ensure(c, c->fast_tables, tab_bytes));
  TRY(ensure(c, c->fast_fdesc, n_fine * sizeof(FineDesc)));
  TRY(ensure(c, c->fast_l1, (l1_keys + slack) * sizeof(KeyT)));
  // the level-1 array and the table's key column trade places after every job (64-bit keys): size both, or the
  // smaller one would be freed and reallocated on alternate jobs
  if (!kWide) TRY(ensure(c, c->t_lo, (l1_keys + slack) * sizeof(KeyT)));
  TRY(ensure(c, c->fast_l2, (l2_keys + 2 * slack) * sizeof(KeyT)));
  if (kWide) { TRY(ensure(c, c->t_lo, l1_keys * 8)); TRY(ensure(c, c->t_hi, l1_keys * 8)); }
  TRY(ensure(c, c->t_cnt, l1_keys * 4));
  const size_t off_status = (off_fine + n_fine * 4 + 15) & ~size_t(15);
  if (off_status + n_fine * 8 + 64 > c->fast_state.cap) {
    TRY(ensure(c, c->fast_state, off_status + n_fine * 8 + 64));
  }
  HOST_MARK("buffers");
  CK(cudaMemsetAsync(c->fast_state.p, 0, off_status + n_fine * 8, c->stream));
  TRY(h2d_small(c, c->fast_tables.p, c->fast_host.data(), tab_bytes));
  HOST_MARK("uploaded");
  unsigned char *st = (unsigned char *)c->fast_state.p, *tb = (unsigned char *)c->fast_tables.p;
  FastPlan &pl = J.pl;
  pl = FastPlan{};
  pl.kb = kb; pl.b1 = b1; pl.n_l1 = n_l1; pl.n_fine = (uint32_t)n_fine; pl.l1_base = l1_base;
  pl.l1_trash = l1_keys; pl.l2_trash = l2_keys + slack;
  pl.fdesc = (const FineDesc *)c->fast_fdesc.p;
  pl.l1_start = (const uint64_t *)(tb + o_l1s); pl.l1_cap = (const uint64_t *)(tb + o_cap);
  pl.l1_tile0 = (const uint32_t *)(tb + o_t0); pl.l1_fine0 = (const uint32_t *)(tb + o_f0); pl.l1_e = (const uint8_t *)(tb + o_e);
  pl.l1_cursor = (unsigned long long *)(st + off_l1cur); pl.fine_cursor = (uint32_t *)(st + off_fine);
  J.ticket = (unsigned int *)(st + off_ticket);
  J.d_total = (unsigned long long *)(st + off_dtotal);
  J.status = (unsigned long long *)(st + off_status);
  J.l1_done = (unsigned long long *)(st + off_l1done);
  LAUNCH(plan_expand_kernel, n_cb, 128, 0, (FineDesc *)c->fast_fdesc.p, (const uint64_t *)(tb + o_cs), (const uint32_t *)(tb + o_cf),
         (const uint16_t *)(tb + o_cc), pl.l1_fine0, pl.l1_e, cshift, l1_base, kb, b1, (uint32_t)kWide);
  c->launches--; // plumbing


  J.key32 = key32; J.split64 = split64; J.ranged = ranged; J.nb_max = nb_max; J.c_lo = c_lo; J.c_hi = c_hi;
  J.t_max = t_max; J.n_fine = n_fine; J.fed = 0;
  J.active = true;
  c->fast_variant = kWide ? "u128" : key32 ? "u32" : split64 ? "split64" : "u64";
  *ok = true;
  return KMC_OK;
}

// level-1 scatter of a key array; incremental: follow it with the level-2 scatter of the whole tiles that have come in
template <typename KeyT>
int fast_feed_array(kmc_ctx *c, const void *keys, uint64_t n, bool incremental) {
  FastJob &J = job_of(c);
  if (!n) return KMC_OK;
  const size_t smem = L1Smem<KeyT>::bytes(arr_tile<KeyT>(), J.pl.n_l1);
  auto fast_part1_array = fast_part1_array_kernel<KeyT>;
  CK(cudaFuncSetAttribute(fast_part1_array, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PHASE_BEGIN("fast_part1");
  const uint32_t grid = (uint32_t)std::min<uint64_t>(grid_for(n, arr_tile<KeyT>()), (uint64_t)c->n_sms);
  LAUNCH(fast_part1_array, grid, kFastThreads, smem, (const KeyT *)keys, n, J.pl, (KeyT *)c->fast_l1.p, d_err(c));
  PHASE_END();
  J.fed += n;
  if (incremental) {
    PHASE_BEGIN("fast_part2");
    const uint64_t t_arr = (uint64_t)((double)n / J.pl.n_l1 / p2_tile<KeyT>() * 1.25) + 2;
    TRY(launch_part2<KeyT>(c, J.pl, (const KeyT *)c->fast_l1.p, J.key32, J.nb_max, std::min(t_arr, J.t_max), J.l1_done, false));
    PHASE_END();
  }
  return KMC_OK;
}

template <typename KeyT>
int fast_end(kmc_ctx *c, bool incremental, bool *used) {
  constexpr bool kWide = sizeof(KeyT) == 16;
  FastJob &J = job_of(c);
  *used = false;
  J.active = false;
  const FastPlan &pl = J.pl;
  const bool key32 = J.key32, split64 = J.split64;
  const uint64_t n_fine = J.n_fine;
  const uint32_t n_l1 = pl.n_l1;
  // ---- level 2 (all of it, or what the incremental rounds have left)
  PHASE_BEGIN("fast_part2");
  TRY(launch_part2<KeyT>(c, pl, (const KeyT *)c->fast_l1.p, key32, J.nb_max, J.t_max, incremental ? J.l1_done : nullptr, true));
  PHASE_END();
  // ---- finish: the level-1 array is dead after part2 and (64-bit keys) becomes the table's key column
  PHASE_BEGIN("fast_finish");
  {
    unsigned long long *prof = nullptr;
    static const bool want_prof = getenv("KMC_FINISH_PROF") && getenv("KMC_FINISH_PROF")[0] == '1';
    if (want_prof) { prof = (unsigned long long *)c->fast_state.p; CK(cudaMemsetAsync(prof, 0, 16 * 8, c->stream)); } // the histogram is dead by now
    uint32_t grid;
    if constexpr (kWide) {
      size_t fsmem = sizeof(FinishSmem<U128>);
      auto fast_finish = fast_finish_kernel<U128>;
      CK(cudaFuncSetAttribute(fast_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
      grid = (uint32_t)std::min<uint64_t>(n_fine, (uint64_t)c->n_sms * 2);
      LAUNCH(fast_finish, grid, kFinThreads, fsmem, pl, (const U128 *)c->fast_l2.p, (uint64_t *)c->t_lo.p, (uint64_t *)c->t_hi.p,
             (uint32_t *)c->t_cnt.p, J.status, J.ticket, d_err(c), J.d_total, prof);
    } else if (key32) {
      size_t fsmem = sizeof(FinishSmem<uint32_t>);
      auto fast_finish = fast_finish_kernel<uint32_t>;
      CK(cudaFuncSetAttribute(fast_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
      grid = (uint32_t)std::min<uint64_t>(n_fine, (uint64_t)c->n_sms * KMC_FINISH_MINB32);
      LAUNCH(fast_finish, grid, kFinThreads, fsmem, pl, (const uint32_t *)c->fast_l2.p, (uint64_t *)c->fast_l1.p, (uint64_t *)nullptr,
             (uint32_t *)c->t_cnt.p, J.status, J.ticket, d_err(c), J.d_total, prof);
    } else if (split64) {
      size_t fsmem = sizeof(FinishSmem<Split64>);
      auto fast_finish = fast_finish_kernel<Split64>;
      CK(cudaFuncSetAttribute(fast_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
      grid = (uint32_t)std::min<uint64_t>(n_fine, (uint64_t)c->n_sms * 2);
      LAUNCH(fast_finish, grid, kFinThreads, fsmem, pl, (const uint64_t *)c->fast_l2.p, (uint64_t *)c->fast_l1.p, (uint64_t *)nullptr,
             (uint32_t *)c->t_cnt.p, J.status, J.ticket, d_err(c), J.d_total, prof);
    } else {
      size_t fsmem = sizeof(FinishSmem<uint64_t>);
      auto fast_finish = fast_finish_kernel<uint64_t>;
      CK(cudaFuncSetAttribute(fast_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
      grid = (uint32_t)std::min<uint64_t>(n_fine, (uint64_t)c->n_sms * KMC_FINISH_MINB64);
      LAUNCH(fast_finish, grid, kFinThreads, fsmem, pl, (const uint64_t *)c->fast_l2.p, (uint64_t *)c->fast_l1.p, (uint64_t *)nullptr,
             (uint32_t *)c->t_cnt.p, J.status, J.ticket, d_err(c), J.d_total, prof);
    }
    if (want_prof) {
      unsigned long long h[16];
      CK(cudaMemcpyAsync(h, prof, sizeof h, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      unsigned long long tot = 0;
      for (int i = 0; i < 10; i++) tot += h[i];
      fprintf(stderr, "[kmc] fast_finish cycles per phase (thread 0, summed over %u CTAs), %% of total:", grid);
      for (int i = 0; i < 11; i++) fprintf(stderr, " p%d=%.1f%%", i, 100.0 * (double)h[i] / (double)std::max<unsigned long long>(1, tot));
      fprintf(stderr, "  total=%.0f cycles/CTA\n", (double)tot / grid);
    }
  }
  PHASE_END();
  uint64_t d = 0;
  uint32_t err = 0;
  std::vector<unsigned long long> cur(n_l1);
  TRY(d2h_small(c, &d, J.d_total, 8));
  TRY(d2h_small(c, cur.data(), pl.l1_cursor, (size_t)n_l1 * 8));
  TRY(read_scalars(c, nullptr, &err));
  if (err & kFlagSpin) return fail(c, KMC_E_CUDA, "fast_finish: look-back did not make progress");
  if (err & kFlagOverflow) {
    c->fast_fallbacks++;
    TRY(zero_scalars(c));
    return KMC_OK; // recount with the data-independent path
  }

  uint64_t N = 0;
  for (unsigned long long v : cur) N += v;
  if (!kWide) std::swap(c->t_lo, c->fast_l1);
  c->n_total = N; c->n_distinct = d;
  c->strategy_used = KMC_STRATEGY_SORT;
  *used = true;
  return KMC_OK;
}

template <typename KeyT>
int finish_fast(kmc_ctx *c, bool *used, int relax = 0) {
  *used = false;
  const uint32_t kb = c->key_bits;
  const uint32_t cb = std::min<uint32_t>(kCoarseBitsMax, kb);
  const uint32_t ncoarse = 1u << cb;
  KeyArrays ka;
  TRY(key_sources<KeyT>(c, &ka));
  TRY(ensure(c, c->fast_state, 4096 * 8 + 64));
  std::vector<uint64_t> hist;
  uint32_t step = 1;
  if (c->range_on && c->part_hist_step) { hist = c->part_hist; step = c->part_hist_step; } // partial count: computed once per input
  else TRY(coarse_hist<KeyT>(c, ka, hist, &step));
  // a partial count sees only the coarse bins of its key range
  const bool ranged = c->range_on;
  const uint32_t c_lo = ranged ? c->range_lo : 0u, c_hi = ranged ? c->range_lo + c->range_n : ncoarse;
  if (ranged) for (uint32_t ci = 0; ci < ncoarse; ci++) if (ci < c_lo || ci >= c_hi) hist[ci] = 0;
  if (c_hi <= c_lo) return KMC_OK; // empty range: the generic path returns the empty table
  // the histogram is a 1-in-step sample: scale it to an upper estimate (+5 sigma of the sampling noise)
  uint64_t n_est = 0;
  for (uint64_t &v : hist) {
    double est = (double)v * step;
    if (step > 1) est += 5.0 * std::sqrt(est * step) + step;
    v = (uint64_t)est;
    n_est += v;
  }
  // small job: the generic path is as fast (a few passes over a few MB) and does not care what the keys look like.
  // lr-gapped keys come in groups of up to d_max - d_min + 1 that share their L-mer, 2 * l_len bits; when the input is
  // repetitive as well (the reference's own fixture: 3.55 M keys, 54-bit prefixes shared by the thousand) no prefix
  // partition can separate them, so such jobs take the generic path up to a larger size.
  if (n_est < (c->cfg.mode == KMC_MODE_LR_GAPPED ? kFastMinKeysGapped : kFastMinKeys)) return KMC_OK;
  HOST_MARK("hist_read");
  bool ok = false;
  TRY(fast_begin<KeyT>(c, hist, n_est, relax, &ok));
  if (!ok) return KMC_OK;
  FastJob &J = job_of(c);
  const FastPlan &pl = J.pl;
  const uint32_t b1 = pl.b1, l1_base = pl.l1_base, n_l1 = pl.n_l1;

  // ---- level 1
  bool incremental = false;
  if (ka.from_array) {
    for (auto &a : ka.arrays) TRY(fast_feed_array<KeyT>(c, a.first, a.second, false));
  } else {
    PHASE_BEGIN("fast_part1");
    size_t smem = L1Smem<KeyT>::bytes(part1_stage<KeyT>(), n_l1);
    auto fast_part1 = fast_part1_kernel<KeyT, true, PrefixBucket>;
    auto fast_part1_ranged = fast_part1_kernel<KeyT, true, PrefixBucketT<true>>;
    if (ranged) CK(cudaFuncSetAttribute(fast_part1_ranged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else CK(cudaFuncSetAttribute(fast_part1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const PrefixBucket bucket = make_prefix_bucket<false>(kb, b1);
    const PrefixBucketT<true> bucket_ranged = make_prefix_bucket<true>(kb, b1, l1_base, kb - cb, c_lo, c_hi - c_lo);
    // A large pinned submit arrives in chunks (submit_chunked): the level-2 scatter then follows every chunk's level-1
    // scatter for the keys that have come in so far (whole tiles only; fast_end takes the rest), so that when
    // the last chunk has landed only its own share of the two scatters and the bucket sort remain.
    size_t n_live = 0, i_last = 0;
    for (size_t i = 0; i < c->n_segs; i++) if (c->segs[i].n_bases) { n_live++; i_last = i; }
    incremental = n_live >= 4;
    for (size_t i = 0; i < c->n_segs; i++) {
      Segment &s = c->segs[i];
      if (!s.n_bases) continue;
      TRY(seg_wait(c, s));
      ExtractParams P = seg_params(c, s);
      uint64_t tiles = num_warp_tiles(s.n_bases, win_lanes<KeyT>());
      uint32_t grid = (uint32_t)std::min<uint64_t>((tiles + kFastWarps - 1) / kFastWarps, (uint64_t)c->n_sms);
      const uint64_t n_ct = (tiles + kFastWarps - 1) / kFastWarps;
      if (ranged) LAUNCH(fast_part1_ranged, grid, kFastThreads, smem, P, tiles, pl, bucket_ranged, (KeyT *)c->fast_l1.p, d_err(c), (uint64_t)0, n_ct);
      else LAUNCH(fast_part1, grid, kFastThreads, smem, P, tiles, pl, bucket, (KeyT *)c->fast_l1.p, d_err(c), (uint64_t)0, n_ct);
      if (incremental && i != i_last) {
        PHASE_END();
        PHASE_BEGIN("fast_part2");
        const uint64_t t_seg = (uint64_t)((double)s.n_bases / n_l1 / p2_tile<KeyT>() * 1.25) + 2;
        TRY(launch_part2<KeyT>(c, pl, (const KeyT *)c->fast_l1.p, J.key32, J.nb_max, std::min(t_seg, J.t_max), J.l1_done, false));
        PHASE_END();
        PHASE_BEGIN("fast_part1");
      }
    }
    PHASE_END();
  }
  return fast_end<KeyT>(c, incremental, used);
}

// ---- partial counts of one input (kmc_finish_part): all keys scattered ONCE by their top bits --------------------------
// A job too large to be counted in one go (1e10 bases at k=31: level-1 + level-2 arrays + table exceed HBM) is counted
// in key ranges.  Extracting the whole input again for every range cost 47 ms x 8 of 634 ms at that size; instead the
// first call runs the level-1 scatter kernel once over everything, 2^b1 buckets by key prefix (74 GB of keys fit beside
// the 10 GB of bases), and part p is then counted from the buckets of its range through the key-array front end.
template <typename KeyT>
int kept_scatter(kmc_ctx *c) {
  const uint32_t kb = c->key_bits, cb = coarse_bits(c), ncoarse = 1u << cb;
  const uint32_t b1 = std::min<uint32_t>(cb, (uint32_t)env_int("KMC_KEPT_BITS", 8)), n_l1 = 1u << b1, cshift = cb - b1;
  c->kept_valid = false;
  c->kept_start.assign(n_l1 + 1, 0); c->kept_count.assign(n_l1, 0);
  std::vector<uint64_t> cap(n_l1, 0);
  uint64_t total = 0;
  for (uint32_t b = 0; b < n_l1; b++) {
    double nb = 0;
    for (uint32_t ci = b << cshift; ci < ((b + 1) << cshift) && ci < ncoarse; ci++) {
      double est = (double)c->part_hist[ci] * c->part_hist_step;
      if (c->part_hist_step > 1) est += 5.0 * std::sqrt(est * c->part_hist_step) + c->part_hist_step;
      nb += est;
    }
    cap[b] = ((uint64_t)(nb * 1.02) + 8192 + 15) & ~15ull;
    c->kept_start[b] = total;
    total += cap[b];
  }
  c->kept_start[n_l1] = total;
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  const size_t need = (total + 2 * kMaxTile) * sizeof(KeyT);
  // room for the array AND for one part's own buffers afterwards, or the old way (extract per part) is the only way
  if (need > c->kept_keys.cap && need + need / 2 > free_b + c->kept_keys.cap) return KMC_OK;
  TRY(ensure(c, c->kept_keys, need));
  auto al16 = [](size_t x) { return (x + 15) & ~size_t(15); };
  const size_t o_s = 0, o_c = al16((size_t)(n_l1 + 1) * 8), tab_bytes = o_c + al16((size_t)n_l1 * 8);
  std::vector<unsigned char> host(tab_bytes, 0);
  memcpy(host.data() + o_s, c->kept_start.data(), (size_t)(n_l1 + 1) * 8);
  memcpy(host.data() + o_c, cap.data(), (size_t)n_l1 * 8);
  TRY(ensure(c, c->kept_tables, tab_bytes));
  TRY(ensure(c, c->kept_state, (size_t)kMaxL1 * 8 + 64));
  TRY(zero_scalars(c));
  CK(cudaMemsetAsync(c->kept_state.p, 0, (size_t)kMaxL1 * 8, c->stream));
  TRY(h2d_small(c, c->kept_tables.p, host.data(), tab_bytes));
  FastPlan pl{};
  pl.kb = kb; pl.b1 = b1; pl.n_l1 = n_l1; pl.l1_base = 0; pl.l1_trash = total;
  pl.l1_start = (const uint64_t *)((unsigned char *)c->kept_tables.p + o_s);
  pl.l1_cap = (const uint64_t *)((unsigned char *)c->kept_tables.p + o_c);
  pl.l1_cursor = (unsigned long long *)c->kept_state.p;
  PHASE_BEGIN("kept_scatter");
  {
    size_t smem = L1Smem<KeyT>::bytes(part1_stage<KeyT>(), n_l1);
    auto fast_part1 = fast_part1_kernel<KeyT, true, PrefixBucket>;
    CK(cudaFuncSetAttribute(fast_part1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const PrefixBucket bucket = make_prefix_bucket<false>(kb, b1);
    for (size_t i = 0; i < c->n_segs; i++) {
      Segment &s = c->segs[i];
      if (!s.n_bases) continue;
      TRY(seg_wait(c, s));
      ExtractParams P = seg_params(c, s);
      const uint64_t tiles = num_warp_tiles(s.n_bases, win_lanes<KeyT>()), n_ct = (tiles + kFastWarps - 1) / kFastWarps;
      LAUNCH(fast_part1, (uint32_t)std::min<uint64_t>(n_ct, (uint64_t)c->n_sms), kFastThreads, smem, P, tiles, pl, bucket, (KeyT *)c->kept_keys.p,
             d_err(c), (uint64_t)0, n_ct);
    }
  }
  PHASE_END();
  std::vector<unsigned long long> cur(n_l1);
  uint32_t err = 0;
  TRY(d2h_small(c, cur.data(), pl.l1_cursor, (size_t)n_l1 * 8));
  TRY(read_scalars(c, nullptr, &err));
  if (err & kFlagOverflow) { TRY(zero_scalars(c)); return KMC_OK; } // skewed beyond the estimate: the old way
  for (uint32_t b = 0; b < n_l1; b++) c->kept_count[b] = cur[b];
  c->kept_b1 = b1;
  c->kept_valid = true;
  return KMC_OK;
}
