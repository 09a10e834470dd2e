// kmc_api_dist.cuh — a section of kmc_api.cu (included there, inside its anonymous namespace, after the ctx and its
// helpers; not a stand-alone header): the multi-GPU owner side — range partition (plan, chunked scatter + slab copies, owner level 2) and the streaming owner of the hash route.

// ---- multi-GPU: range partition, the level-1 scatter done by the SENDERS, the exchange by the copy engines --
// Every rank holds a shard of the reads.  Instead of routing keys to owners by hash and letting every owner run the
// level-1 scatter over what it received (an extra pass over all keys), the ranks agree on ONE plan for the whole key
// space — from the all-gathered coarse histograms, so every rank computes the same plan by itself — whose level-1
// buckets are dealt to the owners in consecutive, equally populated runs.  A sender's level-1 scatter is then the same
// kernel, at the same speed, as on one GPU: it writes into a LOCAL staging array laid out owner by owner, and each
// owner's slab of it crosses NVLink as one large device-to-device copy (the buckets this rank owns itself are
// scattered straight into its own receive buffer).  (The first form of this path stored every bucket run — ~250 B —
// into peer memory from the scatter kernel: 14.5 ms per 1e9 bases at 2 GPUs against 4.5 for the local scatter.)
// The input is cut into chunks: while chunk c + 1 is being scattered, chunk c is on the links and the owners run the
// level-2 scatter over chunk c - 1 on a second stream, so the exchange costs no SM time and hides behind the count.
// The owner runs fast_part2 + fast_finish only, and its table is the key range it owns: the ranks' tables, in rank
// order, are the globally sorted table.
//
// Receive buffer of an owner (kmc_recv_buffer, mapped by the peers with CUDA IPC):
//   [ cursor table: (chunk, bucket, sender) -> keys stored, u64, kDistHeader bytes ][ level-1 array ]
// level-1 array of owner o: for chunk c, for sender s, for bucket b of o: a region of cap(s, b) keys — so the slab
// (c, s) is contiguous, and is what sender s copies in one piece.
constexpr uint32_t kDistMaxWorld = 16, kDistMaxChunks = 16;
constexpr size_t kDistHeader = (size_t)kMaxL1 * kDistMaxWorld * kDistMaxChunks * 8; // 2 MB: any owner may hold most buckets

// sum of a u64 array (the keys an owner received = the sum of its cursor table)
__global__ void __launch_bounds__(256) sum_u64_kernel(const unsigned long long *__restrict__ v, uint64_t n, unsigned long long *__restrict__ total) {
  unsigned long long s = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) s += v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(total, s);
}

__global__ void dist_publish_kernel(const unsigned long long *__restrict__ cursor, const uint64_t *__restrict__ cap,
                                    const uint32_t *__restrict__ own_lo, const uint64_t *__restrict__ peer_header,
                                    uint32_t n_all, uint32_t world, uint32_t rank, uint32_t chunk) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_all) return;
  uint32_t o = 0;
  while (o + 1 < world && own_lo[o + 1] <= b) o++;
  unsigned long long v = cursor[b];
  if (v > cap[b]) v = cap[b]; // overflow was flagged by the scatter; the job is recounted
  const uint32_t my_n = own_lo[o + 1] - own_lo[o];
  unsigned long long *dst = reinterpret_cast<unsigned long long *>(peer_header[o]) + ((size_t)chunk * my_n + (b - own_lo[o])) * world + rank;
  *dst = v;
}

template <typename KeyT>
int dist_hist_impl(kmc_ctx *c, uint64_t *hist_out, uint32_t *low_cardinality) {
  const uint32_t ncoarse = 1u << coarse_bits(c);
  TRY(zero_scalars(c));
  bool low = c->cfg.strategy == KMC_STRATEGY_HASH;
  if (c->cfg.strategy == KMC_STRATEGY_AUTO && sizeof(KeyT) == 8) TRY(hash_probe(c, &low));
  KeyArrays ka;
  std::vector<uint64_t> hist;
  uint32_t step = 1;
  TRY(coarse_hist<KeyT>(c, ka, hist, &step));
  for (uint32_t i = 0; i < 4096; i++) {
    double est = i < ncoarse ? (double)hist[i] * step : 0.0;
    if (step > 1 && i < ncoarse) est += 5.0 * std::sqrt(est * step) + step; // upper estimate, as in finish_fast
    hist_out[i] = (uint64_t)est;
  }
  if (low_cardinality) *low_cardinality = low ? 1u : 0u;
  return KMC_OK;
}

template <typename KeyT>
int dist_plan_impl(kmc_ctx *c, uint32_t world, uint32_t rank, const uint64_t *all_hist, uint32_t n_chunks, uint64_t *need_bytes) {
  constexpr bool kWide = sizeof(KeyT) == 16;
  int kTarget = kWide ? 3200 : kFineTarget;
  const uint32_t kb = c->key_bits, cb = coarse_bits(c), ncoarse = 1u << cb;
  DistPlan &D = c->dist;
  D.valid = false; D.scattered = false;
  for (uint32_t o = 0; o < world; o++) need_bytes[o] = 0;
  std::vector<uint64_t> G(ncoarse, 0);
  uint64_t n_est = 0;
  for (uint32_t s = 0; s < world; s++)
    for (uint32_t ci = 0; ci < ncoarse; ci++) { G[ci] += all_hist[(size_t)s * 4096 + ci]; n_est += all_hist[(size_t)s * 4096 + ci]; }
  if (n_est < ((uint64_t)world << 20)) return KMC_OK; // small job: not worth a plan
  PlanShape shape;
  // at least 64 level-1 buckets per owner, so that owners can be balanced to a few percent
  uint32_t b1_min = 6;
  while ((1u << (b1_min - 6)) < world) b1_min++;
  const uint32_t min_e = getenv("KMC_NO_SPLIT64") ? 0u : split64_min_e(kb, n_est, kWide);
  if (!plan_shape(G, kb, 0, ncoarse, false, kTarget, shape, b1_min, min_e)) return KMC_OK;
  D.fine_cap = kWide ? 4096 : kFineCap;
  if (!kWide && (kFineCap64 != kFineCap || kFineTarget64 != kFineTarget)) {
    // buckets that leave more than 32 key bits are sorted as 64-bit elements, whose bucket shape is smaller: plan again
    // (as fast_begin does; every rank sees the same global histogram, so every rank decides the same)
    bool wide_elems = false;
    for (uint32_t b = 0; b < shape.n_l1; b++) if (kb - shape.b1 - shape.l1e[b] > 32) wide_elems = true;
    if (wide_elems) {
      kTarget = kFineTarget64;
      D.fine_cap = kFineCap64;
      if (!plan_shape(G, kb, 0, ncoarse, false, kTarget, shape, b1_min, min_e)) return KMC_OK;
    }
  }
  const uint32_t b1 = shape.b1, n_all = 1u << b1, cshift = cb - b1;
  if (n_all < world) return KMC_OK;
  // owners: consecutive level-1 buckets, about equal population
  std::vector<uint64_t> pop(n_all, 0);
  unsigned __int128 total = 0;
  for (uint32_t b = 0; b < n_all; b++) {
    for (uint32_t ci = b << cshift; ci < ((b + 1) << cshift); ci++) pop[b] += G[ci];
    total += pop[b];
  }
  D.own_lo.assign(world + 1, 0);
  {
    unsigned __int128 before = 0;
    uint32_t b = 0;
    for (uint32_t o = 1; o < world; o++) {
      const unsigned __int128 want = (total * o + world - 1) / world;
      while (b < n_all && before < want) before += pop[b++];
      D.own_lo[o] = b;
    }
    D.own_lo[world] = n_all;
  }
  for (uint32_t o = 0; o < world; o++) // the owner's cursor table must hold (chunk, bucket, sender)
    if ((uint64_t)n_chunks * (D.own_lo[o + 1] - D.own_lo[o]) * world > kDistHeader / 8) return KMC_OK;
  // chunks: equal slices of every sender's input, except that the first and the last are half as long — the exchange is
  // a chain (scatter chunk 0, then one copy after the other, then the level-2 scatter of the last chunk), and its two
  // ends are the part nothing overlaps
  D.cum_frac.assign(n_chunks + 1, 0.0);
  {
    const double unit = n_chunks > 2 ? 1.0 / (n_chunks - 1) : 1.0 / n_chunks;
    for (uint32_t ch = 0; ch < n_chunks; ch++)
      D.cum_frac[ch + 1] = D.cum_frac[ch] + ((n_chunks > 2 && (ch == 0 || ch + 1 == n_chunks)) ? 0.5 * unit : unit);
    D.cum_frac[n_chunks] = 1.0;
  }
  // region (chunk, sender, bucket): capacity from that sender's own histogram and the chunk's share of its input
  D.s_cap.assign((size_t)n_chunks * n_all, 0); D.s_in.assign((size_t)n_chunks * n_all, 0);
  D.slab_pre.assign((size_t)n_chunks * world, 0); D.slab_len.assign((size_t)n_chunks * world, 0);
  D.chunk_off.assign((size_t)n_chunks * world, 0); D.stage_off.assign((size_t)n_chunks * world, 0);
  D.x_cap.clear(); D.x_off.clear();
  const uint64_t slack = 2 * kMaxTile;
  std::vector<uint64_t> stage(n_chunks, 0);
  for (uint32_t o = 0; o < world; o++) {
    uint64_t off = 0;
    const uint32_t my_n = D.own_lo[o + 1] - D.own_lo[o];
    if (o == rank) { D.x_cap.assign((size_t)n_chunks * my_n * world, 0); D.x_off.assign((size_t)n_chunks * my_n * world, 0); }
    for (uint32_t ch = 0; ch < n_chunks; ch++) {
      const double frac = D.cum_frac[ch + 1] - D.cum_frac[ch];
      D.chunk_off[(size_t)ch * world + o] = off;
      for (uint32_t s = 0; s < world; s++) {
        const uint64_t slab0 = off;
        for (uint32_t b = D.own_lo[o]; b < D.own_lo[o + 1]; b++) {
          uint64_t nb = 0;
          for (uint32_t ci = b << cshift; ci < ((b + 1) << cshift); ci++) nb += all_hist[(size_t)s * 4096 + ci];
          const uint64_t cap1 = n_chunks > 1 ? (((uint64_t)((double)nb * frac * 1.04) + 2048 + 15) & ~15ull)
                                             : (((uint64_t)((double)nb * 1.03) + 4096 + 15) & ~15ull);
          if (s == rank) { D.s_in[(size_t)ch * n_all + b] = off - slab0; D.s_cap[(size_t)ch * n_all + b] = cap1; }
          if (o == rank) {
            const size_t x = ((size_t)ch * my_n + (b - D.own_lo[o])) * world + s;
            D.x_cap[x] = cap1; D.x_off[x] = off;
          }
          off += cap1;
        }
        if (s == rank) { D.slab_pre[(size_t)ch * world + o] = slab0; D.slab_len[(size_t)ch * world + o] = off - slab0; }
      }
      if (o != rank) { D.stage_off[(size_t)ch * world + o] = stage[ch]; stage[ch] += D.slab_len[(size_t)ch * world + o]; }
    }
    need_bytes[o] = kDistHeader + (off + slack) * sizeof(KeyT);
    if (o == rank) D.l1_keys = off;
  }
  D.stage_len = 0;
  for (uint64_t v : stage) D.stage_len = std::max(D.stage_len, v);
  D.world = world; D.rank = rank; D.b1 = b1; D.n_all = n_all; D.n_chunks = n_chunks;
  D.l1e = shape.l1e;
  D.fine_hist = G;
  D.owner_ready = false; D.chunks_sent = 0; D.chunks_owned = 0;
  D.valid = true;
  return KMC_OK;
}

struct StreamSwap {   // helpers launch on c->stream: run them on another stream of the ctx for a while
  kmc_ctx *c; cudaStream_t saved;
  StreamSwap(kmc_ctx *c_, cudaStream_t s) : c(c_), saved(c_->stream) { c->stream = s; }
  ~StreamSwap() { c->stream = saved; }
};

// The owner's plan over what the senders will leave in the receive buffer: one pseudo-bucket per (chunk, bucket,
// sender) region, all regions of a bucket feeding the same fine buckets.  Tables, buffers, descriptors; no key is touched.
template <typename KeyT>
int dist_owner_begin(kmc_ctx *c) {
  constexpr bool kWide = sizeof(KeyT) == 16;
  DistPlan &D = c->dist;
  DistOwner &O = D.owner;
  const int kCap = (int)D.fine_cap;
  const uint32_t kb = c->key_bits, cb = coarse_bits(c), b1 = D.b1, cshift = cb - b1, world = D.world, C = D.n_chunks;
  const uint32_t my_lo = D.own_lo[D.rank], my_n = D.own_lo[D.rank + 1] - my_lo, n_xc = my_n * world, n_x = n_xc * C, n_cb = my_n << cshift;
  if (!c->recv_keys.p || c->recv_keys.cap < kDistHeader + (D.l1_keys + 2 * kMaxTile) * sizeof(KeyT))
    return fail(c, KMC_E_ARG, "kmc_dist_scatter: the receive buffer is smaller than kmc_dist_plan asked for");
  auto al16 = [](size_t x) { return (x + 15) & ~size_t(15); };
  // owner tables: per pseudo-bucket x = (chunk, bucket, sender): start | cap | tile0 | fine0 | e;  per bucket: fine0 | e;
  // per coarse bin: start | fine0 | cap
  const size_t o_xs = 0, o_xc = o_xs + al16((size_t)(n_x + 1) * 8), o_xt = o_xc + al16((size_t)n_x * 8),
               o_xf = o_xt + al16((size_t)(n_x + 1) * 4), o_xe = o_xf + al16((size_t)(n_x + 1) * 4), o_rf = o_xe + al16(n_x),
               o_re = o_rf + al16((size_t)(my_n + 1) * 4), o_cs = o_re + al16(my_n), o_cf = o_cs + al16((size_t)n_cb * 8),
               o_cc = o_cf + al16((size_t)n_cb * 4), tab_bytes = o_cc + al16((size_t)n_cb * 2);
  c->fast_host.assign(tab_bytes, 0);
  unsigned char *hb = c->fast_host.data();
  uint64_t *xs = (uint64_t *)(hb + o_xs), *xc = (uint64_t *)(hb + o_xc);
  uint32_t *xf = (uint32_t *)(hb + o_xf), *rf = (uint32_t *)(hb + o_rf);
  uint8_t *xe = hb + o_xe, *re = hb + o_re;
  uint64_t *cstart = (uint64_t *)(hb + o_cs);
  uint32_t *cfine0 = (uint32_t *)(hb + o_cf);
  uint16_t *ccap = (uint16_t *)(hb + o_cc);
  uint64_t l2_keys = 0, tiles2 = 0, t_max = 1;
  uint32_t fb = 0, nb_max = 1;
  bool key32 = !kWide, split64 = !kWide && !getenv("KMC_NO_SPLIT64");
  for (uint32_t rb = 0; rb < my_n; rb++) {
    if (kb - b1 - D.l1e[my_lo + rb] > 32) key32 = false;
    if (kb - b1 - D.l1e[my_lo + rb] > 32 + (uint32_t)kFinishBits) split64 = false;
  }
  if (key32) split64 = false;
  for (uint32_t rb = 0; rb < my_n; rb++) {
    const uint32_t b = my_lo + rb, e = D.l1e[b], sub_bits = e - cshift;
    rf[rb] = fb; re[rb] = (uint8_t)e;
    for (uint32_t ch = 0; ch < C; ch++)
      for (uint32_t s = 0; s < world; s++) {
        const uint32_t x = ch * n_xc + rb * world + s;
        const uint64_t cap1 = D.x_cap[((size_t)ch * my_n + rb) * world + s];
        xs[x] = D.x_off[((size_t)ch * my_n + rb) * world + s];
        xc[x] = cap1; xf[x] = fb; xe[x] = (uint8_t)e;
        tiles2 += (cap1 + p2_tile<KeyT>() - 1) / p2_tile<KeyT>();
        t_max = std::max<uint64_t>(t_max, (cap1 + p2_tile<KeyT>() - 1) / p2_tile<KeyT>());
      }
    nb_max = std::max<uint32_t>(nb_max, 1u << e);
    for (uint32_t ci = b << cshift; ci < ((b + 1) << cshift); ci++) {
      double avg = (double)D.fine_hist[ci] / (double)(1ull << sub_bits);
      const uint32_t cp = fine_cap_for(avg, kCap);
      const uint32_t ci_rel = ci - (my_lo << cshift);
      cstart[ci_rel] = l2_keys; cfine0[ci_rel] = fb; ccap[ci_rel] = (uint16_t)cp;
      l2_keys += (uint64_t)cp << sub_bits;
      fb += 1u << sub_bits;
    }
  }
  const uint64_t l1_keys = D.l1_keys;
  xs[n_x] = l1_keys; xf[n_x] = fb; rf[my_n] = fb;
  const uint64_t n_fine = fb;
  if (tiles2 > 0x7FFFFFFFull || n_fine == 0) return fail(c, KMC_E_CAPACITY, "kmc_dist_scatter: range-partition plan too large");
  const uint64_t slack = 2 * kMaxTile;
  const size_t off_ticket = 4096 * 8, off_dtotal = off_ticket + 8, off_l1cur = off_dtotal + 8, off_fine = off_l1cur + kMaxL1 * 8;
  const size_t off_status = (off_fine + n_fine * 4 + 15) & ~size_t(15);
  TRY(ensure(c, c->fast_tables, tab_bytes));
  TRY(ensure(c, c->fast_fdesc, n_fine * sizeof(FineDesc)));
  TRY(ensure(c, c->fast_l2, (l2_keys + 2 * slack) * sizeof(KeyT)));
  TRY(ensure(c, c->t_lo, l1_keys * 8 + 64));
  if (kWide) TRY(ensure(c, c->t_hi, l1_keys * 8 + 64));
  TRY(ensure(c, c->t_cnt, l1_keys * 4 + 64));
  TRY(ensure(c, c->fast_state, off_status + n_fine * 8 + 64));
  CK(cudaMemsetAsync(c->fast_state.p, 0, off_status + n_fine * 8, c->stream));
  TRY(h2d_small(c, c->fast_tables.p, c->fast_host.data(), tab_bytes));
  unsigned char *st = (unsigned char *)c->fast_state.p, *tb = (unsigned char *)c->fast_tables.p;
  FastPlan &pl = O.pl;
  pl = FastPlan{};
  pl.kb = kb; pl.b1 = b1; pl.n_l1 = n_x; pl.n_fine = (uint32_t)n_fine; pl.l1_base = my_lo;
  pl.l1_trash = l1_keys; pl.l2_trash = l2_keys + slack;
  pl.fdesc = (const FineDesc *)c->fast_fdesc.p;
  pl.l1_start = (const uint64_t *)(tb + o_xs); pl.l1_cap = (const uint64_t *)(tb + o_xc);
  pl.l1_tile0 = (const uint32_t *)(tb + o_xt); pl.l1_fine0 = (const uint32_t *)(tb + o_xf); pl.l1_e = tb + o_xe;
  pl.l1_cursor = (unsigned long long *)c->recv_keys.p; pl.fine_cursor = (uint32_t *)(st + off_fine);
  O.ticket = (unsigned int *)(st + off_ticket);
  O.d_total = (unsigned long long *)(st + off_dtotal);
  O.status = (unsigned long long *)(st + off_status);
  O.key32 = key32; O.split64 = split64; O.nb_max = nb_max; O.t_max = t_max; O.n_fine = n_fine; O.n_xc = n_xc;
  LAUNCH(plan_expand_kernel, n_cb, 128, 0, (FineDesc *)c->fast_fdesc.p, (const uint64_t *)(tb + o_cs), (const uint32_t *)(tb + o_cf),
         (const uint16_t *)(tb + o_cc), (const uint32_t *)(tb + o_rf), (const uint8_t *)(tb + o_re), cshift, my_lo, kb, b1, (uint32_t)kWide);
  c->launches--;
  if (!c->owner_stream) CK(cudaStreamCreateWithFlags(&c->owner_stream, cudaStreamNonBlocking));
  if (!D.ev_ready) CK(cudaEventCreateWithFlags(&D.ev_ready, cudaEventDisableTiming));
  CK(cudaEventRecord(D.ev_ready, c->stream));
  CK(cudaStreamWaitEvent(c->owner_stream, D.ev_ready, 0));
  c->fast_variant = kWide ? "u128" : key32 ? "u32" : split64 ? "split64" : "u64";
  D.owner_ready = true;
  return KMC_OK;
}

// sender: level-1 scatter of input chunk `chunk` into the staging array (own buckets: into the own receive buffer), then
// — on the peer stream, so that the next chunk's scatter runs meanwhile — one copy per owner and the chunk's cursors.
template <typename KeyT>
int dist_scatter_part_impl(kmc_ctx *c, void *const *peer_buf, uint32_t chunk) {
  DistPlan &D = c->dist;
  const uint32_t n_all = D.n_all, world = D.world, kb = c->key_bits, C = D.n_chunks;
  if (chunk != D.chunks_sent || chunk >= C) return fail(c, KMC_E_ARG, "kmc_dist_scatter_part: chunks go in order, 0..%u", C - 1);
  auto al16 = [](size_t x) { return (x + 15) & ~size_t(15); };
  // sender tables: per chunk l1_start (absolute address / key size) | l1_cap | own_lo | peer header pointers
  const size_t o_s = 0, o_c = o_s + al16((size_t)C * (n_all + 1) * 8), o_own = o_c + al16((size_t)C * n_all * 8),
               o_ph = o_own + al16((size_t)(world + 1) * 4), tab_bytes = o_ph + al16((size_t)world * 8);
  if (chunk == 0) {
    TRY(zero_scalars(c));
    TRY(dist_owner_begin<KeyT>(c)); // buffers first: nothing below may be freed under a running kernel
    TRY(ensure(c, c->dist_stage, (2 * D.stage_len + 64) * sizeof(KeyT)));
    TRY(ensure(c, c->dist_tables, tab_bytes));
    TRY(ensure(c, c->route_keys, (size_t)2 * kMaxTile * sizeof(KeyT) + 256)); // trash area for runs that do not fit
    TRY(ensure(c, c->dist_cursors, (size_t)kDistMaxChunks * kMaxL1 * 8));
    std::vector<unsigned char> host(tab_bytes, 0);
    uint64_t *l1s = (uint64_t *)(host.data() + o_s), *l1c = (uint64_t *)(host.data() + o_c);
    uint32_t *own = (uint32_t *)(host.data() + o_own);
    uint64_t *ph = (uint64_t *)(host.data() + o_ph);
    for (uint32_t ch = 0; ch < C; ch++)
      for (uint32_t o = 0; o < world; o++) {
        const size_t co = (size_t)ch * world + o;
        const uint64_t base = o == D.rank
            ? ((uint64_t)(uintptr_t)c->recv_keys.p + kDistHeader) / sizeof(KeyT) + D.slab_pre[co]
            : (uint64_t)(uintptr_t)c->dist_stage.p / sizeof(KeyT) + (uint64_t)(ch & 1) * D.stage_len + D.stage_off[co];
        for (uint32_t b = D.own_lo[o]; b < D.own_lo[o + 1]; b++) l1s[(size_t)ch * (n_all + 1) + b] = base + D.s_in[(size_t)ch * n_all + b];
      }
    for (uint32_t ch = 0; ch < C; ch++)
      for (uint32_t b = 0; b < n_all; b++) l1c[(size_t)ch * n_all + b] = D.s_cap[(size_t)ch * n_all + b];
    for (uint32_t o = 0; o < world; o++) ph[o] = (uint64_t)(uintptr_t)peer_buf[o];
    for (uint32_t o = 0; o <= world; o++) own[o] = D.own_lo[o];
    CK(cudaMemsetAsync(c->dist_cursors.p, 0, (size_t)C * kMaxL1 * 8, c->stream));
    TRY(h2d_small(c, c->dist_tables.p, host.data(), tab_bytes));
    if (!c->peer_stream) CK(cudaStreamCreateWithFlags(&c->peer_stream, cudaStreamNonBlocking));
    for (uint32_t ch = 0; ch < C; ch++)
      for (cudaEvent_t *e : {&D.ev_scattered[ch], &D.ev_copied[ch]})
        if (!*e) CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
  unsigned char *tb = (unsigned char *)c->dist_tables.p;
  FastPlan pl{};
  pl.kb = kb; pl.b1 = D.b1; pl.n_l1 = n_all; pl.n_fine = 0; pl.l1_base = 0;
  pl.l1_trash = ((uint64_t)(uintptr_t)c->route_keys.p + sizeof(KeyT) - 1) / sizeof(KeyT);
  pl.l1_start = (const uint64_t *)(tb + o_s) + (size_t)chunk * (n_all + 1);
  pl.l1_cap = (const uint64_t *)(tb + o_c) + (size_t)chunk * n_all;
  pl.l1_cursor = (unsigned long long *)c->dist_cursors.p + (size_t)chunk * kMaxL1;
  // the staging half this chunk scatters into was copied out two chunks ago
  if (chunk >= 2) CK(cudaStreamWaitEvent(c->stream, D.ev_copied[chunk - 2], 0));
  // this chunk's share of the CTA tiles of all segments, in segment order
  uint64_t all_ct = 0;
  for (size_t i = 0; i < c->n_segs; i++)
    if (c->segs[i].n_bases) all_ct += (num_warp_tiles(c->segs[i].n_bases, win_lanes<KeyT>()) + kFastWarps - 1) / kFastWarps;
  const uint64_t g0 = (uint64_t)((double)all_ct * D.cum_frac[chunk]), g1 = chunk + 1 == C ? all_ct : (uint64_t)((double)all_ct * D.cum_frac[chunk + 1]);
  PHASE_BEGIN("route");
  {
    size_t smem = L1Smem<KeyT>::bytes(part1_stage<KeyT>(), n_all);
    auto fast_scatter_to_owners = fast_part1_kernel<KeyT, true, PrefixBucket>;
    CK(cudaFuncSetAttribute(fast_scatter_to_owners, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const PrefixBucket bucket = make_prefix_bucket<false>(kb, D.b1);
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
      uint32_t grid = (uint32_t)std::min<uint64_t>(hi - lo, (uint64_t)c->n_sms);
      LAUNCH(fast_scatter_to_owners, grid, kFastThreads, smem, P, tiles, pl, bucket, (KeyT *)nullptr, d_err(c), lo - (seg0 - n_ct), hi - (seg0 - n_ct));
    }
  }
  PHASE_END();
  CK(cudaEventRecord(D.ev_scattered[chunk], c->stream));
  CK(cudaStreamWaitEvent(c->peer_stream, D.ev_scattered[chunk], 0));
  const KeyT *stage = (const KeyT *)c->dist_stage.p + (size_t)(chunk & 1) * D.stage_len;
  // the slabs leave staggered, so that at any moment every rank writes to a different peer.  KMC_PEER_STREAMS > 1 puts
  // them on several streams at once (the first then waits for the others); measured at 8 GPUs it does not help — 28.9
  // ms/step with 4 streams against 27.5 with one: the links, not a copy engine, are the bound (~480 GB/s leave a GPU)
  static const int n_lanes = std::max(1, std::min(env_int("KMC_PEER_STREAMS", 1), 8));
  for (int l = 1; l < n_lanes; l++) {
    if (!c->peer_lane[l]) CK(cudaStreamCreateWithFlags(&c->peer_lane[l], cudaStreamNonBlocking));
    if (!c->peer_lane_ev[l]) CK(cudaEventCreateWithFlags(&c->peer_lane_ev[l], cudaEventDisableTiming));
    CK(cudaStreamWaitEvent(c->peer_lane[l], D.ev_scattered[chunk], 0));
  }
  for (uint32_t d = 1; d < world; d++) {
    const uint32_t o = (D.rank + d) % world;
    const size_t co = (size_t)chunk * world + o;
    if (!D.slab_len[co]) continue;
    const int l = (int)((d - 1) % (uint32_t)n_lanes);
    KeyT *dst = (KeyT *)((unsigned char *)peer_buf[o] + kDistHeader) + D.slab_pre[co];
    CK(cudaMemcpyAsync(dst, stage + D.stage_off[co], D.slab_len[co] * sizeof(KeyT), cudaMemcpyDeviceToDevice, l ? c->peer_lane[l] : c->peer_stream));
  }
  for (int l = 1; l < n_lanes; l++) {
    CK(cudaEventRecord(c->peer_lane_ev[l], c->peer_lane[l]));
    CK(cudaStreamWaitEvent(c->peer_stream, c->peer_lane_ev[l], 0));
  }
  {
    StreamSwap sw(c, c->peer_stream);
    const bool kt = c->ktiming;
    c->ktiming = false; // per-kernel event pairs belong to the compute stream
    LAUNCH(dist_publish_kernel, grid_for(n_all, 256), 256, 0, pl.l1_cursor, pl.l1_cap, (const uint32_t *)(tb + o_own),
           (const uint64_t *)(tb + o_ph), n_all, world, D.rank, chunk);
    c->launches--;
    c->ktiming = kt;
  }
  CK(cudaEventRecord(D.ev_copied[chunk], c->peer_stream));
  D.chunks_sent = chunk + 1;
  return KMC_OK;
}

// owner: level-2 scatter over the regions of one chunk (every sender's copy of it has landed: the caller's hand-over)
template <typename KeyT>
int dist_owner_part_impl(kmc_ctx *c, uint32_t chunk) {
  DistPlan &D = c->dist;
  DistOwner &O = D.owner;
  if (!D.owner_ready) return fail(c, KMC_E_ARG, "kmc_dist_owner_part before kmc_dist_scatter_part");
  if (chunk != D.chunks_owned || chunk >= D.n_chunks) return fail(c, KMC_E_ARG, "kmc_dist_owner_part: chunks go in order");
  StreamSwap sw(c, c->owner_stream);
  FastPlan pl = O.pl;
  const size_t x0 = (size_t)chunk * O.n_xc;
  pl.n_l1 = O.n_xc;
  pl.l1_start += x0; pl.l1_cap += x0; pl.l1_tile0 += x0; pl.l1_fine0 += x0; pl.l1_e += x0; pl.l1_cursor += x0;
  const KeyT *l1 = (const KeyT *)((unsigned char *)c->recv_keys.p + kDistHeader);
  const bool kt = c->ktiming;
  c->ktiming = false;
  PHASE_BEGIN("fast_part2");
  int rc = launch_part2<KeyT>(c, pl, l1, O.key32, O.nb_max, O.t_max);
  c->ktiming = kt;
  if (rc) return rc;
  PHASE_END();
  D.chunks_owned = chunk + 1;
  return KMC_OK;
}

// sender: everything this rank had to store has landed; did it fit?
int dist_scatter_end_impl(kmc_ctx *c, uint32_t *overflow) {
  DistPlan &D = c->dist;
  if (D.chunks_sent != D.n_chunks) return fail(c, KMC_E_ARG, "kmc_dist_scatter_end: %u of %u chunks scattered", D.chunks_sent, D.n_chunks);
  CK(cudaStreamSynchronize(c->peer_stream));
  uint32_t err = 0;
  TRY(read_scalars(c, nullptr, &err));
  *overflow = (err & kFlagOverflow) ? 1u : 0u;
  if (*overflow) {
    if (c->owner_stream) CK(cudaStreamSynchronize(c->owner_stream));
  if (c->peer_stream) CK(cudaStreamSynchronize(c->peer_stream));
    TRY(zero_scalars(c));
  }
  D.scattered = !*overflow;
  return KMC_OK;
}

// the owner's last part: (the level-2 scatter of chunks not handed over one by one, then) the bucket sort
template <typename KeyT>
int finish_dist(kmc_ctx *c) {
  constexpr bool kWide = sizeof(KeyT) == 16;
  DistPlan &D = c->dist;
  DistOwner &O = D.owner;
  while (D.chunks_owned < D.n_chunks) TRY(dist_owner_part_impl<KeyT>(c, D.chunks_owned));
  const FastPlan &pl = O.pl;
  const uint64_t n_fine = O.n_fine;
  unsigned long long *header = (unsigned long long *)c->recv_keys.p;
  {
    StreamSwap sw(c, c->owner_stream);
    PHASE_BEGIN("fast_finish");
    unsigned long long *prof = nullptr;
    if constexpr (kWide) {
      size_t fsmem = sizeof(FinishSmem<U128>);
      auto fast_finish = fast_finish_kernel<U128>;
      CK(cudaFuncSetAttribute(fast_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
      LAUNCH(fast_finish, (uint32_t)std::min<uint64_t>(n_fine, (uint64_t)c->n_sms * 2), kFinThreads, fsmem, pl, (const U128 *)c->fast_l2.p,
             (uint64_t *)c->t_lo.p, (uint64_t *)c->t_hi.p, (uint32_t *)c->t_cnt.p, O.status, O.ticket, d_err(c), O.d_total, prof);
    } else if (O.key32) {
      size_t fsmem = sizeof(FinishSmem<uint32_t>);
      auto fast_finish = fast_finish_kernel<uint32_t>;
      CK(cudaFuncSetAttribute(fast_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
      LAUNCH(fast_finish, (uint32_t)std::min<uint64_t>(n_fine, (uint64_t)c->n_sms * KMC_FINISH_MINB32), kFinThreads, fsmem, pl,
             (const uint32_t *)c->fast_l2.p, (uint64_t *)c->t_lo.p, (uint64_t *)nullptr, (uint32_t *)c->t_cnt.p, O.status, O.ticket, d_err(c),
             O.d_total, prof);
    } else if (O.split64) {
      size_t fsmem = sizeof(FinishSmem<Split64>);
      auto fast_finish = fast_finish_kernel<Split64>;
      CK(cudaFuncSetAttribute(fast_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
      LAUNCH(fast_finish, (uint32_t)std::min<uint64_t>(n_fine, (uint64_t)c->n_sms * 2), kFinThreads, fsmem, pl, (const uint64_t *)c->fast_l2.p,
             (uint64_t *)c->t_lo.p, (uint64_t *)nullptr, (uint32_t *)c->t_cnt.p, O.status, O.ticket, d_err(c), O.d_total, prof);
    } else {
      size_t fsmem = sizeof(FinishSmem<uint64_t>);
      auto fast_finish = fast_finish_kernel<uint64_t>;
      CK(cudaFuncSetAttribute(fast_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
      LAUNCH(fast_finish, (uint32_t)std::min<uint64_t>(n_fine, (uint64_t)c->n_sms * KMC_FINISH_MINB64), kFinThreads, fsmem, pl, (const uint64_t *)c->fast_l2.p,
             (uint64_t *)c->t_lo.p, (uint64_t *)nullptr, (uint32_t *)c->t_cnt.p, O.status, O.ticket, d_err(c), O.d_total, prof);
    }
    PHASE_END();
    uint64_t d = 0;
    uint32_t err = 0;
    TRY(d2h_small(c, &d, O.d_total, 8));
    TRY(read_scalars(c, nullptr, &err));
    if (err & kFlagSpin) return fail(c, KMC_E_CUDA, "fast_finish: look-back did not make progress");
    if (err & kFlagOverflow) {
      c->fast_fallbacks++;
      TRY(zero_scalars(c));
      return fail(c, KMC_E_CAPACITY, "range-partitioned count: a fine bucket overflowed (recount through kmc_route_to_peers)");
    }
    // keys I own = what the senders' cursor table says
    uint64_t N = 0;
    const size_t n_x = (size_t)O.n_xc * D.n_chunks;
    CK(cudaMemsetAsync(d_total_all(c), 0, 8, c->stream));
    LAUNCH(sum_u64_kernel, std::min<uint32_t>(grid_for(n_x, 256), 64), 256, 0, header, (uint64_t)n_x, d_total_all(c));
    c->launches--;
    TRY(d2h_small(c, &N, d_total_all(c), 8));
    c->n_total = N; c->n_distinct = d;
  }
  c->strategy_used = KMC_STRATEGY_SORT;
  D.scattered = false; D.owner_ready = false;
  return KMC_OK;
}


// ---- streaming owner (multi-GPU): count what the other ranks route here WHILE they are still routing ------
// The routing pass is cut into chunks (kmc_route_to_peers_part); after every chunk the ranks agree on the counts and
// each owner feeds the keys that have just arrived to its partitioned count — level-1 scatter and the whole tiles of
// the level-2 scatter — on a second stream, beside the routing kernel of the next chunk (which leaves it some SMs).
// kmc_finish then only has the rest of the level-2 scatter and the bucket sort left.

template <typename KeyT>
int owner_begin_impl(kmc_ctx *c, const uint64_t *global_hist, uint32_t n_owners, uint32_t *streaming) {
  *streaming = 0;
  const uint32_t ncoarse = 1u << coarse_bits(c);
  // this owner's share of every coarse bin: the owner function is a hash, so 1 / n_owners of it, Poisson-distributed
  std::vector<uint64_t> hist(ncoarse);
  uint64_t n_est = 0;
  for (uint32_t ci = 0; ci < ncoarse; ci++) {
    const double m = (double)global_hist[ci] / n_owners;
    hist[ci] = (uint64_t)(m * 1.02 + 6.0 * std::sqrt(m) + 64.0);
    n_est += hist[ci];
  }
  if (n_est < (1u << 22)) return KMC_OK; // small job: not worth the choreography
  if (!c->owner_stream) CK(cudaStreamCreateWithFlags(&c->owner_stream, cudaStreamNonBlocking));
  CK(cudaStreamSynchronize(c->stream)); // buffers the plan touches may still be read by the previous job's tail
  StreamSwap sw(c, c->owner_stream);
  TRY(zero_scalars(c));
  bool ok = false;
  TRY(fast_begin<KeyT>(c, hist, n_est, 0, &ok));
  if (!ok) return KMC_OK;
  c->owner_on = true;
  c->owner_fed.clear();
  *streaming = 1;
  return KMC_OK;
}

template <typename KeyT>
int owner_feed_impl(kmc_ctx *c, const void *d_keys, uint64_t n) {
  if (!n) return KMC_OK;
  c->owner_fed.emplace_back(d_keys, n);
  StreamSwap sw(c, c->owner_stream);
  return fast_feed_array<KeyT>(c, d_keys, n, true);
}

template <typename KeyT>
int finish_impl(kmc_ctx *c);

template <typename KeyT>
int owner_finish(kmc_ctx *c) {
  bool used = false;
  {
    StreamSwap sw(c, c->owner_stream);
    TRY(fast_end<KeyT>(c, true, &used));
  }
  c->owner_on = false;
  if (used) return KMC_OK;
  // the count did not suit the partitioned path (a bucket overflowed): recount what was fed, any way that works
  CK(cudaStreamSynchronize(c->owner_stream));
  c->ingested.clear();
  for (auto &e : c->owner_fed) {
    if (!c->ingested.empty() && (const char *)c->ingested.back().first + c->ingested.back().second * sizeof(KeyT) == (const char *)e.first)
      c->ingested.back().second += e.second;       // chunks of one region are adjacent
    else c->ingested.push_back(e);
  }
  c->owner_fed.clear();
  TRY(zero_scalars(c));
  return finish_impl<KeyT>(c);
}
