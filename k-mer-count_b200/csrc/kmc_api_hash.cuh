// kmc_api_hash.cuh — a section of kmc_api.cu (included there, inside its anonymous namespace, after the ctx and its
// helpers; not a stand-alone header): the hash strategies — 64-bit keys, wider keys, (key, count) rows — and the cardinality probe.

// ---- hash strategy (kmc_hash.cuh), 64-bit keys ---------------------------------------------------------------
// step > 1: cardinality probe on a sample (table stays, nothing else is produced); *ok = table did not fill.
int hash_run(kmc_ctx *c, const KeyArrays &ka, uint32_t log2_slots, uint64_t limit, uint32_t step, bool *ok, HashTable *out,
             uint32_t n_hot = 0, bool throttle = true) {
  *ok = false;
  TRY(ensure(c, c->hash_hot, kHotMax * 8 + 64));
  const uint64_t *hot = (const uint64_t *)c->hash_hot.p;
  // the fill limit is only checked between tiles: keep the keys in flight (one tile per resident warp) well below
  // the table size, or a high-cardinality input would swamp the table before anybody notices
  // (the real run's table is sized from the probe, so only the probe itself — step > 1 or forced — is throttled)
  const uint64_t max_warps = std::max<uint64_t>(64, (1ull << log2_slots) / 4 / 992);
  const uint32_t max_ctas = throttle ? (uint32_t)std::min<uint64_t>((uint64_t)c->n_sms * 8, std::max<uint64_t>(8, max_warps / 8))
                                     : (uint32_t)c->n_sms * 8;
  const uint64_t slots = 1ull << log2_slots;
  TRY(ensure(c, c->hash_slots, slots * sizeof(HashSlot)));
  TRY(ensure(c, c->hash_scalars, 64));
  LAUNCH(hash_init_kernel, std::min<uint32_t>(grid_for(slots, 256), c->n_sms * 16), 256, 0, (HashSlot *)c->hash_slots.p, slots);
  CK(cudaMemsetAsync(c->hash_scalars.p, 0, 64, c->stream));
  HashTable T;
  T.slots = (HashSlot *)c->hash_slots.p;
  T.mask = slots - 1; T.shift = 64 - log2_slots;
  T.n_used = (unsigned long long *)c->hash_scalars.p; T.n_total = T.n_used + 1; T.n_ones = T.n_used + 2;
  T.limit = limit; T.flags = d_err(c);
  if (ka.from_array) {
    for (auto &a : ka.arrays) {
      uint32_t grid = (uint32_t)std::min<uint64_t>(grid_for(a.second, 1024ull * step), (uint64_t)max_ctas);
      LAUNCH(hash_count_array_kernel, grid, 256, 0, (const uint64_t *)a.first, a.second, step, T, hot, n_hot);
    }
  } else {
    for (size_t i = 0; i < c->n_segs; i++) {
      Segment &s = c->segs[i];
      if (!s.n_bases) continue;
      const bool sample_host = step > 1 && s.host_alias && s.wait_ready;
      if (!sample_host) TRY(seg_wait(c, s));
      ExtractParams P = seg_params(c, s, sample_host);
      uint64_t tiles = num_warp_tiles(s.n_bases, 31);
      uint32_t grid = (uint32_t)std::min<uint64_t>((tiles / step + 8) / 8, (uint64_t)max_ctas);
      auto hash_count = hash_count_kernel<true>;
      LAUNCH(hash_count, grid, 256, 0, P, tiles, step, T, hot, n_hot);
    }
  }
  uint32_t err = 0;
  TRY(read_scalars(c, nullptr, &err));
  if (err & kFlagHashFull) {
    CK(cudaMemsetAsync(d_err(c), 0, 4, c->stream)); // a full table is not an error: the caller picks another route
    return KMC_OK;
  }
  *ok = true;
  if (out) *out = T;
  return KMC_OK;
}

int finish_hash(kmc_ctx *c, uint32_t log2_slots, uint64_t limit, bool *used) {
  *used = false;
  KeyArrays ka;
  TRY(key_sources<uint64_t>(c, &ka));
  HashTable T;
  bool ok = false;
  PHASE_BEGIN("hash_count");
  TRY(hash_run(c, ka, log2_slots, limit, 1, &ok, &T, c->n_hot, /*throttle=*/c->probe_distinct == 0));
  PHASE_END();
  if (!ok) { c->hash_aborts++; return KMC_OK; }
  unsigned long long sc[3];
  TRY(d2h_small(c, sc, c->hash_scalars.p, sizeof sc));
  const uint64_t d = sc[0], n_total = sc[1], n_ones = sc[2];
  if (n_ones > 0xFFFFFFFFull) return fail(c, KMC_E_COUNT_OVERFLOW, "a k-mer occurs more than 2^32-1 times");
  PHASE_BEGIN("hash_sort");
  // distinct keys → dense array → sorted; the key arrays of key_sources() are no longer needed
  uint64_t *dense = nullptr, *scratch = nullptr, *sorted = nullptr;
  TRY(ensure(c, c->t_lo, (d + 2) * 8));
  TRY(ensure(c, c->keys_b, (d + 2) * 8));
  TRY(ensure(c, c->t_cnt, (d + 2) * 4));
  dense = (uint64_t *)c->t_lo.p; scratch = (uint64_t *)c->keys_b.p;
  if (d) {
    CK(cudaMemsetAsync(d_cursor(c), 0, 8, c->stream));
    LAUNCH(hash_compact_kernel, std::min<uint32_t>(grid_for(T.mask + 1, 256), c->n_sms * 16), 256, 0, T, dense, d_cursor(c));
    TRY(radix_sort<uint64_t>(c, dense, scratch, d, c->key_bits, &sorted));
    if (sorted != dense) CK(cudaMemcpyAsync(dense, sorted, d * 8, cudaMemcpyDeviceToDevice, c->stream));
    LAUNCH(hash_lookup_kernel, std::min<uint32_t>(grid_for(d, 256), c->n_sms * 16), 256, 0, T, (const uint64_t *)dense, d,
           (uint32_t *)c->t_cnt.p);
  }
  uint64_t rows = d;
  if (n_ones) { // the all-ones key (k = 32) sorts last
    uint64_t k1 = kHashEmpty;
    uint32_t c1 = (uint32_t)n_ones;
    CK(cudaMemcpyAsync(dense + d, &k1, 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync((uint32_t *)c->t_cnt.p + d, &c1, 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    rows++;
  }
  PHASE_END();
  c->n_total = n_total; c->n_distinct = rows;
  c->strategy_used = KMC_STRATEGY_HASH;
  *used = true;
  return KMC_OK;
}

// ---- hash strategy, keys of more than 64 bits (kmc_hash128.cuh) --------------------------------------------------
// The same steps as hash_run / finish_hash / hash_probe for 64-bit keys, without the hot-key counters.
int hash128_run(kmc_ctx *c, const KeyArrays &ka, uint32_t log2_slots, uint64_t limit, uint32_t step, bool *ok, HashTable128 *out,
                bool throttle = true) {
  *ok = false;
  const uint64_t max_warps = std::max<uint64_t>(64, (1ull << log2_slots) / 4 / 992);
  const uint32_t max_ctas = throttle ? (uint32_t)std::min<uint64_t>((uint64_t)c->n_sms * 8, std::max<uint64_t>(8, max_warps / 8))
                                     : (uint32_t)c->n_sms * 8;
  const uint64_t slots = 1ull << log2_slots;
  TRY(ensure(c, c->hash_slots, slots * sizeof(HashSlot128)));
  TRY(ensure(c, c->hash_scalars, 64));
  LAUNCH(hash128_init_kernel, std::min<uint32_t>(grid_for(slots, 256), c->n_sms * 16), 256, 0, (HashSlot128 *)c->hash_slots.p, slots);
  CK(cudaMemsetAsync(c->hash_scalars.p, 0, 64, c->stream));
  HashTable128 T;
  T.slots = (HashSlot128 *)c->hash_slots.p;
  T.mask = slots - 1; T.shift = 64 - log2_slots;
  T.n_used = (unsigned long long *)c->hash_scalars.p; T.n_total = T.n_used + 1; T.n_ones = T.n_used + 2;
  T.limit = limit; T.flags = d_err(c);
  if (ka.from_array) {
    for (auto &a : ka.arrays) {
      uint32_t grid = (uint32_t)std::min<uint64_t>(grid_for(a.second, 1024ull * step), (uint64_t)max_ctas);
      LAUNCH(hash128_count_array_kernel, grid, 256, 0, (const U128 *)a.first, a.second, step, T);
    }
  } else {
    for (size_t i = 0; i < c->n_segs; i++) {
      Segment &s = c->segs[i];
      if (!s.n_bases) continue;
      const bool sample_host = step > 1 && s.host_alias && s.wait_ready;
      if (!sample_host) TRY(seg_wait(c, s));
      ExtractParams P = seg_params(c, s, sample_host);
      uint64_t tiles = num_warp_tiles(s.n_bases, win_lanes<U128>());
      uint32_t grid = (uint32_t)std::min<uint64_t>((tiles / step + 8) / 8, (uint64_t)max_ctas);
      auto hash128_count = hash128_count_kernel<true>;
      LAUNCH(hash128_count, grid, 256, 0, P, tiles, step, T);
    }
  }
  uint32_t err = 0;
  TRY(read_scalars(c, nullptr, &err));
  if (err & kFlagHashFull) {
    CK(cudaMemsetAsync(d_err(c), 0, 4, c->stream)); // a full table is not an error: the caller picks another route
    return KMC_OK;
  }
  *ok = true;
  if (out) *out = T;
  return KMC_OK;
}

int finish_hash128(kmc_ctx *c, uint32_t log2_slots, uint64_t limit, bool *used) {
  *used = false;
  KeyArrays ka;
  TRY(key_sources<U128>(c, &ka));
  HashTable128 T;
  bool ok = false;
  PHASE_BEGIN("hash_count");
  TRY(hash128_run(c, ka, log2_slots, limit, 1, &ok, &T, /*throttle=*/c->probe_distinct == 0));
  PHASE_END();
  if (!ok) { c->hash_aborts++; return KMC_OK; }
  unsigned long long sc[3];
  TRY(d2h_small(c, sc, c->hash_scalars.p, sizeof sc));
  const uint64_t d = sc[0], n_total = sc[1], n_ones = sc[2];
  if (n_ones > 0xFFFFFFFFull) return fail(c, KMC_E_COUNT_OVERFLOW, "a k-mer occurs more than 2^32-1 times");
  PHASE_BEGIN("hash_sort");
  // distinct keys → dense array → sorted → the table's columns; the key arrays of key_sources() are no longer needed
  TRY(ensure(c, c->keys_a, (d + 2) * sizeof(U128)));
  TRY(ensure(c, c->keys_b, (d + 2) * sizeof(U128)));
  TRY(ensure(c, c->t_lo, (d + 2) * 8));
  TRY(ensure(c, c->t_hi, (d + 2) * 8));
  TRY(ensure(c, c->t_cnt, (d + 2) * 4));
  if (d) {
    U128 *dense = (U128 *)c->keys_a.p, *sorted = nullptr;
    CK(cudaMemsetAsync(d_cursor(c), 0, 8, c->stream));
    LAUNCH(hash128_compact_kernel, std::min<uint32_t>(grid_for(T.mask + 1, 256), c->n_sms * 16), 256, 0, T, dense, d_cursor(c));
    TRY(radix_sort<U128>(c, dense, (U128 *)c->keys_b.p, d, c->key_bits, &sorted));
    LAUNCH(hash128_lookup_kernel, std::min<uint32_t>(grid_for(d, 256), c->n_sms * 16), 256, 0, T, (const U128 *)sorted, d,
           (uint64_t *)c->t_lo.p, (uint64_t *)c->t_hi.p, (uint32_t *)c->t_cnt.p);
  }
  uint64_t rows = d;
  if (n_ones) { // the all-ones key (128 key bits: k = 64, non-canonical poly-T) sorts last
    uint64_t k1 = kHashEmpty;
    uint32_t c1 = (uint32_t)n_ones;
    CK(cudaMemcpyAsync((uint64_t *)c->t_lo.p + d, &k1, 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync((uint64_t *)c->t_hi.p + d, &k1, 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync((uint32_t *)c->t_cnt.p + d, &c1, 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    rows++;
  }
  PHASE_END();
  c->n_total = n_total; c->n_distinct = rows;
  c->strategy_used = KMC_STRATEGY_HASH;
  *used = true;
  return KMC_OK;
}

// cardinality probe for wide keys: the same insert kernel on a 1-in-step sample into a 2^22-slot table with a fill limit
int hash128_probe(kmc_ctx *c, bool *low_cardinality) {
  *low_cardinality = false;
  c->probe_distinct = 0;
  c->n_hot = 0;
  KeyArrays ka;
  ka.from_array = !c->ingested.empty();
  if (c->cfg.mode == KMC_MODE_LR_GAPPED) return KMC_OK; // keys would have to be materialised first: skip the probe
  if (ka.from_array) for (auto &e : c->ingested) if (e.second) { ka.arrays.emplace_back(e.first, e.second); ka.n += e.second; }
  const uint64_t n_in = ka.from_array ? ka.n : c->total_bases;
  if (n_in < (1u << 18)) return KMC_OK;
  uint32_t step = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(kHistSampleMax, n_in >> 26));
  bool ok = false;
  HashTable128 T;
  PHASE_BEGIN("hash_probe");
  TRY(hash128_run(c, ka, 22, 1ull << 20, step, &ok, &T));
  if (ok) {
    unsigned long long sc[3];
    TRY(d2h_small(c, sc, c->hash_scalars.p, sizeof sc));
    c->probe_distinct = sc[0];
  }
  PHASE_END();
  *low_cardinality = ok;
  return KMC_OK;
}

// (key, count) rows handed over with kmc_ingest_pairs: merge them through the hash table (equal keys add up), then
// the same compaction / sort / look-up as finish_hash.  Kept apart from finish_hash on purpose: that one is the measured
// single-GPU path.
int finish_pairs(kmc_ctx *c) {
  uint64_t rows_in = 0;
  for (auto &a : c->ingested_pairs) rows_in += a.n;
  uint32_t lg = 10;
  while (lg < 33 && (1ull << lg) < 2 * rows_in + 2) lg++; // load factor <= 1/2 even if every row is a distinct key
  const uint64_t slots = 1ull << lg;
  TRY(ensure(c, c->hash_slots, slots * sizeof(HashSlot)));
  TRY(ensure(c, c->hash_scalars, 64));
  PHASE_BEGIN("hash_count");
  LAUNCH(hash_init_kernel, std::min<uint32_t>(grid_for(slots, 256), c->n_sms * 16), 256, 0, (HashSlot *)c->hash_slots.p, slots);
  CK(cudaMemsetAsync(c->hash_scalars.p, 0, 64, c->stream));
  HashTable T;
  T.slots = (HashSlot *)c->hash_slots.p;
  T.mask = slots - 1; T.shift = 64 - lg;
  T.n_used = (unsigned long long *)c->hash_scalars.p; T.n_total = T.n_used + 1; T.n_ones = T.n_used + 2;
  T.limit = slots; T.flags = d_err(c);
  for (auto &a : c->ingested_pairs)
    if (a.n) LAUNCH(hash_count_pairs_kernel, std::min<uint32_t>(grid_for(a.n, 1024), c->n_sms * 8), 256, 0, a.keys, a.counts, a.n, T);
  PHASE_END();
  uint32_t err = 0;
  TRY(read_scalars(c, nullptr, &err));
  if (err & kFlagHashFull) return fail(c, KMC_E_CAPACITY, "kmc_finish: the merge table filled up (internal sizing error)");
  unsigned long long sc[3];
  TRY(d2h_small(c, sc, c->hash_scalars.p, sizeof sc));
  const uint64_t d = sc[0], n_total = sc[1], n_ones = sc[2];
  if (n_ones > 0xFFFFFFFFull) return fail(c, KMC_E_COUNT_OVERFLOW, "a k-mer occurs more than 2^32-1 times");
  PHASE_BEGIN("hash_sort");
  TRY(ensure(c, c->t_lo, (d + 2) * 8));
  TRY(ensure(c, c->keys_b, (d + 2) * 8));
  TRY(ensure(c, c->t_cnt, (d + 2) * 4));
  uint64_t *dense = (uint64_t *)c->t_lo.p, *scratch = (uint64_t *)c->keys_b.p, *sorted = nullptr;
  if (d) {
    CK(cudaMemsetAsync(d_cursor(c), 0, 8, c->stream));
    LAUNCH(hash_compact_kernel, std::min<uint32_t>(grid_for(T.mask + 1, 256), c->n_sms * 16), 256, 0, T, dense, d_cursor(c));
    TRY(radix_sort<uint64_t>(c, dense, scratch, d, c->key_bits, &sorted));
    if (sorted != dense) CK(cudaMemcpyAsync(dense, sorted, d * 8, cudaMemcpyDeviceToDevice, c->stream));
    LAUNCH(hash_lookup_kernel, std::min<uint32_t>(grid_for(d, 256), c->n_sms * 16), 256, 0, T, (const uint64_t *)dense, d,
           (uint32_t *)c->t_cnt.p);
  }
  uint64_t rows = d;
  if (n_ones) { // the all-ones key (k = 32) sorts last
    uint64_t k1 = kHashEmpty;
    uint32_t c1 = (uint32_t)n_ones;
    CK(cudaMemcpyAsync(dense + d, &k1, 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync((uint32_t *)c->t_cnt.p + d, &c1, 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    rows++;
  }
  PHASE_END();
  c->n_total = n_total; c->n_distinct = rows;
  c->strategy_used = KMC_STRATEGY_HASH;
  return KMC_OK;
}

// AUTO: is the number of distinct keys small enough for an L2-resident table?  Insert a sample into a small table.
int hash_probe(kmc_ctx *c, bool *low_cardinality) {
  *low_cardinality = false;
  c->probe_distinct = 0;
  KeyArrays ka;
  ka.from_array = !c->ingested.empty();
  if (c->cfg.mode == KMC_MODE_LR_GAPPED) return KMC_OK; // keys would have to be materialised first: skip the probe
  if (ka.from_array) for (auto &e : c->ingested) if (e.second) { ka.arrays.emplace_back(e.first, e.second); ka.n += e.second; }
  const uint64_t n_in = ka.from_array ? ka.n : c->total_bases;
  if (n_in < (1u << 18)) return KMC_OK;
  uint32_t step = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(kHistSampleMax, n_in >> 26));
  if (!ka.from_array) for (size_t i = 0; i < c->n_segs; i++) if (c->segs[i].wait_ready && c->segs[i].host_alias && step > 1) { step = std::min<uint32_t>(64, step * 4); break; }
  bool ok = false;
  HashTable T;
  c->n_hot = 0;
  PHASE_BEGIN("hash_probe");
  TRY(hash_run(c, ka, 23, 1ull << 21, step, &ok, &T));
  if (ok) {
    // keys that make up more than 1/50000 of the sampled occurrences get private shared-memory counters later
    unsigned long long sc[3];
    TRY(d2h_small(c, sc, c->hash_scalars.p, sizeof sc));
    const uint32_t thr = (uint32_t)std::max<uint64_t>(64, sc[1] / 50000);
    unsigned int *d_nhot = (unsigned int *)((unsigned char *)c->hash_hot.p + kHotMax * 8);
    CK(cudaMemsetAsync(d_nhot, 0, 4, c->stream));
    LAUNCH(hash_hot_kernel, c->n_sms * 8, 256, 0, T, thr, (uint64_t *)c->hash_hot.p, d_nhot);
    unsigned int nh = 0;
    TRY(d2h_small(c, &nh, d_nhot, 4));
    c->n_hot = std::min<uint32_t>(nh, kHotMax);
    c->probe_distinct = sc[0];
  }
  PHASE_END();
  *low_cardinality = ok;
  return KMC_OK;
}
