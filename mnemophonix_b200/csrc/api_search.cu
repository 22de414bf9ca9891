// api_search.cu — C ABI (include/mnemo_cuda.h): LSH database handle, batched / sharded search.
//
// One probe driver (probe_batch / probe_tile) serves every entry point: the query signatures go to the device once,
// the queries are processed in tiles whose dense per-(query, entry) accumulators fit a fixed budget, and each tile is
// handed to the entry point's own epilogue (exact ranking on the GPU, candidate records, or the ranked-shard lists).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/mnemo_cuda.h"
#include "api_internal.h"
#include "search_internal.h"

using namespace mnemo;

struct mnemo_db {
  mnemo_ctx* ctx = nullptr;
  uint64_t n_sigs = 0;
  uint32_t n_entries = 0, size = 0, entry_base = 0;
  std::vector<uint32_t> h_entry_start;
  uint8_t* d_sigs = nullptr;
  uint32_t *d_entry_start = nullptr, *d_sig_entry = nullptr, *d_start = nullptr, *d_ids = nullptr;
  double expected_list_ids = 0;            // sum over buckets of len^2 / n_sigs (see mnemo_db_create)
  uint32_t* d_dir = nullptr;               // with the pair index, memory allowing: run directory, 600 * n_sigs slots
  uint32_t* d_rows = nullptr;              // with the pair index: the signatures again, one 128-byte aligned row each
  unsigned long long* d_pairs = nullptr;   // pair index (k_search.cu: launch_pair_build), 300 * n_sigs entries, or nullptr
  // scratch kept between search calls (grown on demand): cudaMalloc/cudaFree cost more than the kernels
  struct Scratch {
    void* p = nullptr;
    uint64_t bytes = 0;
  } scratch[24];
  bool pair_wanted = false;           // the pair index pays for this database (decided at create) but is not built yet
  uint64_t pair_after = 4096;         // ... until this many query signatures have been searched; MNEMO_PAIR_LAZY_AFTER
  uint64_t served_qsigs = 0;
  uint32_t pair_cap = 4096;           // distinct candidates per round of the pair kernel; MNEMO_PAIR_CAP shrinks it (tests)
  uint32_t range_ids = search_range_bits();   // ids per pass of the probe's id-range mode; MNEMO_RANGE_IDS shrinks it (tests)
  uint64_t acc_budget = 1ull << 30;   // bytes of dense accumulators per query tile; MNEMO_ACC_BYTES shrinks it (tests)
  uint32_t dense_nq = 0, dense_nqs = 0;   // query batch held on the device between mnemo_search_dense_begin / _lists
};

enum {   // scratch slots
  kScrQ = 0, kScrQstart, kScrAccS, kScrAccN, kScrLast, kScrStats, kScrTotal, kScrQbase, kScrQcnt, kScrCands, kScrLeaf,
  kScrSel, kScrSelCnt, kScrOut, kScrHas, kScrHasAll, kScrLists, kScrBack, kScrTodo, kScrQsQuery
};

static cudaError_t scratch_get(mnemo_db* db, int slot, uint64_t bytes, void** out) {
  mnemo_db::Scratch& s = db->scratch[slot];
  if (s.bytes < bytes || !s.p) {
    if (s.p) cudaFree(s.p);
    s.p = nullptr;
    s.bytes = 0;
    const uint64_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&s.p, want);
    if (e != cudaSuccess) return e;
    s.bytes = want;
  }
  *out = s.p;
  return cudaSuccess;
}

template <typename T>
static cudaError_t dalloc(T** p, uint64_t count) {
  *p = nullptr;
  return cudaMalloc((void**)p, (count ? count : 1) * sizeof(T));
}

#define TRY(x)                      \
  do {                              \
    if (e == cudaSuccess) e = (x);  \
  } while (0)

// Pair index + run directory + 128-byte rows of a database whose CSR stands (k_search.cu: launch_pair_build).  Skipped,
// for good, when less than half of the free device memory would remain.
static cudaError_t build_pair_index(mnemo_db* db) {
  db->pair_wanted = false;
  if (db->d_pairs || !db->n_sigs || !pair_index_possible(db->size)) return cudaSuccess;
  cudaStream_t st = db->ctx->stream;
  cudaError_t e = cudaSuccess;
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  const size_t temp_bytes = lsh_sort_temp_bytes(db->n_sigs, db->size);
  const uint64_t need = db->n_sigs * 300ull * 8ull + db->n_sigs * (5ull * 25ull + 32ull) * 4ull + temp_bytes;   // index + rows + scratch
  if (need >= free_b / 2) return cudaSuccess;
  uint32_t *d_slots = nullptr, *d_vals_out = nullptr, *d_keys_in = nullptr, *d_vals_in = nullptr, *d_keys_out = nullptr;
  void* d_temp = nullptr;
  TRY(dalloc(&db->d_pairs, db->n_sigs * 300ull + 8));   // + 8: the last probe window may reach past the last entry
  TRY(dalloc(&d_slots, db->n_sigs * kTables));
  TRY(dalloc(&d_vals_out, db->n_sigs * kTables));
  TRY(dalloc(&d_keys_in, db->n_sigs * kTables));
  TRY(dalloc(&d_vals_in, db->n_sigs * kTables));
  TRY(dalloc(&d_keys_out, db->n_sigs * kTables));
  TRY(cudaMalloc(&d_temp, temp_bytes ? temp_bytes : 1));
  // run directory (one probe instead of a search per pair): as much memory again as the index; MNEMO_PAIR_DIR=0 / 1
  bool want_dir = need + db->n_sigs * 2400ull < free_b / 2;
  if (const char* env = getenv("MNEMO_PAIR_DIR")) want_dir = atoi(env) > 0;
  if (want_dir) TRY(dalloc(&db->d_dir, db->n_sigs * 600ull));
  TRY(launch_pair_build(db->d_sigs, db->n_sigs, db->size, db->d_start, db->d_ids, d_slots, d_keys_in, d_vals_in, d_keys_out, d_vals_out,
                        d_temp, db->d_pairs, db->d_dir, st));
  // the deep check reads whole signatures at random: 100-byte rows straddle three 64-byte DRAM bursts 56 % of
  // the time, 128-byte aligned rows always take two
  TRY(dalloc(&db->d_rows, db->n_sigs * 32ull));
  if (e == cudaSuccess) launch_pad_rows(db->d_sigs, db->n_sigs, db->d_rows, st);
  TRY(cudaStreamSynchronize(st));
  void* tmp[] = {d_slots, d_vals_out, d_keys_in, d_vals_in, d_keys_out, d_temp};
  for (void* q : tmp)
    if (q) cudaFree(q);
  if (e != cudaSuccess) {   // the streaming kernel goes on answering
    void** mine[] = {(void**)&db->d_pairs, (void**)&db->d_dir, (void**)&db->d_rows};
    for (void** q : mine) {
      if (*q) cudaFree(*q);
      *q = nullptr;
    }
  }
  return e;
}

extern "C" int mnemo_db_create(mnemo_ctx* ctx, const uint8_t* sigs, const uint32_t* entry_start, uint32_t n_entries,
                               uint32_t table_size, uint32_t entry_base, mnemo_db** out) {
  if (!ctx || !out || !entry_start) return -3;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  mnemo_db* db = new (std::nothrow) mnemo_db();
  if (!db) return -1;
  db->ctx = ctx;
  db->n_entries = n_entries;
  db->entry_base = entry_base;
  db->h_entry_start.assign(entry_start, entry_start + n_entries + 1);
  db->n_sigs = entry_start[n_entries];
  db->size = table_size ? table_size : (uint32_t)(db->n_sigs / 2);   // lsh.c:61
  if (const char* env = getenv("MNEMO_RANGE_IDS")) {   // test hook: many small id ranges on a small database
    const long v = atol(env);
    if (v > 0) db->range_ids = (uint32_t)std::min<long>(search_range_bits(), (v + 127) / 128 * 128);
  }
  if (const char* env = getenv("MNEMO_PAIR_CAP")) {   // test hook: id ranges, halving and widening in the pair kernel
    const long v = atol(env);
    if (v > 0) db->pair_cap = (uint32_t)std::min<long>(4096, v);
  }
  if (const char* env = getenv("MNEMO_ACC_BYTES")) {   // test hook: several query tiles on a small batch
    const long long v = atoll(env);
    if (v > 0) db->acc_budget = (uint64_t)v;
  }
  if ((uint64_t)db->size * kTables + 1 >= (1ull << 32) || db->n_sigs * kTables >= (1ull << 32)) {
    ctx->err = "mnemo_db_create: shard too large for 32-bit bucket offsets; use more shards";
    delete db;
    return -3;
  }
  const uint64_t n_slots = (uint64_t)db->size * kTables;
  uint32_t *d_counts = nullptr, *d_cursor = nullptr, *d_block_sums = nullptr;
  uint32_t *d_keys_in = nullptr, *d_vals_in = nullptr, *d_keys_out = nullptr;
  void* d_temp = nullptr;
  const size_t temp_bytes = (db->size && db->n_sigs) ? lsh_sort_temp_bytes(db->n_sigs, db->size) : 0;
  cudaError_t e = cudaSuccess;
  TRY(dalloc(&db->d_sigs, db->n_sigs * kSigLen + 4));
  TRY(dalloc(&db->d_entry_start, (uint64_t)n_entries + 1));
  TRY(dalloc(&db->d_sig_entry, db->n_sigs));
  TRY(dalloc(&db->d_start, n_slots + 1));
  TRY(dalloc(&db->d_ids, db->n_sigs * kTables + 4));   // + 4: the probe kernel stages whole 16-byte units
  TRY(dalloc(&d_counts, n_slots));
  TRY(dalloc(&d_cursor, n_slots));
  TRY(dalloc(&d_block_sums, (uint64_t)lsh_scan_blocks(n_slots) + 1));
  TRY(dalloc(&d_keys_in, db->n_sigs * kTables));
  TRY(dalloc(&d_vals_in, db->n_sigs * kTables));
  TRY(dalloc(&d_keys_out, db->n_sigs * kTables));
  TRY(cudaMalloc(&d_temp, temp_bytes ? temp_bytes : 1));
  cudaStream_t st = ctx->stream;
  if (db->n_sigs) TRY(cudaMemcpyAsync(db->d_sigs, sigs, db->n_sigs * kSigLen, cudaMemcpyHostToDevice, st));
  TRY(cudaMemcpyAsync(db->d_entry_start, entry_start, 4ull * (n_entries + 1), cudaMemcpyHostToDevice, st));
  TRY(cudaMemsetAsync(db->d_start, 0, (n_slots + 1) * 4, st));
  if (e == cudaSuccess && db->size && db->n_sigs) {
    TRY(launch_lsh_build(db->d_sigs, db->n_sigs, db->d_entry_start, n_entries, db->size, d_counts, db->d_start, d_cursor,
                         d_block_sums, d_keys_in, d_vals_in, d_keys_out, d_temp, temp_bytes, db->d_ids, db->d_sig_entry, st));
    TRY(cudaGetLastError());
  }
  TRY(cudaStreamSynchronize(st));
  // Pair index (k_pairprobe.cu): worth its memory (2400 bytes per signature) and build time once the bucket lists a
  // query signature probes are long.  What a query that resembles the database meets is sum(len^2) / n_sigs ids per
  // table; measured on config #4's corpus cut to 380 k / 760 k / 1.52 M / 3.04 M signatures (L = 5.3 k / 10.6 k / 21 k /
  // 42 k ids over the 25 tables; tools/crossover_run.sh, whole search of 10 k queries) the streaming kernel takes 15.2 /
  // 25.9 / 45.3 / 95 ms and the pair kernel with its run directory 16.2 / 19.4 / 25.2 / 38.3 ms, so the index is built from
  // L = 7.5 k on.  MNEMO_PAIR_INDEX=1 / 0 forces it on (tests run the small cases through it) / off.
  bool want_pairs = false, build_now = false;
  if (e == cudaSuccess && db->size && db->n_sigs) {
    unsigned long long sum_sq = 0, *d_sum = reinterpret_cast<unsigned long long*>(d_block_sums);   // (idle by now)
    launch_bucket_load(db->d_start, n_slots, d_sum, st);
    TRY(cudaMemcpyAsync(&sum_sq, d_sum, 8, cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
    db->expected_list_ids = (double)sum_sq / (double)db->n_sigs;
    want_pairs = db->expected_list_ids >= 7500.0;
  }
  // The index is built when the database has answered pair_after query signatures, not at create: 0.2-1.3 s and 4.8 KB
  // per signature are wasted on a process that loads a database for one query (`mnemophonix search`); the streaming
  // kernel answers the first few in the meantime (same results).  MNEMO_PAIR_INDEX=1 builds it here and now.
  if (const char* env = getenv("MNEMO_PAIR_INDEX")) {
    const int v = atoi(env);
    want_pairs = v > 0;
    build_now = v == 1;
  }
  if (const char* env = getenv("MNEMO_PAIR_LAZY_AFTER")) db->pair_after = (uint64_t)std::max<long long>(0, atoll(env));
  db->pair_wanted = want_pairs && db->n_sigs && pair_index_possible(db->size);
  cudaFree(d_counts);
  cudaFree(d_cursor);
  cudaFree(d_block_sums);
  cudaFree(d_keys_in);
  cudaFree(d_vals_in);
  cudaFree(d_keys_out);
  cudaFree(d_temp);
  if (e == cudaSuccess && db->pair_wanted && (build_now || db->pair_after == 0)) e = build_pair_index(db);
  if (e != cudaSuccess) {
    int rc = mnemo_fail(ctx, e, "mnemo_db_create");
    mnemo_db_destroy(db);
    return rc;
  }
  *out = db;
  return 0;
}

extern "C" int mnemo_db_has_pair_index(const mnemo_db* db) { return db && db->d_pairs ? 1 : 0; }

extern "C" void mnemo_db_destroy(mnemo_db* db) {
  if (!db) return;
  cudaSetDevice(db->ctx->device);
  cudaStreamSynchronize(db->ctx->stream);
  void* ptrs[] = {db->d_sigs, db->d_entry_start, db->d_sig_entry, db->d_start, db->d_ids, db->d_pairs, db->d_rows, db->d_dir};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (auto& sc : db->scratch)
    if (sc.p) cudaFree(sc.p);
  delete db;
}

extern "C" uint32_t mnemo_db_table_size(const mnemo_db* db) { return db ? db->size : 0; }

extern "C" uint32_t mnemo_db_tile_queries(const mnemo_db* db) {
  if (!db) return 0;
  const uint64_t per_query = (uint64_t)std::max<uint32_t>(db->n_entries, 1) * 8;
  return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(0xFFFFFFFFu, db->acc_budget / per_query));
}

static SearchParams base_params(const mnemo_db* db) {
  SearchParams p;
  memset(&p, 0, sizeof p);
  p.db_sigs = db->d_sigs;
  p.sig_entry = db->d_sig_entry;
  p.start = db->d_start;
  p.ids = db->d_ids;
  p.pairs = db->d_pairs;
  p.pair_fbits = (uint32_t)pair_index_fbits(db->size);
  p.pair_dir = db->d_dir;
  p.rows128 = db->d_rows;
  p.pair_cap = db->pair_cap;
  p.size = db->size;
  p.size_magic = db->size ? 0xFFFFFFFFFFFFFFFFull / db->size + 1 : 0;
  if (db->size) {   // size = 2^shift * odd; inverse of the odd part modulo 2^32 by Newton's iteration
    uint32_t odd = db->size, shift = 0;
    while ((odd & 1u) == 0) {
      odd >>= 1;
      ++shift;
    }
    uint32_t inv = odd;                                   // correct to 3 bits
    for (int i = 0; i < 5; i++) inv *= 2u - odd * inv;    // 6, 12, 24, 48 bits
    p.div_inv = inv;
    p.div_shift = shift;
    p.div_thr = 0xFFFFFFFFu / db->size;
  }
  p.n_entries = db->n_entries;
  p.n_sigs = (uint32_t)db->n_sigs;
  p.range_ids = db->range_ids;
  return p;
}

static uint32_t local_entry_of(const mnemo_db* db, uint32_t g) {
  const auto& es = db->h_entry_start;
  return (uint32_t)(std::upper_bound(es.begin(), es.end(), g) - es.begin() - 1);
}

extern "C" int64_t mnemo_db_get_matches(mnemo_db* db, const uint8_t* sig, uint32_t* pairs_out, uint64_t cap_pairs) {
  if (!db || !sig) return -3;
  mnemo_ctx* ctx = db->ctx;
  cudaSetDevice(ctx->device);
  if (db->size == 0 || db->n_sigs == 0) return 0;
  cudaStream_t st = ctx->stream;
  uint8_t* d_q = nullptr;
  uint32_t *d_out = nullptr, *d_counts = nullptr;
  std::vector<uint32_t> ids, counts(kTables);
  SearchParams p = base_params(db);
  cudaError_t e = dalloc(&d_q, 128);
  TRY(dalloc(&d_counts, kTables));
  TRY(cudaMemcpyAsync(d_q, sig, kSigLen, cudaMemcpyHostToDevice, st));
  p.query_sigs = d_q;
  // first the list lengths alone, then a gather buffer of exactly that size: nothing is truncated however long
  // the chains are
  if (e == cudaSuccess) launch_search_gather(p, nullptr, 0, d_counts, st);
  TRY(cudaMemcpyAsync(counts.data(), d_counts, 4 * kTables, cudaMemcpyDeviceToHost, st));
  TRY(cudaStreamSynchronize(st));
  uint64_t total = 0;
  if (e == cudaSuccess)
    for (uint32_t c : counts) total += c;
  if (e == cudaSuccess && total) {
    TRY(dalloc(&d_out, total));
    if (e == cudaSuccess) launch_search_gather(p, d_out, (uint32_t)total, d_counts, st);
    ids.resize(total);
    TRY(cudaMemcpyAsync(ids.data(), d_out, total * 4, cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
  }
  cudaFree(d_q);
  cudaFree(d_out);
  cudaFree(d_counts);
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_db_get_matches");
  // lsh.c:89-112 prepends while walking tables 0..24 and each chain from its head (members in
  // descending (entry, signature) order): the returned list reads table 24 ascending, ..., table 0.
  std::vector<uint32_t> off(kTables + 1, 0);
  for (int k = 0; k < kTables; k++) off[k + 1] = off[k] + counts[k];
  uint64_t w = 0;
  for (int k = kTables - 1; k >= 0; k--) {
    std::sort(ids.begin() + off[k], ids.begin() + off[k + 1]);
    for (uint32_t i = off[k]; i < off[k + 1]; i++, w++) {
      if (w >= cap_pairs) continue;
      const uint32_t le = local_entry_of(db, ids[i]);
      pairs_out[2 * w] = db->entry_base + le;
      pairs_out[2 * w + 1] = ids[i] - db->h_entry_start[le];
    }
  }
  return (int64_t)total;
}

// ------------------------------------------------------------------ the probe driver ----

struct ProbeBatch {
  uint8_t* d_q = nullptr;
  uint32_t *d_qstart = nullptr, *d_acc_s = nullptr, *d_acc_n = nullptr;
  mnemo_lastgroup_dev* d_last = nullptr;
  unsigned long long* d_stats = nullptr;
  uint32_t n_queries = 0, n_qsigs = 0, tile = 0;
  std::vector<uint32_t> h_qstart;
};

static bool db_is_empty(const mnemo_db* db) { return db->size == 0 || db->n_sigs == 0 || db->n_entries == 0; }

// Query signatures to the device, scratch for one tile of `tile` queries (0 = as many as the budget allows).
static cudaError_t probe_batch(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start, uint32_t n_queries,
                               uint32_t tile, ProbeBatch& b) {
  cudaStream_t st = db->ctx->stream;
  b.n_queries = n_queries;
  b.n_qsigs = query_start[n_queries];
  // the pair index is due once the database has seen enough query signatures to pay for it (mnemo_db_create)
  db->served_qsigs += b.n_qsigs;
  if (db->pair_wanted && !db->d_pairs && db->served_qsigs >= db->pair_after && build_pair_index(db) != cudaSuccess)
    cudaGetLastError();   // no memory for it after all: the streaming kernel goes on answering
  b.tile = std::max<uint32_t>(1, std::min(n_queries, tile ? tile : mnemo_db_tile_queries(db)));
  const uint64_t cells = (uint64_t)b.tile * db->n_entries;
  cudaError_t e = cudaSuccess;
  TRY(scratch_get(db, kScrQ, (uint64_t)b.n_qsigs * kSigLen + 4, (void**)&b.d_q));
  TRY(scratch_get(db, kScrQstart, 4ull * ((uint64_t)b.tile + 1), (void**)&b.d_qstart));
  TRY(scratch_get(db, kScrAccS, 4ull * cells, (void**)&b.d_acc_s));
  TRY(scratch_get(db, kScrAccN, 4ull * cells, (void**)&b.d_acc_n));
  TRY(scratch_get(db, kScrLast, sizeof(mnemo_lastgroup_dev) * (uint64_t)b.n_qsigs, (void**)&b.d_last));
  TRY(scratch_get(db, kScrStats, 32, (void**)&b.d_stats));
  TRY(cudaMemcpyAsync(b.d_q, query_sigs, (uint64_t)b.n_qsigs * kSigLen, cudaMemcpyHostToDevice, st));
  TRY(cudaMemsetAsync(b.d_stats, 0, 32, st));
  return e;
}

// Probe + score queries [q0, q0 + nq) into the (zeroed) accumulators [nq][n_entries]; last-group records of the
// tile's query signatures land in d_last at their batch-wide index.
static cudaError_t probe_tile(mnemo_db* db, ProbeBatch& b, const uint32_t* query_start, uint32_t q0, uint32_t nq) {
  cudaStream_t st = db->ctx->stream;
  cudaError_t e = cudaSuccess;
  // tile-local query_start (offsets relative to the tile's first signature); a pageable source is staged by the
  // driver before cudaMemcpyAsync returns, so the vector can be rewritten for the next tile
  b.h_qstart.assign(query_start + q0, query_start + q0 + nq + 1);
  const uint32_t s0 = b.h_qstart[0];
  for (auto& v : b.h_qstart) v -= s0;
  const uint32_t ns = b.h_qstart[nq];
  TRY(cudaMemcpyAsync(b.d_qstart, b.h_qstart.data(), 4ull * (nq + 1), cudaMemcpyHostToDevice, st));
  TRY(cudaMemsetAsync(b.d_acc_s, 0, (uint64_t)nq * db->n_entries * 4, st));
  TRY(cudaMemsetAsync(b.d_acc_n, 0, (uint64_t)nq * db->n_entries * 4, st));
  if (e != cudaSuccess) return e;
  SearchParams p = base_params(db);
  p.query_sigs = b.d_q + (uint64_t)s0 * kSigLen;
  p.query_start = b.d_qstart;
  uint32_t* d_qs_query = nullptr;
  TRY(scratch_get(db, kScrQsQuery, 4ull * ((uint64_t)ns + 1), (void**)&d_qs_query));
  if (e != cudaSuccess) return e;
  launch_qs_query(b.d_qstart, nq, ns, d_qs_query, st);
  p.qs_query = d_qs_query;
  p.n_queries = nq;
  p.acc_score = b.d_acc_s;
  p.acc_n = b.d_acc_n;
  p.last = b.d_last + s0;
  p.stat_nodes = b.d_stats;
  p.stat_deep = b.d_stats + 1;
  if (db->d_pairs) {
    static const bool trace = getenv("MNEMO_TRACE_PAIRS") != nullptr;   // diagnostics: pair-index entries visited
    unsigned long long* d_vis = nullptr;
    if (trace) {
      TRY(scratch_get(db, kScrBack, 128, (void**)&d_vis));
      TRY(cudaMemsetAsync(d_vis, 0, 128, st));
      if (e != cudaSuccess) return e;
    }
    p.stat_pairs = d_vis;
    launch_search_pairs(p, ns, st);
    if (trace) {
      unsigned long long v[16] = {0};
      cudaMemcpyAsync(v, d_vis, 128, cudaMemcpyDeviceToHost, st);
      cudaStreamSynchronize(st);
      fprintf(stderr,
              "[mnemo] pair index, per query signature: %.0f entries visited in %.1f runs (%.1f longer than the sectors read "
              "inline), %.1f pairs searched, %.1f sector probes; %.3f of the query signatures took id ranges; CTA cycles: %.0f "
              "per one-pass signature (thread 0: %.0f own search, %.0f waiting for the others, %.0f flat pass, %.0f deep checks), "
              "%.0f per id-range signature\n",
              (double)v[0] / ns, (double)v[1] / ns, (double)v[2] / ns, (double)v[3] / ns, (double)v[4] / ns, (double)v[5] / ns,
              (double)v[6] / std::max<double>(1.0, (double)ns - (double)v[5]),
              (double)v[8] / std::max<double>(1.0, (double)ns - (double)v[5]), (double)v[9] / std::max<double>(1.0, (double)ns - (double)v[5]),
              (double)v[10] / std::max<double>(1.0, (double)ns - (double)v[5]), (double)v[11] / std::max<double>(1.0, (double)ns - (double)v[5]),
              (double)v[7] / std::max<double>(1.0, (double)v[5]));
    }
  } else {
    launch_search(p, ns, st);
  }
  return cudaGetLastError();
}

static cudaError_t read_stats(const unsigned long long* d_stats, uint64_t* stats, cudaStream_t st) {
  if (!stats) return cudaSuccess;
  unsigned long long h[2] = {0, 0};
  cudaError_t e = cudaMemcpyAsync(h, d_stats, 16, cudaMemcpyDeviceToHost, st);
  TRY(cudaStreamSynchronize(st));
  stats[0] = h[0];
  stats[1] = h[1];
  return e;
}

// ------------------------------------------------------------------ search: one database on one GPU ----

// probe kernel -> per-(query, entry) accumulators -> exact ranking on the GPU (k_select.cu).  Only 12 bytes per query
// come back to the host, however many entries a query touches; with lists_out the first ten records of the sorted
// array (search.c:175: what `verbose` prints) come back as well, 164 bytes per query.
static int search_batch_impl(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start, uint32_t n_queries,
                             mnemo_match* out, mnemo_sel_rec* lists_out, uint32_t* counts_out, uint64_t* stats) {
  for (uint32_t q = 0; q < n_queries; q++) {
    out[q] = mnemo_match{-7, 0.0f, 0};
    if (counts_out) counts_out[q] = 0;
  }
  if (stats) stats[0] = stats[1] = 0;
  if (n_queries == 0 || query_start[n_queries] == 0 || db_is_empty(db)) return 0;
  mnemo_ctx* ctx = db->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  ProbeBatch b;
  cudaError_t e = probe_batch(db, query_sigs, query_start, n_queries, 0, b);
  const uint32_t n_total = db->entry_base + db->n_entries;   // entries below entry_base have no matches here
  const uint32_t depth = select_depth(n_total);
  const uint32_t grid = select_grid(b.tile, ctx->sm_count);
  uint32_t* d_leaf = nullptr;
  SelRec* d_sel = nullptr;
  uint8_t* d_cnt = nullptr;
  mnemo_match_dev* d_out = nullptr;
  SelRec* d_lists = nullptr;
  uint32_t* d_counts = nullptr;
  const uint64_t list_bytes = sizeof(SelRec) * kSelTop * (uint64_t)n_queries;
  TRY(scratch_get(db, kScrLeaf, 4ull * ((1ull << depth) + 1), (void**)&d_leaf));
  TRY(scratch_get(db, kScrSel, sizeof(SelRec) * (uint64_t)grid * n_total, (void**)&d_sel));
  TRY(scratch_get(db, kScrSelCnt, (uint64_t)grid << depth, (void**)&d_cnt));
  TRY(scratch_get(db, kScrOut, sizeof(mnemo_match_dev) * (uint64_t)n_queries, (void**)&d_out));
  uint32_t* d_todo = nullptr;
  TRY(scratch_get(db, kScrTodo, 4ull * (b.tile + 1), (void**)&d_todo));
  if (lists_out) {
    TRY(scratch_get(db, kScrLists, list_bytes + 4ull * n_queries, (void**)&d_lists));
    d_counts = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(d_lists) + list_bytes);
  }
  if (e == cudaSuccess) launch_select_bounds(n_total, depth, d_leaf, st);
  for (uint32_t q0 = 0; q0 < n_queries && e == cudaSuccess; q0 += b.tile) {
    const uint32_t nq = std::min(b.tile, n_queries - q0);
    TRY(probe_tile(db, b, query_start, q0, nq));
    if (e != cudaSuccess) break;
    launch_search_select(b.d_acc_s, b.d_acc_n, nq, db->n_entries, db->entry_base, n_total, 0, depth, d_leaf, d_sel, d_cnt,
                         select_grid(nq, ctx->sm_count), d_todo, lists_out ? nullptr : d_out + q0,
                         lists_out ? d_lists + (uint64_t)q0 * kSelTop : nullptr, lists_out ? d_counts + q0 : nullptr, st);
    TRY(cudaGetLastError());
  }
  static_assert(sizeof(mnemo_match_dev) == sizeof(mnemo_match), "match layouts");
  static_assert(sizeof(SelRec) == sizeof(mnemo_sel_rec), "record layouts");
  if (lists_out) {
    TRY(cudaMemcpyAsync(lists_out, d_lists, list_bytes, cudaMemcpyDeviceToHost, st));
    TRY(cudaMemcpyAsync(counts_out, d_counts, 4ull * n_queries, cudaMemcpyDeviceToHost, st));
  } else {
    TRY(cudaMemcpyAsync(out, d_out, sizeof(mnemo_match) * (uint64_t)n_queries, cudaMemcpyDeviceToHost, st));
  }
  TRY(cudaStreamSynchronize(st));
  TRY(read_stats(b.d_stats, stats, st));
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_search_batch");
  if (lists_out)   // search.c:173-185 over the list
    for (uint32_t q = 0; q < n_queries; q++) {
      float best_avg = 0.0f;
      for (uint32_t k = 0; k < counts_out[q] && k < (uint32_t)kSelTop; k++) {
        const mnemo_sel_rec& r = lists_out[(uint64_t)q * kSelTop + k];
        if ((r.n_matches >= 10 || (r.avg >= 35.0f && r.n_matches >= 5)) && r.avg >= 30.0f && r.avg > best_avg) {
          best_avg = r.avg;
          out[q] = mnemo_match{r.entry, r.score, r.n_matches};
        }
      }
    }
  return 0;
}

extern "C" int mnemo_search_batch(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start,
                                  uint32_t n_queries, mnemo_match* out, uint64_t* stats) {
  if (!db || !out || !query_start) return -3;
  return search_batch_impl(db, query_sigs, query_start, n_queries, out, nullptr, nullptr, stats);
}

extern "C" int mnemo_search_batch_top(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start, uint32_t n_queries,
                                      mnemo_match* out, mnemo_sel_rec* lists_out, uint32_t* counts_out) {
  if (!db || !out || !query_start || !lists_out || !counts_out) return -3;
  return search_batch_impl(db, query_sigs, query_start, n_queries, out, lists_out, counts_out, nullptr);
}

// Diagnostic form: the dense per-entry rows search.c:55-59 accumulates (score sum, n_matches), [n_queries][n_entries]
// in this shard's entry order.  What the parity tests and bench.py compare with the reference's scores[] array.
extern "C" int mnemo_search_scores(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start, uint32_t n_queries,
                                   uint32_t* score_out, uint32_t* n_matches_out) {
  if (!db || !query_start || !score_out || !n_matches_out) return -3;
  const uint64_t row = db->n_entries;
  memset(score_out, 0, 4ull * n_queries * row);
  memset(n_matches_out, 0, 4ull * n_queries * row);
  if (n_queries == 0 || query_start[n_queries] == 0 || db_is_empty(db)) return 0;
  mnemo_ctx* ctx = db->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  ProbeBatch b;
  cudaError_t e = probe_batch(db, query_sigs, query_start, n_queries, 0, b);
  for (uint32_t q0 = 0; q0 < n_queries && e == cudaSuccess; q0 += b.tile) {
    const uint32_t nq = std::min(b.tile, n_queries - q0);
    TRY(probe_tile(db, b, query_start, q0, nq));
    TRY(cudaMemcpyAsync(score_out + (uint64_t)q0 * row, b.d_acc_s, 4ull * nq * row, cudaMemcpyDeviceToHost, st));
    TRY(cudaMemcpyAsync(n_matches_out + (uint64_t)q0 * row, b.d_acc_n, 4ull * nq * row, cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
  }
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_search_scores");
  return 0;
}

// ------------------------------------------------------------------ record form (arbitrary shard cuts) ----

extern "C" int mnemo_search_shard(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start,
                                  uint32_t n_queries, mnemo_cand* cands, uint64_t cap_cands, uint64_t* n_cands,
                                  mnemo_lastgroup* last, uint64_t* stats) {
  if (!db || !query_start || !n_cands) return -3;
  *n_cands = 0;
  if (stats) stats[0] = stats[1] = 0;
  const uint32_t n_qsigs = query_start[n_queries];
  if (last)
    for (uint32_t i = 0; i < n_qsigs; i++) last[i] = mnemo_lastgroup{0, 0, 0, 0, 0};
  if (n_queries == 0 || n_qsigs == 0 || db_is_empty(db)) return 0;
  mnemo_ctx* ctx = db->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  ProbeBatch b;
  cudaError_t e = probe_batch(db, query_sigs, query_start, n_queries, 0, b);
  uint32_t *d_total = nullptr, *d_qbase = nullptr, *d_qcnt = nullptr;
  CandDev* d_cands = nullptr;
  uint64_t cand_cap = std::min<uint64_t>((uint64_t)b.tile * db->n_entries, 1u << 22);
  TRY(scratch_get(db, kScrTotal, 4, (void**)&d_total));
  TRY(scratch_get(db, kScrQbase, 4ull * b.tile, (void**)&d_qbase));
  TRY(scratch_get(db, kScrQcnt, 4ull * b.tile, (void**)&d_qcnt));
  TRY(scratch_get(db, kScrCands, sizeof(CandDev) * cand_cap, (void**)&d_cands));
  std::vector<CandDev> h_cands;
  std::vector<mnemo_cand> all;
  for (uint32_t q0 = 0; q0 < n_queries && e == cudaSuccess; q0 += b.tile) {
    const uint32_t nq = std::min(b.tile, n_queries - q0);
    TRY(probe_tile(db, b, query_start, q0, nq));
    uint32_t total = 0;
    for (int attempt = 0; attempt < 2 && e == cudaSuccess; attempt++) {
      TRY(cudaMemsetAsync(d_total, 0, 4, st));
      if (e == cudaSuccess)
        launch_search_compact(b.d_acc_s, b.d_acc_n, nq, db->n_entries, db->entry_base, d_cands, (uint32_t)cand_cap, d_total,
                              d_qbase, d_qcnt, st);
      TRY(cudaGetLastError());
      TRY(cudaMemcpyAsync(&total, d_total, 4, cudaMemcpyDeviceToHost, st));
      TRY(cudaStreamSynchronize(st));
      if (e != cudaSuccess || total <= cand_cap) break;
      cand_cap = total;   // more records than the buffer held: compact again into one that fits (the accumulators are intact)
      TRY(scratch_get(db, kScrCands, sizeof(CandDev) * cand_cap, (void**)&d_cands));
    }
    if (e != cudaSuccess) break;
    h_cands.resize(total);
    if (total) TRY(cudaMemcpy(h_cands.data(), d_cands, sizeof(CandDev) * total, cudaMemcpyDeviceToHost));
    // records come out grouped per query but in reservation order: sort (query, entry)
    std::sort(h_cands.begin(), h_cands.end(), [](const CandDev& x, const CandDev& y) {
      return x.query != y.query ? x.query < y.query : x.entry < y.entry;
    });
    for (const CandDev& c : h_cands) all.push_back(mnemo_cand{c.query + q0, c.entry, c.score, c.n_matches});
  }
  if (e == cudaSuccess && last) {
    std::vector<mnemo_lastgroup_dev> h_last(n_qsigs);
    TRY(cudaMemcpy(h_last.data(), b.d_last, sizeof(mnemo_lastgroup_dev) * n_qsigs, cudaMemcpyDeviceToHost));
    if (e == cudaSuccess)
      for (uint32_t i = 0; i < n_qsigs; i++) {
        last[i].has = h_last[i].has;
        if (!h_last[i].has) continue;
        const uint32_t le = local_entry_of(db, h_last[i].g);
        last[i].entry = db->entry_base + le;
        last[i].sig = h_last[i].g - db->h_entry_start[le];
        last[i].hits = h_last[i].hits;
        last[i].score = h_last[i].score;
      }
  }
  TRY(read_stats(b.d_stats, stats, st));
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_search_shard");
  *n_cands = all.size();
  if (all.size() > cap_cands) {
    ctx->err = "mnemo_search_shard: cands capacity too small (n_cands holds the number needed)";
    return -1;
  }
  if (cands && !all.empty()) memcpy(cands, all.data(), all.size() * sizeof(mnemo_cand));
  return 0;
}

// ---- final ranking on the host: search.c:65-107 comparator + the merge order of glibc's qsort ----

namespace {
struct Rec {
  int32_t entry;
  float score;
  int32_t n;
};

// search.c:65-107, restated (cmp_entry_avg, search_internal.h).  Not a strict weak order.
int cmp_entry_scores(const Rec& a, const Rec& b) {
  const float avg_a = a.n == 0 ? 0 : a.score / (float)a.n;
  const float avg_b = b.n == 0 ? 0 : b.score / (float)b.n;
  return cmp_entry_avg(avg_a, a.n, avg_b, b.n);
}

// The reference sorts ALL n_entries records with libc qsort (search.c:170).  glibc's qsort is a
// top-down merge sort (split n/2 | n - n/2, take the left element when cmp <= 0).  Entries
// without matches compare equal to one another and lose to every entry with matches (whose
// average is >= 30), so they end up after the others in index order and the merge tree over
// the full index range can be replayed on the few non-zero records alone.
void msort_sparse(std::vector<Rec>& v, size_t i0, size_t i1, uint32_t lo, uint32_t n, std::vector<Rec>& tmp) {
  if (i1 - i0 <= 1 || n <= 1) return;
  const uint32_t n1 = n / 2, mid = lo + n1;
  size_t split = i0;
  while (split < i1 && (uint32_t)v[split].entry < mid) split++;
  msort_sparse(v, i0, split, lo, n1, tmp);
  msort_sparse(v, split, i1, mid, n - n1, tmp);
  size_t a = i0, b = split, w = 0;
  tmp.resize(i1 - i0);
  while (a < split && b < i1) tmp[w++] = cmp_entry_scores(v[a], v[b]) <= 0 ? v[a++] : v[b++];
  while (a < split) tmp[w++] = v[a++];
  while (b < i1) tmp[w++] = v[b++];
  std::copy(tmp.begin(), tmp.begin() + w, v.begin() + i0);
}

// search.c:173-185 over the first ten records of the sorted array
mnemo_match select_best(std::vector<Rec>& recs, uint32_t n_entries_total) {
  std::vector<Rec> tmp;
  msort_sparse(recs, 0, recs.size(), 0, n_entries_total, tmp);
  mnemo_match best = {-7, 0.0f, 0};
  float best_avg = 0;
  for (size_t i = 0; i < recs.size() && i < 10; i++) {
    const Rec& r = recs[i];
    const float avg = r.n == 0 ? 0 : r.score / (float)r.n;
    if ((r.n >= 10 || (avg >= 35 && r.n >= 5)) && avg >= 30)
      if (avg > best_avg) {
        best_avg = avg;
        best.entry = r.entry;
        best.score = r.score;
        best.n_matches = r.n;
      }
  }
  return best;
}
}  // namespace

extern "C" int mnemo_search_finish(const mnemo_cand* cands, uint64_t n_cands, const mnemo_lastgroup* last,
                                   uint32_t n_shards, const uint32_t* query_start, uint32_t n_queries,
                                   uint32_t n_entries_total, mnemo_match* out) {
  if (!out || !query_start || (n_cands && !cands)) return -3;
  const uint32_t n_qsigs = query_start[n_queries];
  // group candidate records per query
  std::vector<std::vector<Rec>> per(n_queries);
  for (uint64_t i = 0; i < n_cands; i++) {
    if (cands[i].query >= n_queries) return -3;
    per[cands[i].query].push_back(Rec{cands[i].entry, cands[i].score, cands[i].n_matches});
  }
  // The group the reference never flushes is the greatest (entry, signature) over the WHOLE
  // database (search.c:147-165).  Every shard left out its own greatest group; all but the globally
  // greatest one have to be scored after all.  last is laid out [shard][query signature].
  if (last && n_shards > 1) {
    for (uint32_t q = 0; q < n_queries; q++)
      for (uint32_t s = query_start[q]; s < query_start[q + 1]; s++) {
        int top = -1;
        for (uint32_t r = 0; r < n_shards; r++) {
          const mnemo_lastgroup& g = last[(uint64_t)r * n_qsigs + s];
          if (!g.has) continue;
          if (top < 0) {
            top = (int)r;
            continue;
          }
          const mnemo_lastgroup& t = last[(uint64_t)top * n_qsigs + s];
          if (g.entry > t.entry || (g.entry == t.entry && g.sig > t.sig)) top = (int)r;
        }
        for (uint32_t r = 0; r < n_shards; r++) {
          const mnemo_lastgroup& g = last[(uint64_t)r * n_qsigs + s];
          if (!g.has || (int)r == top) continue;
          if (g.hits >= 2 && g.score >= 30) {
            bool found = false;
            for (Rec& rec : per[q])
              if (rec.entry == (int32_t)g.entry) {
                rec.score += (float)g.score;
                rec.n += 1;
                found = true;
                break;
              }
            if (!found) per[q].push_back(Rec{(int32_t)g.entry, (float)g.score, 1});
          }
        }
      }
  }
  for (uint32_t q = 0; q < n_queries; q++) {
    std::sort(per[q].begin(), per[q].end(), [](const Rec& a, const Rec& b) { return a.entry < b.entry; });
    out[q] = select_best(per[q], n_entries_total);
  }
  return 0;
}

// ------------------------------------------------------------------ sharded search, ranked on the GPUs ----
// The database is sharded on node boundaries of the ranking tree (mnemo_tree_cuts), so every shard can rank its own
// subtrees exactly (k_select.cu) and hand over kSelTop records per query and node instead of one record per matched
// entry.  Protocol per query tile (mnemophonix_b200/dist.py drives it; everything stays on the device and on the
// context's stream, the collectives are the caller's):
//   1. every rank: mnemo_search_dense_begin_dev  -> probe kernel into dense accumulators kept on the device;
//                                                   has[s] = this shard holds a candidate for query signature s
//   2. all-gather of the has flags (n_query_signatures bytes per rank: the path's small exchange step)
//   3. every rank: mnemo_search_dense_lists_dev  -> last-group fix-up where a higher shard holds candidates, ranking of
//                                                   each owned node: [per][nq][10] records + [per][nq] counts
//   4. all-gather of those, mnemo_select_merge_dev on every rank (top levels of the tree + thresholds), 12 bytes
//      per query to the host.
// The host-pointer forms (mnemo_search_dense_begin / _lists, mnemo_select_merge) wrap the same kernels for tests and
// for callers without a device-side collective.

extern "C" uint32_t mnemo_tree_cuts(uint32_t n_entries_total, uint32_t n_shards, uint32_t* cuts, uint32_t* node_lo,
                                    uint32_t node_cap) {
  if (n_shards == 0) return 0;
  uint32_t depth = 0;
  while ((1u << depth) < n_shards) ++depth;
  const uint32_t n_nodes = 1u << depth;
  std::vector<uint32_t> lo(n_nodes + 1);
  for (uint32_t i = 0; i < n_nodes; i++) {   // msort's split: n -> n/2 | n - n/2
    uint32_t l = 0, n = n_entries_total;
    for (int b = (int)depth - 1; b >= 0; --b) {
      const uint32_t n1 = n / 2;
      if ((i >> b) & 1u) {
        l += n1;
        n -= n1;
      } else {
        n = n1;
      }
    }
    lo[i] = l;
  }
  lo[n_nodes] = n_entries_total;
  if (node_lo)
    for (uint32_t i = 0; i <= n_nodes && i < node_cap; i++) node_lo[i] = lo[i];
  if (cuts)
    for (uint32_t r = 0; r <= n_shards; r++) cuts[r] = lo[(uint64_t)r * n_nodes / n_shards];   // contiguous node groups
  return n_nodes;
}

extern "C" int mnemo_search_dense_begin_dev(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start,
                                            uint32_t n_queries, uint8_t* d_has_out) {
  if (!db || !query_start || !d_has_out) return -3;
  mnemo_ctx* ctx = db->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const uint32_t n_qsigs = query_start[n_queries];
  db->dense_nq = n_queries;
  db->dense_nqs = n_qsigs;
  if (n_queries == 0 || n_qsigs == 0) return 0;
  cudaError_t e = cudaMemsetAsync(d_has_out, 0, n_qsigs, st);
  if (db_is_empty(db)) {
    if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_search_dense_begin");
    return 0;
  }
  if (n_queries > mnemo_db_tile_queries(db)) {
    ctx->err = "mnemo_search_dense_begin: query batch x shard entries exceeds the accumulator budget; send tiles of "
               "mnemo_db_tile_queries() queries";
    return -1;
  }
  ProbeBatch b;
  TRY(probe_batch(db, query_sigs, query_start, n_queries, n_queries, b));
  TRY(probe_tile(db, b, query_start, 0, n_queries));
  if (e == cudaSuccess) launch_search_has(b.d_last, n_qsigs, d_has_out, st);
  TRY(cudaGetLastError());
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_search_dense_begin");
  return 0;
}

extern "C" int mnemo_search_dense_lists_dev(mnemo_db* db, const uint8_t* d_has_all, uint32_t n_shards, uint32_t shard,
                                            const uint32_t* node_bounds, uint32_t n_nodes, mnemo_sel_rec* d_lists_out,
                                            uint32_t* d_counts_out) {
  if (!db || !node_bounds || !d_lists_out || !d_counts_out || n_shards == 0 || shard >= n_shards || (n_shards > 1 && !d_has_all))
    return -3;
  mnemo_ctx* ctx = db->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const uint32_t n_queries = db->dense_nq, n_qsigs = db->dense_nqs;
  if (n_queries == 0 || n_nodes == 0) return 0;
  if (node_bounds[0] != db->entry_base || node_bounds[n_nodes] != db->entry_base + db->n_entries) {
    ctx->err = "mnemo_search_dense_lists: the nodes do not tile this shard's entry range";
    return -3;
  }
  cudaError_t e = cudaMemsetAsync(d_counts_out, 0, 4ull * n_nodes * n_queries, st);
  if (n_qsigs == 0 || db_is_empty(db)) {
    if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_search_dense_lists");
    return 0;
  }
  uint8_t* d_cnt = nullptr;
  uint32_t* d_leaf = nullptr;
  SelRec* d_sel = nullptr;
  uint32_t max_n = 0;
  for (uint32_t i = 0; i < n_nodes; i++) max_n = std::max(max_n, node_bounds[i + 1] - node_bounds[i]);
  const uint32_t max_depth = select_depth(max_n ? max_n : 1);
  const uint32_t grid = select_grid(n_queries, ctx->sm_count);
  // one leaf table per node: the nodes are ranked back to back on the stream without a host round trip
  const uint64_t leaf_stride = (1ull << max_depth) + 1;
  TRY(scratch_get(db, kScrLeaf, 4ull * leaf_stride * n_nodes, (void**)&d_leaf));
  TRY(scratch_get(db, kScrSel, sizeof(SelRec) * (uint64_t)grid * (max_n ? max_n : 1), (void**)&d_sel));
  TRY(scratch_get(db, kScrSelCnt, (uint64_t)grid << max_depth, (void**)&d_cnt));
  uint32_t* d_todo = nullptr;
  TRY(scratch_get(db, kScrTodo, 4ull * (n_queries + 1), (void**)&d_todo));
  uint32_t* d_qstart = (uint32_t*)db->scratch[kScrQstart].p;
  uint32_t* d_acc_s = (uint32_t*)db->scratch[kScrAccS].p;
  uint32_t* d_acc_n = (uint32_t*)db->scratch[kScrAccN].p;
  mnemo_lastgroup_dev* d_last = (mnemo_lastgroup_dev*)db->scratch[kScrLast].p;
  if (e == cudaSuccess && shard + 1 < n_shards)
    launch_search_fixup(d_last, d_has_all, n_shards, shard, n_qsigs, d_qstart, n_queries, db->d_sig_entry, db->n_entries,
                        d_acc_s, d_acc_n, st);
  static_assert(sizeof(SelRec) == sizeof(mnemo_sel_rec), "record layouts");
  for (uint32_t i = 0; i < n_nodes && e == cudaSuccess; i++) {
    const uint32_t lo = node_bounds[i], n = node_bounds[i + 1] - lo;
    if (n == 0) continue;
    const uint32_t depth = select_depth(n);
    uint32_t* leaf = d_leaf + leaf_stride * i;
    launch_select_bounds(n, depth, leaf, st);
    // tree over the node's own entries: position p <-> accumulator column (lo - entry_base) + p, global entry lo + p
    launch_search_select(d_acc_s + (lo - db->entry_base), d_acc_n + (lo - db->entry_base), n_queries, db->n_entries, 0, n, lo,
                         depth, leaf, d_sel, d_cnt, grid, d_todo, nullptr,
                         reinterpret_cast<SelRec*>(d_lists_out) + (uint64_t)i * n_queries * kSelTop,
                         d_counts_out + (uint64_t)i * n_queries, st);
    TRY(cudaGetLastError());
  }
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_search_dense_lists");
  return 0;
}

extern "C" uint64_t mnemo_shard_record_bytes(uint32_t nodes_per_shard, uint32_t n_queries) {
  return (uint64_t)nodes_per_shard * n_queries * (sizeof(SelRec) * kSelTop + 4);
}

extern "C" int mnemo_select_merge_dev(mnemo_ctx* ctx, const void* d_gathered, uint32_t n_shards, uint32_t n_nodes,
                                      uint32_t nodes_per_shard, uint32_t n_queries, mnemo_match* out) {
  if (!ctx || !out || (n_queries && n_nodes && !d_gathered)) return -3;
  if ((n_nodes & (n_nodes - 1)) || n_nodes > (uint32_t)kMergeMaxNodes || n_shards == 0 || n_shards > n_nodes) return -3;
  for (uint32_t q = 0; q < n_queries; q++) out[q] = mnemo_match{-7, 0.0f, 0};
  if (n_queries == 0 || n_nodes == 0) return 0;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  mnemo_match_dev* d_out = nullptr;
  // stream-ordered allocation: no device-wide synchronisation per tile (cudaFree has one)
  cudaError_t e = cudaMallocAsync((void**)&d_out, sizeof(mnemo_match_dev) * (uint64_t)n_queries, st);
  if (e == cudaSuccess) launch_select_merge(d_gathered, n_shards, n_nodes, nodes_per_shard, n_queries, d_out, st);
  TRY(cudaGetLastError());
  TRY(cudaMemcpyAsync(out, d_out, sizeof(mnemo_match) * (uint64_t)n_queries, cudaMemcpyDeviceToHost, st));
  if (d_out) cudaFreeAsync(d_out, st);
  TRY(cudaStreamSynchronize(st));
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_select_merge_dev");
  return 0;
}

extern "C" int mnemo_search_dense_stats(mnemo_db* db, uint64_t* stats) {
  if (!db || !stats) return -3;
  stats[0] = stats[1] = 0;
  if (!db->scratch[kScrStats].p) return 0;
  cudaSetDevice(db->ctx->device);
  cudaError_t e = read_stats((const unsigned long long*)db->scratch[kScrStats].p, stats, db->ctx->stream);
  if (e != cudaSuccess) return mnemo_fail(db->ctx, e, "mnemo_search_dense_stats");
  return 0;
}

// Accumulator row of one query of the batch held since mnemo_search_dense_begin (after the fix-up when _lists ran).
extern "C" int mnemo_search_dense_read(mnemo_db* db, uint32_t query, uint32_t* score_out, uint32_t* n_matches_out) {
  if (!db || !score_out || !n_matches_out || query >= db->dense_nq) return -3;
  if (db->n_entries == 0) return 0;
  memset(score_out, 0, 4ull * db->n_entries);
  memset(n_matches_out, 0, 4ull * db->n_entries);
  if (db->dense_nqs == 0 || db_is_empty(db)) return 0;
  cudaSetDevice(db->ctx->device);
  cudaStream_t st = db->ctx->stream;
  const uint32_t* d_acc_s = (const uint32_t*)db->scratch[kScrAccS].p;
  const uint32_t* d_acc_n = (const uint32_t*)db->scratch[kScrAccN].p;
  cudaError_t e = cudaMemcpyAsync(score_out, d_acc_s + (uint64_t)query * db->n_entries, 4ull * db->n_entries,
                                  cudaMemcpyDeviceToHost, st);
  TRY(cudaMemcpyAsync(n_matches_out, d_acc_n + (uint64_t)query * db->n_entries, 4ull * db->n_entries, cudaMemcpyDeviceToHost, st));
  TRY(cudaStreamSynchronize(st));
  if (e != cudaSuccess) return mnemo_fail(db->ctx, e, "mnemo_search_dense_read");
  return 0;
}

// ---- host-pointer forms of the same protocol ----

extern "C" int mnemo_search_dense_begin(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start,
                                        uint32_t n_queries, uint8_t* has_out, uint64_t* stats) {
  if (!db || !query_start || !has_out) return -3;
  if (stats) stats[0] = stats[1] = 0;
  const uint32_t n_qsigs = query_start[n_queries];
  memset(has_out, 0, n_qsigs);
  db->dense_nq = n_queries;
  db->dense_nqs = n_qsigs;
  if (n_queries == 0 || n_qsigs == 0) return 0;
  mnemo_ctx* ctx = db->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  uint8_t* d_has = nullptr;
  cudaError_t e = scratch_get(db, kScrHas, n_qsigs, (void**)&d_has);
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_search_dense_begin");
  const int rc = mnemo_search_dense_begin_dev(db, query_sigs, query_start, n_queries, d_has);
  if (rc != 0) return rc;
  TRY(cudaMemcpyAsync(has_out, d_has, n_qsigs, cudaMemcpyDeviceToHost, st));
  TRY(cudaStreamSynchronize(st));
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_search_dense_begin");
  return stats && !db_is_empty(db) ? mnemo_search_dense_stats(db, stats) : 0;
}

extern "C" int mnemo_search_dense_lists(mnemo_db* db, const uint8_t* has_higher, const uint32_t* node_bounds,
                                        uint32_t n_nodes, mnemo_sel_rec* lists_out, uint32_t* counts_out) {
  if (!db || !node_bounds || !lists_out || !counts_out) return -3;
  mnemo_ctx* ctx = db->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const uint32_t n_queries = db->dense_nq, n_qsigs = db->dense_nqs;
  for (uint64_t i = 0; i < (uint64_t)n_nodes * n_queries; i++) counts_out[i] = 0;
  if (n_queries == 0 || n_nodes == 0) return 0;
  uint8_t* d_hh = nullptr;
  SelRec* d_lists = nullptr;
  const uint64_t list_bytes = sizeof(SelRec) * kSelTop * (uint64_t)n_nodes * n_queries, cnt_bytes = 4ull * n_nodes * n_queries;
  cudaError_t e = scratch_get(db, kScrLists, list_bytes + cnt_bytes, (void**)&d_lists);
  uint32_t* d_counts = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(d_lists) + list_bytes);
  // has_higher as the flags of one virtual higher shard: [this shard's row (unused)][the higher shards' OR]
  if (has_higher && n_qsigs) {
    TRY(scratch_get(db, kScrHasAll, 2ull * n_qsigs, (void**)&d_hh));
    TRY(cudaMemsetAsync(d_hh, 0, n_qsigs, st));
    TRY(cudaMemcpyAsync(d_hh + n_qsigs, has_higher, n_qsigs, cudaMemcpyHostToDevice, st));
  }
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_search_dense_lists");
  const int rc = mnemo_search_dense_lists_dev(db, d_hh, d_hh ? 2 : 1, 0, node_bounds, n_nodes,
                                              reinterpret_cast<mnemo_sel_rec*>(d_lists), d_counts);
  if (rc != 0) return rc;
  TRY(cudaMemcpyAsync(lists_out, d_lists, list_bytes, cudaMemcpyDeviceToHost, st));
  TRY(cudaMemcpyAsync(counts_out, d_counts, cnt_bytes, cudaMemcpyDeviceToHost, st));
  TRY(cudaStreamSynchronize(st));
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_search_dense_lists");
  return 0;
}

// Top levels of the ranking tree on the host: lists[node][query][kSelTop] of the 2^depth nodes (in node order) are
// merged pairwise exactly as glibc's merge sort would merge the nodes' sorted ranges, truncated to the first ten;
// then search.c:173-185.
extern "C" int mnemo_select_merge(const mnemo_sel_rec* lists, const uint32_t* counts, uint32_t n_nodes, uint32_t n_queries,
                                  mnemo_match* out) {
  if (!out || (n_queries && n_nodes && (!lists || !counts))) return -3;
  if (n_nodes & (n_nodes - 1)) return -3;
  std::vector<mnemo_sel_rec> cur, nxt;
  std::vector<uint32_t> ccnt, ncnt;
  for (uint32_t q = 0; q < n_queries; q++) {
    out[q] = mnemo_match{-7, 0.0f, 0};
    if (n_nodes == 0) continue;
    cur.assign((size_t)n_nodes * kSelTop, mnemo_sel_rec{0, 0.0f, 0, 0.0f});
    ccnt.assign(n_nodes, 0);
    for (uint32_t i = 0; i < n_nodes; i++) {
      ccnt[i] = std::min<uint32_t>(counts[(uint64_t)i * n_queries + q], kSelTop);
      for (uint32_t k = 0; k < ccnt[i]; k++) cur[(size_t)i * kSelTop + k] = lists[((uint64_t)i * n_queries + q) * kSelTop + k];
    }
    for (uint32_t m = n_nodes; m > 1; m >>= 1) {
      nxt.assign((size_t)(m / 2) * kSelTop, mnemo_sel_rec{0, 0.0f, 0, 0.0f});
      ncnt.assign(m / 2, 0);
      for (uint32_t i = 0; i < m / 2; i++) {
        const mnemo_sel_rec* A = &cur[(size_t)(2 * i) * kSelTop];
        const mnemo_sel_rec* B = &cur[(size_t)(2 * i + 1) * kSelTop];
        const uint32_t ca = ccnt[2 * i], cb = ccnt[2 * i + 1];
        uint32_t ia = 0, ib = 0, w = 0;
        mnemo_sel_rec* O = &nxt[(size_t)i * kSelTop];
        while (w < (uint32_t)kSelTop && ia < ca && ib < cb)
          O[w++] = cmp_entry_avg(A[ia].avg, A[ia].n_matches, B[ib].avg, B[ib].n_matches) <= 0 ? A[ia++] : B[ib++];
        while (w < (uint32_t)kSelTop && ia < ca) O[w++] = A[ia++];
        while (w < (uint32_t)kSelTop && ib < cb) O[w++] = B[ib++];
        ncnt[i] = w;
      }
      cur.swap(nxt);
      ccnt.swap(ncnt);
    }
    float best_avg = 0.0f;
    for (uint32_t k = 0; k < ccnt[0]; k++) {
      const mnemo_sel_rec& r = cur[k];
      if ((r.n_matches >= 10 || (r.avg >= 35.0f && r.n_matches >= 5)) && r.avg >= 30.0f && r.avg > best_avg) {
        best_avg = r.avg;
        out[q] = mnemo_match{r.entry, r.score, r.n_matches};
      }
    }
  }
  return 0;
}

// Self-test of the ranking kernel on caller-supplied per-entry (score, n_matches) rows [n_queries][n_entries]
// (tests: adversarial averages for the non-transitive comparator, against the qsort-based restatement).
extern "C" int mnemo_selftest_select(mnemo_ctx* ctx, const uint32_t* score, const uint32_t* n_matches, uint32_t n_queries,
                                     uint32_t n_entries, uint32_t entry_base, mnemo_match* out) {
  if (!ctx || !out || (n_queries && n_entries && (!score || !n_matches))) return -3;
  for (uint32_t q = 0; q < n_queries; q++) out[q] = mnemo_match{-7, 0.0f, 0};
  if (n_queries == 0 || n_entries == 0) return 0;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const uint32_t n_total = entry_base + n_entries, depth = select_depth(n_total);
  const uint32_t grid = select_grid(n_queries, ctx->sm_count);
  const uint64_t cells = (uint64_t)n_queries * n_entries;
  uint32_t *d_s = nullptr, *d_n = nullptr, *d_leaf = nullptr;
  SelRec* d_sel = nullptr;
  uint8_t* d_cnt = nullptr;
  mnemo_match_dev* d_out = nullptr;
  cudaError_t e = dalloc(&d_s, cells);
  TRY(dalloc(&d_n, cells));
  TRY(dalloc(&d_leaf, (1ull << depth) + 1));
  TRY(dalloc(&d_sel, (uint64_t)grid * n_total));
  TRY(dalloc(&d_cnt, (uint64_t)grid << depth));
  TRY(dalloc(&d_out, (uint64_t)n_queries));
  uint32_t* d_todo = nullptr;
  TRY(dalloc(&d_todo, (uint64_t)n_queries + 1));
  TRY(cudaMemcpyAsync(d_s, score, cells * 4, cudaMemcpyHostToDevice, st));
  TRY(cudaMemcpyAsync(d_n, n_matches, cells * 4, cudaMemcpyHostToDevice, st));
  if (e == cudaSuccess) {
    launch_select_bounds(n_total, depth, d_leaf, st);
    launch_search_select(d_s, d_n, n_queries, n_entries, entry_base, n_total, 0, depth, d_leaf, d_sel, d_cnt, grid, d_todo,
                         d_out, nullptr, nullptr, st);
    e = cudaGetLastError();
  }
  TRY(cudaMemcpyAsync(out, d_out, sizeof(mnemo_match) * (uint64_t)n_queries, cudaMemcpyDeviceToHost, st));
  TRY(cudaStreamSynchronize(st));
  void* ptrs[] = {d_s, d_n, d_leaf, d_sel, d_cnt, d_out, d_todo};
  for (void* q : ptrs)
    if (q) cudaFree(q);
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_selftest_select");
  return 0;
}
