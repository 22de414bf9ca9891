// api_search.cu — C ABI (include/mnemo_cuda.h): LSH database handle, batched / sharded search.
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
  // scratch kept between search calls (grown on demand): cudaMalloc/cudaFree cost more than the kernels
  struct Scratch {
    void* p = nullptr;
    uint64_t bytes = 0;
  } scratch[16];
  uint32_t range_ids = 1u << 18;   // ids per pass of the probe's id-range mode; MNEMO_RANGE_IDS shrinks it (tests)
  uint32_t dense_nq = 0, dense_nqs = 0;   // query batch held on the device between mnemo_search_dense_begin / _lists
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
    if (v > 0) db->range_ids = (uint32_t)std::min<long>(1L << 18, (v + 127) / 128 * 128);
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
#define TRY(x)                      \
  do {                              \
    if (e == cudaSuccess) e = (x);  \
  } while (0)
  TRY(dalloc(&db->d_sigs, db->n_sigs * kSigLen + 4));
  TRY(dalloc(&db->d_entry_start, (uint64_t)n_entries + 1));
  TRY(dalloc(&db->d_sig_entry, db->n_sigs));
  TRY(dalloc(&db->d_start, n_slots + 1));
  TRY(dalloc(&db->d_ids, db->n_sigs * kTables));
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
#undef TRY
  cudaFree(d_counts);
  cudaFree(d_cursor);
  cudaFree(d_block_sums);
  cudaFree(d_keys_in);
  cudaFree(d_vals_in);
  cudaFree(d_keys_out);
  cudaFree(d_temp);
  if (e != cudaSuccess) {
    int rc = mnemo_fail(ctx, e, "mnemo_db_create");
    mnemo_db_destroy(db);
    return rc;
  }
  *out = db;
  return 0;
}

extern "C" void mnemo_db_destroy(mnemo_db* db) {
  if (!db) return;
  cudaSetDevice(db->ctx->device);
  cudaStreamSynchronize(db->ctx->stream);
  void* ptrs[] = {db->d_sigs, db->d_entry_start, db->d_sig_entry, db->d_start, db->d_ids};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (auto& sc : db->scratch)
    if (sc.p) cudaFree(sc.p);
  delete db;
}

extern "C" uint32_t mnemo_db_table_size(const mnemo_db* db) { return db ? db->size : 0; }

static SearchParams base_params(const mnemo_db* db) {
  SearchParams p;
  memset(&p, 0, sizeof p);
  p.db_sigs = db->d_sigs;
  p.sig_entry = db->d_sig_entry;
  p.start = db->d_start;
  p.ids = db->d_ids;
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
  const uint32_t cap = (uint32_t)std::min<uint64_t>(db->n_sigs * kTables, 1u << 26);
  cudaError_t e = dalloc(&d_q, 128);
  if (e == cudaSuccess) e = dalloc(&d_out, cap);
  if (e == cudaSuccess) e = dalloc(&d_counts, kTables);
  std::vector<uint32_t> ids, counts(kTables);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_q, sig, kSigLen, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    SearchParams p = base_params(db);
    p.query_sigs = d_q;
    launch_search_gather(p, d_out, cap, d_counts, st);
    e = cudaMemcpyAsync(counts.data(), d_counts, 4 * kTables, cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  uint64_t total = 0;
  if (e == cudaSuccess) {
    for (uint32_t c : counts) total += c;
    ids.resize(std::min<uint64_t>(total, cap));
    if (!ids.empty()) e = cudaMemcpy(ids.data(), d_out, ids.size() * 4, cudaMemcpyDeviceToHost);
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
    if (off[k + 1] > ids.size()) continue;
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

// ------------------------------------------------------------------ search ----

// Records of all queries, in (query, entry) order, appended to `out` (no capacity limit).
static int search_shard_impl(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start, uint32_t n_queries,
                             std::vector<mnemo_cand>& out, mnemo_lastgroup* last, uint64_t* stats) {
  if (!db || !query_start) return -3;
  mnemo_ctx* ctx = db->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  if (stats) stats[0] = stats[1] = 0;
  const uint32_t n_qsigs = query_start[n_queries];
  if (last)
    for (uint32_t i = 0; i < n_qsigs; i++) last[i] = mnemo_lastgroup{0, 0, 0, 0, 0};
  if (n_queries == 0 || n_qsigs == 0 || db->size == 0 || db->n_sigs == 0 || db->n_entries == 0) return 0;

  // queries are processed in groups whose dense accumulators stay under ~1 GiB
  const uint64_t per_query = (uint64_t)db->n_entries * 8;
  const uint32_t group = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n_queries, (1ull << 30) / per_query));
  uint8_t* d_q = nullptr;
  uint32_t *d_qstart = nullptr, *d_acc_s = nullptr, *d_acc_n = nullptr, *d_total = nullptr, *d_qbase = nullptr,
           *d_qcnt = nullptr;
  mnemo_lastgroup_dev* d_last = nullptr;
  unsigned long long* d_stats = nullptr;
  CandDev* d_cands = nullptr;
  const uint32_t cand_cap = (uint32_t)std::min<uint64_t>((uint64_t)group * db->n_entries, 1u << 24);
  cudaError_t e = cudaSuccess;
#define TRY(x)                      \
  do {                              \
    if (e == cudaSuccess) e = (x);  \
  } while (0)
  TRY(scratch_get(db, 0, (uint64_t)n_qsigs * kSigLen + 4, (void**)&d_q));
  TRY(scratch_get(db, 1, 4ull * ((uint64_t)n_queries + 1), (void**)&d_qstart));
  TRY(scratch_get(db, 2, 4ull * group * db->n_entries, (void**)&d_acc_s));
  TRY(scratch_get(db, 3, 4ull * group * db->n_entries, (void**)&d_acc_n));
  TRY(scratch_get(db, 4, sizeof(mnemo_lastgroup_dev) * (uint64_t)n_qsigs, (void**)&d_last));
  TRY(scratch_get(db, 5, 16, (void**)&d_stats));
  TRY(scratch_get(db, 6, 4, (void**)&d_total));
  TRY(scratch_get(db, 7, 4ull * group, (void**)&d_qbase));
  TRY(scratch_get(db, 8, 4ull * group, (void**)&d_qcnt));
  TRY(scratch_get(db, 9, sizeof(CandDev) * (uint64_t)cand_cap, (void**)&d_cands));
  TRY(cudaMemcpyAsync(d_q, query_sigs, (uint64_t)n_qsigs * kSigLen, cudaMemcpyHostToDevice, st));
  TRY(cudaMemsetAsync(d_stats, 0, 16, st));
  std::vector<CandDev> h_cands;
  std::vector<uint32_t> h_qstart;
  int rc = 0;
  for (uint32_t q0 = 0; q0 < n_queries && e == cudaSuccess && rc == 0; q0 += group) {
    const uint32_t nq = std::min(group, n_queries - q0);
    // group-local query_start (offsets relative to the group's first signature)
    h_qstart.assign(query_start + q0, query_start + q0 + nq + 1);
    const uint32_t s0 = h_qstart[0];
    for (auto& v : h_qstart) v -= s0;
    const uint32_t ns = h_qstart[nq];
    TRY(cudaMemcpyAsync(d_qstart, h_qstart.data(), 4ull * (nq + 1), cudaMemcpyHostToDevice, st));
    TRY(cudaMemsetAsync(d_acc_s, 0, (uint64_t)nq * db->n_entries * 4, st));
    TRY(cudaMemsetAsync(d_acc_n, 0, (uint64_t)nq * db->n_entries * 4, st));
    TRY(cudaMemsetAsync(d_total, 0, 4, st));
    if (e != cudaSuccess) break;
    SearchParams p = base_params(db);
    p.query_sigs = d_q + (uint64_t)s0 * kSigLen;
    p.query_start = d_qstart;
    p.n_queries = nq;
    p.acc_score = d_acc_s;
    p.acc_n = d_acc_n;
    p.last = d_last + s0;
    p.stat_nodes = d_stats;
    p.stat_deep = d_stats + 1;
    launch_search(p, ns, st);
    launch_search_compact(d_acc_s, d_acc_n, nq, db->n_entries, db->entry_base, d_cands, cand_cap, d_total, d_qbase,
                          d_qcnt, st);
    uint32_t total = 0;
    TRY(cudaGetLastError());
    TRY(cudaMemcpyAsync(&total, d_total, 4, cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
    if (e != cudaSuccess) break;
    if (total > cand_cap) {
      ctx->err = "mnemo_search_shard: candidate buffer overflow";
      rc = -1;
      break;
    }
    h_cands.resize(total);
    if (total) TRY(cudaMemcpy(h_cands.data(), d_cands, sizeof(CandDev) * total, cudaMemcpyDeviceToHost));
    // records come out grouped per query but in reservation order: sort (query, entry)
    std::sort(h_cands.begin(), h_cands.end(), [](const CandDev& a, const CandDev& b) {
      return a.query != b.query ? a.query < b.query : a.entry < b.entry;
    });
    for (const CandDev& c : h_cands) out.push_back(mnemo_cand{c.query + q0, c.entry, c.score, c.n_matches});
  }
  if (e == cudaSuccess && rc == 0 && last) {
    std::vector<mnemo_lastgroup_dev> h_last(n_qsigs);
    TRY(cudaMemcpy(h_last.data(), d_last, sizeof(mnemo_lastgroup_dev) * n_qsigs, cudaMemcpyDeviceToHost));
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
  if (e == cudaSuccess && rc == 0 && stats) {
    unsigned long long h[2] = {0, 0};
    TRY(cudaMemcpy(h, d_stats, 16, cudaMemcpyDeviceToHost));
    stats[0] = h[0];
    stats[1] = h[1];
  }
#undef TRY
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_search_shard");
  return rc;
}

extern "C" int mnemo_search_shard(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start,
                                  uint32_t n_queries, mnemo_cand* cands, uint64_t cap_cands, uint64_t* n_cands,
                                  mnemo_lastgroup* last, uint64_t* stats) {
  if (!db || !query_start || !n_cands) return -3;
  *n_cands = 0;
  std::vector<mnemo_cand> v;
  const int rc = search_shard_impl(db, query_sigs, query_start, n_queries, v, last, stats);
  if (rc != 0) return rc;
  *n_cands = v.size();
  if (v.size() > cap_cands) {
    db->ctx->err = "mnemo_search_shard: cands capacity too small";
    return -1;
  }
  if (cands && !v.empty()) memcpy(cands, v.data(), v.size() * sizeof(mnemo_cand));
  return 0;
}

// ------------------------------------------------------------------ peer-memory merge ----
// The merging rank owns a result buffer in its HBM; the other ranks (one process per GPU) map it with
// cudaIpc and their search epilogue writes candidate and last-group records straight into it over NVLink.

struct mnemo_result {
  mnemo_ctx* ctx = nullptr;
  bool owner = false;
  void* base = nullptr;
  uint32_t cap = 0, n_shards = 0, n_qsigs = 0;
  ResultHeader* hdr = nullptr;
  CandDev* cands = nullptr;
  LastRec* last = nullptr;
};

static uint64_t result_bytes(uint32_t cap, uint32_t n_shards, uint32_t n_qsigs) {
  return sizeof(ResultHeader) + (uint64_t)cap * sizeof(CandDev) + (uint64_t)n_shards * n_qsigs * sizeof(LastRec);
}

static void result_carve(mnemo_result* r) {
  r->hdr = reinterpret_cast<ResultHeader*>(r->base);
  r->cands = reinterpret_cast<CandDev*>(r->hdr + 1);
  r->last = reinterpret_cast<LastRec*>(r->cands + r->cap);
}

extern "C" int mnemo_result_reset(mnemo_result* r) {
  if (!r || !r->owner) return -3;
  cudaSetDevice(r->ctx->device);
  cudaStream_t st = r->ctx->stream;
  const ResultHeader h = {0u, r->cap, r->n_shards, r->n_qsigs};
  cudaError_t e = cudaMemsetAsync(r->base, 0, result_bytes(r->cap, r->n_shards, r->n_qsigs), st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(r->hdr, &h, sizeof h, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return mnemo_fail(r->ctx, e, "mnemo_result_reset");
  return 0;
}

extern "C" int mnemo_result_create(mnemo_ctx* ctx, uint32_t cap_cands, uint32_t n_shards, uint32_t n_qsigs,
                                   mnemo_result** out, uint8_t ipc_handle[64]) {
  if (!ctx || !out || n_shards == 0) return -3;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  mnemo_result* r = new (std::nothrow) mnemo_result();
  if (!r) return -1;
  r->ctx = ctx;
  r->owner = true;
  r->cap = cap_cands;
  r->n_shards = n_shards;
  r->n_qsigs = n_qsigs;
  cudaError_t e = cudaMalloc(&r->base, result_bytes(cap_cands, n_shards, n_qsigs));
  if (e != cudaSuccess) {
    delete r;
    return mnemo_fail(ctx, e, "mnemo_result_create");
  }
  result_carve(r);
  if (ipc_handle) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, r->base);
    if (e != cudaSuccess) {
      cudaFree(r->base);
      delete r;
      return mnemo_fail(ctx, e, "mnemo_result_create (cudaIpcGetMemHandle)");
    }
    memcpy(ipc_handle, &h, 64);
  }
  int rc = mnemo_result_reset(r);
  if (rc != 0) {
    cudaFree(r->base);
    delete r;
    return rc;
  }
  *out = r;
  return 0;
}

extern "C" int mnemo_result_open(mnemo_ctx* ctx, const uint8_t ipc_handle[64], uint32_t cap_cands, uint32_t n_shards,
                                 uint32_t n_qsigs, mnemo_result** out) {
  if (!ctx || !out || !ipc_handle) return -3;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  mnemo_result* r = new (std::nothrow) mnemo_result();
  if (!r) return -1;
  r->ctx = ctx;
  r->owner = false;
  r->cap = cap_cands;
  r->n_shards = n_shards;
  r->n_qsigs = n_qsigs;
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, 64);
  cudaError_t e = cudaIpcOpenMemHandle(&r->base, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    delete r;
    return mnemo_fail(ctx, e, "mnemo_result_open (cudaIpcOpenMemHandle)");
  }
  result_carve(r);
  *out = r;
  return 0;
}

extern "C" void mnemo_result_destroy(mnemo_result* r) {
  if (!r) return;
  cudaSetDevice(r->ctx->device);
  cudaStreamSynchronize(r->ctx->stream);
  if (r->owner) cudaFree(r->base);
  else cudaIpcCloseMemHandle(r->base);
  delete r;
}

extern "C" int mnemo_search_shard_push(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start,
                                       uint32_t n_queries, mnemo_result* dst, uint32_t shard_rank, uint64_t* stats) {
  if (!db || !query_start || !dst || shard_rank >= dst->n_shards) return -3;
  mnemo_ctx* ctx = db->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  if (stats) stats[0] = stats[1] = 0;
  const uint32_t n_qsigs = query_start[n_queries];
  if (n_qsigs != dst->n_qsigs) {
    ctx->err = "mnemo_search_shard_push: result buffer was sized for a different query batch";
    return -3;
  }
  if (n_queries == 0 || n_qsigs == 0 || db->size == 0 || db->n_sigs == 0 || db->n_entries == 0) return 0;
  const uint64_t per_query = (uint64_t)db->n_entries * 8;
  const uint32_t group = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n_queries, (1ull << 30) / per_query));
  uint8_t* d_q = nullptr;
  uint32_t *d_qstart = nullptr, *d_acc_s = nullptr, *d_acc_n = nullptr;
  mnemo_lastgroup_dev* d_last = nullptr;
  unsigned long long* d_stats = nullptr;
  cudaError_t e = cudaSuccess;
#define TRY(x)                      \
  do {                              \
    if (e == cudaSuccess) e = (x);  \
  } while (0)
  TRY(scratch_get(db, 0, (uint64_t)n_qsigs * kSigLen + 4, (void**)&d_q));
  TRY(scratch_get(db, 1, 4ull * ((uint64_t)n_queries + 1), (void**)&d_qstart));
  TRY(scratch_get(db, 2, 4ull * group * db->n_entries, (void**)&d_acc_s));
  TRY(scratch_get(db, 3, 4ull * group * db->n_entries, (void**)&d_acc_n));
  TRY(scratch_get(db, 4, sizeof(mnemo_lastgroup_dev) * (uint64_t)n_qsigs, (void**)&d_last));
  TRY(scratch_get(db, 5, 16, (void**)&d_stats));
  TRY(cudaMemcpyAsync(d_q, query_sigs, (uint64_t)n_qsigs * kSigLen, cudaMemcpyHostToDevice, st));
  TRY(cudaMemsetAsync(d_stats, 0, 16, st));
  std::vector<uint32_t> h_qstart;
  for (uint32_t q0 = 0; q0 < n_queries && e == cudaSuccess; q0 += group) {
    const uint32_t nq = std::min(group, n_queries - q0);
    h_qstart.assign(query_start + q0, query_start + q0 + nq + 1);
    const uint32_t s0 = h_qstart[0];
    for (auto& v : h_qstart) v -= s0;
    const uint32_t ns = h_qstart[nq];
    TRY(cudaMemcpyAsync(d_qstart, h_qstart.data(), 4ull * (nq + 1), cudaMemcpyHostToDevice, st));
    TRY(cudaMemsetAsync(d_acc_s, 0, (uint64_t)nq * db->n_entries * 4, st));
    TRY(cudaMemsetAsync(d_acc_n, 0, (uint64_t)nq * db->n_entries * 4, st));
    if (e != cudaSuccess) break;
    SearchParams p = base_params(db);
    p.query_sigs = d_q + (uint64_t)s0 * kSigLen;
    p.query_start = d_qstart;
    p.n_queries = nq;
    p.acc_score = d_acc_s;
    p.acc_n = d_acc_n;
    p.last = d_last + s0;
    p.stat_nodes = d_stats;
    p.stat_deep = d_stats + 1;
    launch_search(p, ns, st);
    launch_search_push(d_acc_s, d_acc_n, nq, db->n_entries, db->entry_base, q0, dst->hdr, dst->cands, st);
    TRY(cudaGetLastError());
    if (q0 + group < n_queries) TRY(cudaStreamSynchronize(st));   // h_qstart is reused by the next group
  }
  if (e == cudaSuccess) {
    launch_search_last_push(d_last, n_qsigs, db->d_sig_entry, db->d_entry_start, db->entry_base,
                            dst->last + (uint64_t)shard_rank * n_qsigs, st);
    TRY(cudaGetLastError());
  }
  TRY(cudaStreamSynchronize(st));   // the records are in the merging rank's memory when this returns
  if (e == cudaSuccess && stats) {
    unsigned long long h[2] = {0, 0};
    TRY(cudaMemcpy(h, d_stats, 16, cudaMemcpyDeviceToHost));
    stats[0] = h[0];
    stats[1] = h[1];
  }
#undef TRY
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_search_shard_push");
  return 0;
}

// ---- final ranking: search.c:65-107 comparator + the merge order of glibc's qsort ----

namespace {
struct Rec {
  int32_t entry;
  float score;
  int32_t n;
};

// search.c:65-107, restated.  Not a strict weak order (integer abs() of a float difference).
int cmp_entry_scores(const Rec& a, const Rec& b) {
  const float avg_a = a.n == 0 ? 0 : a.score / (float)a.n;
  const float avg_b = b.n == 0 ? 0 : b.score / (float)b.n;
  const float delta = (float)abs((int)(avg_a - avg_b));
  if (delta <= 3) {
    if (delta <= 5 && a.n >= b.n + 5) return -1;
    if (b.n >= a.n + 5) return 1;
  }
  if ((double)delta < 0.5) {
    if (a.n > b.n) return -1;
    if (a.n < b.n) return 1;
  }
  if (avg_a > avg_b) return -1;
  if (avg_b > avg_a) return 1;
  return 0;
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

// Dense path used by mnemo_search_batch: probe kernel -> per-(query, entry) accumulators -> exact ranking on the GPU
// (k_select.cu).  Only 12 bytes per query come back to the host, however many entries a query touches.
static int search_batch_select(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start, uint32_t n_queries,
                               mnemo_match* out, uint64_t* stats) {
  mnemo_ctx* ctx = db->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  if (stats) stats[0] = stats[1] = 0;
  const uint32_t n_qsigs = query_start[n_queries];
  if (n_qsigs == 0 || db->size == 0 || db->n_sigs == 0 || db->n_entries == 0) return 0;
  const uint64_t per_query = (uint64_t)db->n_entries * 8;
  const uint32_t group = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n_queries, (1ull << 30) / per_query));
  const uint32_t n_total = db->entry_base + db->n_entries;   // entries below entry_base have no matches here
  const uint32_t depth = select_depth(n_total);
  const uint32_t grid = select_grid(group, ctx->sm_count);
  uint8_t* d_q = nullptr;
  uint32_t *d_qstart = nullptr, *d_acc_s = nullptr, *d_acc_n = nullptr, *d_leaf = nullptr;
  mnemo_lastgroup_dev* d_last = nullptr;
  unsigned long long* d_stats = nullptr;
  SelRec* d_sel = nullptr;
  uint8_t* d_cnt = nullptr;
  mnemo_match_dev* d_out = nullptr;
  cudaError_t e = cudaSuccess;
#define TRY(x)                      \
  do {                              \
    if (e == cudaSuccess) e = (x);  \
  } while (0)
  TRY(scratch_get(db, 0, (uint64_t)n_qsigs * kSigLen + 4, (void**)&d_q));
  TRY(scratch_get(db, 1, 4ull * ((uint64_t)n_queries + 1), (void**)&d_qstart));
  TRY(scratch_get(db, 2, 4ull * group * db->n_entries, (void**)&d_acc_s));
  TRY(scratch_get(db, 3, 4ull * group * db->n_entries, (void**)&d_acc_n));
  TRY(scratch_get(db, 4, sizeof(mnemo_lastgroup_dev) * (uint64_t)n_qsigs, (void**)&d_last));
  TRY(scratch_get(db, 5, 16, (void**)&d_stats));
  TRY(scratch_get(db, 10, 4ull * ((1ull << depth) + 1), (void**)&d_leaf));
  TRY(scratch_get(db, 11, sizeof(SelRec) * (uint64_t)grid * n_total, (void**)&d_sel));
  TRY(scratch_get(db, 12, (uint64_t)grid << depth, (void**)&d_cnt));
  TRY(scratch_get(db, 13, sizeof(mnemo_match_dev) * (uint64_t)n_queries, (void**)&d_out));
  TRY(cudaMemcpyAsync(d_q, query_sigs, (uint64_t)n_qsigs * kSigLen, cudaMemcpyHostToDevice, st));
  TRY(cudaMemsetAsync(d_stats, 0, 16, st));
  if (e == cudaSuccess) launch_select_bounds(n_total, depth, d_leaf, st);
  std::vector<uint32_t> h_qstart;
  for (uint32_t q0 = 0; q0 < n_queries && e == cudaSuccess; q0 += group) {
    const uint32_t nq = std::min(group, n_queries - q0);
    h_qstart.assign(query_start + q0, query_start + q0 + nq + 1);
    const uint32_t s0 = h_qstart[0];
    for (auto& v : h_qstart) v -= s0;
    const uint32_t ns = h_qstart[nq];
    TRY(cudaMemcpyAsync(d_qstart, h_qstart.data(), 4ull * (nq + 1), cudaMemcpyHostToDevice, st));
    TRY(cudaMemsetAsync(d_acc_s, 0, (uint64_t)nq * db->n_entries * 4, st));
    TRY(cudaMemsetAsync(d_acc_n, 0, (uint64_t)nq * db->n_entries * 4, st));
    if (e != cudaSuccess) break;
    SearchParams p = base_params(db);
    p.query_sigs = d_q + (uint64_t)s0 * kSigLen;
    p.query_start = d_qstart;
    p.n_queries = nq;
    p.acc_score = d_acc_s;
    p.acc_n = d_acc_n;
    p.last = d_last + s0;
    p.stat_nodes = d_stats;
    p.stat_deep = d_stats + 1;
    launch_search(p, ns, st);
    launch_search_select(d_acc_s, d_acc_n, nq, db->n_entries, db->entry_base, n_total, 0, depth, d_leaf, d_sel, d_cnt,
                         select_grid(nq, ctx->sm_count), d_out + q0, nullptr, nullptr, st);
    TRY(cudaGetLastError());
    if (q0 + group < n_queries) TRY(cudaStreamSynchronize(st));   // h_qstart is reused by the next group
  }
  static_assert(sizeof(mnemo_match_dev) == sizeof(mnemo_match), "match layouts");
  TRY(cudaMemcpyAsync(out, d_out, sizeof(mnemo_match) * (uint64_t)n_queries, cudaMemcpyDeviceToHost, st));
  TRY(cudaStreamSynchronize(st));
  if (e == cudaSuccess && stats) {
    unsigned long long h[2] = {0, 0};
    TRY(cudaMemcpy(h, d_stats, 16, cudaMemcpyDeviceToHost));
    stats[0] = h[0];
    stats[1] = h[1];
  }
#undef TRY
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_search_batch");
  return 0;
}

extern "C" int mnemo_search_batch(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start,
                                  uint32_t n_queries, mnemo_match* out, uint64_t* stats) {
  if (!db || !out || !query_start) return -3;
  for (uint32_t q = 0; q < n_queries; q++) out[q] = mnemo_match{-7, 0.0f, 0};
  if (n_queries == 0) return 0;
  return search_batch_select(db, query_sigs, query_start, n_queries, out, stats);
}

// ------------------------------------------------------------------ sharded search, ranked on the GPUs ----
// The database is sharded on node boundaries of the ranking tree (mnemo_tree_cuts), so every shard can rank its own
// subtrees exactly (k_select.cu) and hand over kSelTop records per query and node instead of one record per matched
// entry.  Protocol per query batch (mnemophonix_b200/dist.py drives it):
//   1. every rank: mnemo_search_dense_begin  -> probe kernel into dense accumulators kept on the device;
//                                               has[s] = this shard holds a candidate for query signature s
//   2. all-gather of the has flags (n_query_signatures bytes per rank: the path's small exchange step)
//   3. every rank: mnemo_search_dense_lists  -> last-group fix-up where a higher shard holds candidates, ranking of each
//                                               owned node, lists + counts to the host
//   4. all-gather of the lists, mnemo_select_merge on the merging rank (top levels of the tree + thresholds).

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

extern "C" int mnemo_search_dense_begin(mnemo_db* db, const uint8_t* query_sigs, const uint32_t* query_start,
                                        uint32_t n_queries, uint8_t* has_out, uint64_t* stats) {
  if (!db || !query_start || !has_out) return -3;
  mnemo_ctx* ctx = db->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  if (stats) stats[0] = stats[1] = 0;
  const uint32_t n_qsigs = query_start[n_queries];
  db->dense_nq = n_queries;
  db->dense_nqs = n_qsigs;
  memset(has_out, 0, n_qsigs);
  if (n_queries == 0 || n_qsigs == 0 || db->size == 0 || db->n_sigs == 0 || db->n_entries == 0) return 0;
  const uint64_t cells = (uint64_t)n_queries * db->n_entries;
  if (cells * 8 > (8ull << 30)) {
    ctx->err = "mnemo_search_dense_begin: query batch x shard entries exceeds 8 GiB of accumulators; send smaller batches";
    return -1;
  }
  uint8_t *d_q = nullptr, *d_has = nullptr;
  uint32_t *d_qstart = nullptr, *d_acc_s = nullptr, *d_acc_n = nullptr;
  mnemo_lastgroup_dev* d_last = nullptr;
  unsigned long long* d_stats = nullptr;
  cudaError_t e = cudaSuccess;
#define TRY(x)                      \
  do {                              \
    if (e == cudaSuccess) e = (x);  \
  } while (0)
  TRY(scratch_get(db, 0, (uint64_t)n_qsigs * kSigLen + 4, (void**)&d_q));
  TRY(scratch_get(db, 1, 4ull * ((uint64_t)n_queries + 1), (void**)&d_qstart));
  TRY(scratch_get(db, 2, 4ull * cells, (void**)&d_acc_s));
  TRY(scratch_get(db, 3, 4ull * cells, (void**)&d_acc_n));
  TRY(scratch_get(db, 4, sizeof(mnemo_lastgroup_dev) * (uint64_t)n_qsigs, (void**)&d_last));
  TRY(scratch_get(db, 5, 16, (void**)&d_stats));
  TRY(scratch_get(db, 14, n_qsigs, (void**)&d_has));
  TRY(cudaMemcpyAsync(d_q, query_sigs, (uint64_t)n_qsigs * kSigLen, cudaMemcpyHostToDevice, st));
  TRY(cudaMemcpyAsync(d_qstart, query_start, 4ull * (n_queries + 1), cudaMemcpyHostToDevice, st));
  TRY(cudaMemsetAsync(d_stats, 0, 16, st));
  TRY(cudaMemsetAsync(d_acc_s, 0, cells * 4, st));
  TRY(cudaMemsetAsync(d_acc_n, 0, cells * 4, st));
  if (e == cudaSuccess) {
    SearchParams p = base_params(db);
    p.query_sigs = d_q;
    p.query_start = d_qstart;
    p.n_queries = n_queries;
    p.acc_score = d_acc_s;
    p.acc_n = d_acc_n;
    p.last = d_last;
    p.stat_nodes = d_stats;
    p.stat_deep = d_stats + 1;
    launch_search(p, n_qsigs, st);
    launch_search_has(d_last, n_qsigs, d_has, st);
    TRY(cudaGetLastError());
  }
  TRY(cudaMemcpyAsync(has_out, d_has, n_qsigs, cudaMemcpyDeviceToHost, st));
  TRY(cudaStreamSynchronize(st));
  if (e == cudaSuccess && stats) {
    unsigned long long h[2] = {0, 0};
    TRY(cudaMemcpy(h, d_stats, 16, cudaMemcpyDeviceToHost));
    stats[0] = h[0];
    stats[1] = h[1];
  }
#undef TRY
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_search_dense_begin");
  return 0;
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
  if (node_bounds[0] != db->entry_base || node_bounds[n_nodes] != db->entry_base + db->n_entries) {
    ctx->err = "mnemo_search_dense_lists: the nodes do not tile this shard's entry range";
    return -3;
  }
  if (n_qsigs == 0 || db->size == 0 || db->n_sigs == 0 || db->n_entries == 0) return 0;
  uint8_t *d_hh = nullptr, *d_cnt = nullptr;
  uint32_t *d_leaf = nullptr, *d_counts = nullptr;
  SelRec *d_sel = nullptr, *d_lists = nullptr;
  uint32_t max_n = 0;
  for (uint32_t i = 0; i < n_nodes; i++) max_n = std::max(max_n, node_bounds[i + 1] - node_bounds[i]);
  const uint32_t max_depth = select_depth(max_n ? max_n : 1);
  const uint32_t grid = select_grid(n_queries, ctx->sm_count);
  cudaError_t e = cudaSuccess;
#define TRY(x)                      \
  do {                              \
    if (e == cudaSuccess) e = (x);  \
  } while (0)
  TRY(scratch_get(db, 15, n_qsigs, (void**)&d_hh));
  TRY(scratch_get(db, 10, 4ull * ((1ull << max_depth) + 1), (void**)&d_leaf));
  TRY(scratch_get(db, 11, sizeof(SelRec) * (uint64_t)grid * (max_n ? max_n : 1), (void**)&d_sel));
  TRY(scratch_get(db, 12, (uint64_t)grid << max_depth, (void**)&d_cnt));
  TRY(scratch_get(db, 13, (sizeof(SelRec) * kSelTop + 4) * (uint64_t)n_queries, (void**)&d_lists));
  d_counts = reinterpret_cast<uint32_t*>(d_lists + (uint64_t)n_queries * kSelTop);
  uint32_t* d_qstart = (uint32_t*)db->scratch[1].p;
  uint32_t* d_acc_s = (uint32_t*)db->scratch[2].p;
  uint32_t* d_acc_n = (uint32_t*)db->scratch[3].p;
  mnemo_lastgroup_dev* d_last = (mnemo_lastgroup_dev*)db->scratch[4].p;
  if (has_higher) {
    TRY(cudaMemcpyAsync(d_hh, has_higher, n_qsigs, cudaMemcpyHostToDevice, st));
    if (e == cudaSuccess)
      launch_search_fixup(d_last, d_hh, n_qsigs, d_qstart, n_queries, db->d_sig_entry, db->n_entries, d_acc_s, d_acc_n, st);
  }
  static_assert(sizeof(SelRec) == sizeof(mnemo_sel_rec), "record layouts");
  for (uint32_t i = 0; i < n_nodes && e == cudaSuccess; i++) {
    const uint32_t lo = node_bounds[i], n = node_bounds[i + 1] - lo;
    if (n == 0) continue;
    const uint32_t depth = select_depth(n);
    launch_select_bounds(n, depth, d_leaf, st);
    // tree over the node's own entries: position p <-> accumulator column (lo - entry_base) + p, global entry lo + p
    launch_search_select(d_acc_s + (lo - db->entry_base), d_acc_n + (lo - db->entry_base), n_queries, db->n_entries, 0, n, lo,
                         depth, d_leaf, d_sel, d_cnt, grid, nullptr, d_lists, d_counts, st);
    TRY(cudaGetLastError());
    TRY(cudaMemcpyAsync(lists_out + (uint64_t)i * n_queries * kSelTop, d_lists, sizeof(SelRec) * (uint64_t)n_queries * kSelTop,
                        cudaMemcpyDeviceToHost, st));
    TRY(cudaMemcpyAsync(counts_out + (uint64_t)i * n_queries, d_counts, 4ull * n_queries, cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));   // d_lists / d_leaf are reused by the next node
  }
#undef TRY
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
  if (e == cudaSuccess) e = dalloc(&d_n, cells);
  if (e == cudaSuccess) e = dalloc(&d_leaf, (1ull << depth) + 1);
  if (e == cudaSuccess) e = dalloc(&d_sel, (uint64_t)grid * n_total);
  if (e == cudaSuccess) e = dalloc(&d_cnt, (uint64_t)grid << depth);
  if (e == cudaSuccess) e = dalloc(&d_out, (uint64_t)n_queries);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_s, score, cells * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_n, n_matches, cells * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    launch_select_bounds(n_total, depth, d_leaf, st);
    launch_search_select(d_s, d_n, n_queries, n_entries, entry_base, n_total, 0, depth, d_leaf, d_sel, d_cnt, grid, d_out,
                         nullptr, nullptr, st);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, sizeof(mnemo_match) * (uint64_t)n_queries, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  void* ptrs[] = {d_s, d_n, d_leaf, d_sel, d_cnt, d_out};
  for (void* q : ptrs)
    if (q) cudaFree(q);
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_selftest_select");
  return 0;
}

extern "C" int mnemo_result_finish(mnemo_result* r, const uint32_t* query_start, uint32_t n_queries,
                                   uint32_t n_entries_total, mnemo_match* out) {
  if (!r || !r->owner || !query_start || !out) return -3;
  mnemo_ctx* ctx = r->ctx;
  cudaSetDevice(ctx->device);
  ResultHeader h;
  cudaError_t e = cudaMemcpy(&h, r->hdr, sizeof h, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_result_finish");
  if (h.count > r->cap) {
    char buf[128];
    snprintf(buf, sizeof buf, "mnemo_result_finish: candidate capacity exceeded (%u records, capacity %u)", h.count, r->cap);
    ctx->err = buf;
    return -1;
  }
  static_assert(sizeof(CandDev) == sizeof(mnemo_cand) && sizeof(LastRec) == sizeof(mnemo_lastgroup), "record layouts");
  std::vector<mnemo_cand> cands(h.count ? h.count : 1);
  std::vector<mnemo_lastgroup> last((size_t)r->n_shards * r->n_qsigs + 1);
  if (h.count) e = cudaMemcpy(cands.data(), r->cands, sizeof(mnemo_cand) * h.count, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && r->n_qsigs)
    e = cudaMemcpy(last.data(), r->last, sizeof(mnemo_lastgroup) * (size_t)r->n_shards * r->n_qsigs, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return mnemo_fail(ctx, e, "mnemo_result_finish");
  return mnemo_search_finish(cands.data(), h.count, last.data(), r->n_shards, query_start, n_queries, n_entries_total, out);
}
