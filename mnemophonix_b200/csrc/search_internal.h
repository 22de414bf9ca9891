// search_internal.h — device-side parameter blocks of the search kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mnemo {

struct mnemo_lastgroup_dev {
  uint32_t has;     // the query signature had at least one candidate in this shard
  uint32_t g;       // greatest candidate id (shard-local signature index)
  uint32_t hits;    // number of buckets it shares with the query signature
  uint32_t score;   // number of equal bytes
};

struct CandDev {
  uint32_t query;
  int32_t entry;
  float score;
  int32_t n_matches;
};

// ---- final ranking (k_select.cu) ----
constexpr int kSelTop = 10;   // search.c:175: only the first ten records of the sorted array are looked at
struct SelRec {
  int32_t entry;   // global entry index
  float score;     // sum of deep-check scores (exact below 2^24)
  int32_t n;       // n_matches (> 0: entries without matches are never stored)
  float avg;       // score / (float)n, search.c:66-67
};
struct mnemo_match_dev {
  int32_t entry;
  float score;
  int32_t n_matches;
};
// compare_entry_scores (search.c:65-107) on precomputed averages.  Not a strict weak order: abs() is the integer
// one, applied to the truncated difference.
__host__ __device__ inline int cmp_entry_avg(float avg_a, int32_t n_a, float avg_b, int32_t n_b) {
  int di = (int)(avg_a - avg_b);
  if (di < 0) di = -di;
  const float delta = (float)di;
  if (delta <= 3) {
    if (delta <= 5 && n_a >= n_b + 5) return -1;
    if (n_b >= n_a + 5) return 1;
  }
  if ((double)delta < 0.5) {
    if (n_a > n_b) return -1;
    if (n_a < n_b) return 1;
  }
  if (avg_a > avg_b) return -1;
  if (avg_b > avg_a) return 1;
  return 0;
}

// cmp_entry_avg(a, b) <= 0 ("a is emitted before b by a merge") in integer form: delta is the float of an integer, so
// delta <= 3 is di <= 3 (delta <= 5 then holds) and (double)delta < 0.5 is di == 0.
__host__ __device__ inline bool entry_before(float avg_a, int32_t n_a, float avg_b, int32_t n_b) {
  int di = (int)(avg_a - avg_b);
  if (di < 0) di = -di;
  if (di <= 3) {
    if (n_a >= n_b + 5) return true;
    if (n_b >= n_a + 5) return false;
    if (di == 0 && n_a != n_b) return n_a > n_b;
  }
  return !(avg_b > avg_a);
}

// run directory of the pair index: hash of (slice, slot_i, slot_j) and its reduction to a slot of a table of any capacity
__host__ __device__ inline uint32_t pair_dir_hash(uint32_t slice, uint32_t slot_i, uint32_t slot_j) {
  unsigned long long x = ((unsigned long long)slot_i << 32 | slot_j) * 0x9E3779B97F4A7C15ull;
  x ^= (unsigned long long)(slice + 1u) * 0xC2B2AE3D27D4EB4Full;
  x ^= x >> 29;
  x *= 0xBF58476D1CE4E5B9ull;
  return (uint32_t)(x >> 32);
}
__host__ __device__ inline uint32_t pair_dir_slot(uint32_t hash, uint32_t capacity) {
  return (uint32_t)(((unsigned long long)hash * capacity) >> 32);
}

struct SearchParams {
  const uint8_t* db_sigs;        // [n_sigs][100]
  const uint32_t* sig_entry;     // [n_sigs] shard-local entry index
  const uint32_t* start;         // [25*size + 1] CSR offsets into ids
  const uint32_t* ids;           // [25*n_sigs]
  const unsigned long long* pairs;   // pair index [300*n_sigs]: ((slot_j << pair_fbits) | remaining) << 32 | id, or nullptr
  uint32_t pair_fbits;               // width of the saturated remaining-run-length field (k_search.cu: pair_remaining_kernel)
  const uint32_t* pair_dir;          // run directory [600*n_sigs] (k_search.cu: pair_dir_insert_kernel), or nullptr
  const uint32_t* rows128;           // with the pair index: signature g at words [32 g, 32 g + 25)
  uint32_t pair_cap;                 // distinct candidates the pair kernel lists per round (4096; MNEMO_PAIR_CAP shrinks it: tests)
  uint32_t size;                 // lsh.c:61 (global)
  uint64_t size_magic;           // 0xFFFFFFFFFFFFFFFF / size + 1 (fastmod, k_search.cu)
  uint32_t div_inv, div_shift, div_thr;   // divisibility by size (same_slot, k_search.cu)
  uint32_t n_entries;            // entries in this shard
  uint32_t n_sigs;               // signatures in this shard
  uint32_t range_ids;            // ids per pass of the probe's id-range mode (multiple of 128, <= search_range_bits())
  const uint8_t* query_sigs;     // [n_query_sigs][100]
  const uint32_t* query_start;   // [n_queries + 1]
  const uint32_t* qs_query;      // [n_query_sigs] query of every query signature (k_search.cu: launch_qs_query)
  uint32_t n_queries;
  uint32_t* acc_score;           // [n_queries][n_entries]
  uint32_t* acc_n;
  mnemo_lastgroup_dev* last;     // [n_query_sigs]
  unsigned long long* stat_nodes;
  unsigned long long* stat_deep;
  unsigned long long* stat_pairs;   // optional: pair-index entries visited (diagnostics)
};

size_t lsh_sort_temp_bytes(uint64_t n_sigs, uint32_t size);
cudaError_t launch_lsh_build(const uint8_t* sigs, uint64_t n_sigs, const uint32_t* entry_start, uint32_t n_entries,
                             uint32_t size, uint32_t* counts, uint32_t* start, uint32_t* scan_copy, uint32_t* block_sums,
                             uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, void* temp, size_t temp_bytes,
                             uint32_t* ids, uint32_t* sig_entry, cudaStream_t st);
uint32_t search_range_bits();   // ids per pass of the streaming kernel's id-range mode (one bitmap bit per id)
void launch_search(const SearchParams& p, uint32_t n_query_sigs, cudaStream_t st);
void launch_qs_query(const uint32_t* query_start, uint32_t n_queries, uint32_t n_query_sigs, uint32_t* qs_query, cudaStream_t st);
// pair-index path (k_pairprobe.cu): candidates from the pair index instead of the probed buckets
void launch_bucket_load(const uint32_t* start, uint64_t n_slots, unsigned long long* sum_sq, cudaStream_t st);
void launch_pad_rows(const uint8_t* sigs, uint64_t n_sigs, uint32_t* rows, cudaStream_t st);
void launch_search_pairs(const SearchParams& p, uint32_t n_query_sigs, cudaStream_t st);
bool pair_index_possible(uint32_t size);
int pair_index_fbits(uint32_t size);
cudaError_t launch_pair_build(const uint8_t* sigs, uint64_t n_sigs, uint32_t size, const uint32_t* start, const uint32_t* ids,
                              uint32_t* slots, uint32_t* ka, uint32_t* va, uint32_t* kb, uint32_t* vb, void* temp,
                              unsigned long long* pairs, uint32_t* dir, cudaStream_t st);
void launch_search_compact(const uint32_t* acc_score, const uint32_t* acc_n, uint32_t n_queries, uint32_t n_entries,
                           uint32_t entry_base, CandDev* cands, uint32_t cap, uint32_t* total, uint32_t* q_base,
                           uint32_t* q_cnt, cudaStream_t st);
void launch_search_gather(const SearchParams& p, uint32_t* out, uint32_t cap, uint32_t* table_counts, cudaStream_t st);
uint32_t lsh_scan_blocks(uint64_t n_slots);
void launch_search_fixup(const mnemo_lastgroup_dev* last, const uint8_t* has_all, uint32_t n_shards, uint32_t shard,
                         uint32_t n_qsigs, const uint32_t* query_start, uint32_t n_queries, const uint32_t* sig_entry,
                         uint32_t n_entries, uint32_t* acc_score, uint32_t* acc_n, cudaStream_t st);
constexpr int kMergeMaxNodes = 16;   // shards per database in the ranked-shard protocol (one box: <= 8 GPUs)
void launch_select_merge(const void* gathered, uint32_t n_shards, uint32_t n_nodes, uint32_t nodes_per_shard,
                         uint32_t n_queries, mnemo_match_dev* out, cudaStream_t st);
void launch_search_has(const mnemo_lastgroup_dev* last, uint32_t n_qsigs, uint8_t* has, cudaStream_t st);
uint32_t select_depth(uint32_t n_total);
uint32_t select_grid(uint32_t n_queries, int sm_count);
void launch_select_bounds(uint32_t n_total, uint32_t depth, uint32_t* leaf_lo, cudaStream_t st);
// Tree over positions [0, n_total); accumulator column of position p is p - first_local (columns outside
// [0, n_local) are entries without matches); reported entry = entry_offset + p.  out != nullptr: the selected match
// per query; out_lists != nullptr: the root's list (kSelTop records per query) and its length instead.
void launch_search_select(const uint32_t* acc_score, const uint32_t* acc_n, uint32_t n_queries, uint32_t n_local,
                          uint32_t first_local, uint32_t n_total, uint32_t entry_offset, uint32_t depth,
                          const uint32_t* leaf_lo, SelRec* scratch, uint8_t* cnt_scratch, uint32_t grid, uint32_t* todo,
                          mnemo_match_dev* out, SelRec* out_lists, uint32_t* out_counts, cudaStream_t st);

}  // namespace mnemo
