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

// Result buffer of the merging rank (device memory, possibly mapped into other processes by cudaIpc):
// [ResultHeader][CandDev cands[cap]][LastRec last[n_shards][n_qsigs]]
struct ResultHeader {
  uint32_t count;   // candidate records reserved so far (may exceed cap: overflow is detected by the owner)
  uint32_t cap;
  uint32_t n_shards;
  uint32_t n_qsigs;
};
struct LastRec {
  uint32_t has, entry, sig, hits, score;
};

struct SearchParams {
  const uint8_t* db_sigs;        // [n_sigs][100]
  const uint32_t* sig_entry;     // [n_sigs] shard-local entry index
  const uint32_t* start;         // [25*size + 1] CSR offsets into ids
  const uint32_t* ids;           // [25*n_sigs]
  uint32_t size;                 // lsh.c:61 (global)
  uint32_t n_entries;            // entries in this shard
  const uint8_t* query_sigs;     // [n_query_sigs][100]
  const uint32_t* query_start;   // [n_queries + 1]
  uint32_t n_queries;
  uint32_t* acc_score;           // [n_queries][n_entries]
  uint32_t* acc_n;
  mnemo_lastgroup_dev* last;     // [n_query_sigs]
  unsigned long long* stat_nodes;
  unsigned long long* stat_deep;
};

void launch_lsh_build(const uint8_t* sigs, uint64_t n_sigs, const uint32_t* entry_start, uint32_t n_entries,
                      uint32_t size, uint32_t* counts, uint32_t* start, uint32_t* cursor, uint32_t* block_sums,
                      uint32_t* ids, uint32_t* sig_entry, cudaStream_t st);
void launch_search(const SearchParams& p, uint32_t n_query_sigs, cudaStream_t st);
void launch_search_compact(const uint32_t* acc_score, const uint32_t* acc_n, uint32_t n_queries, uint32_t n_entries,
                           uint32_t entry_base, CandDev* cands, uint32_t cap, uint32_t* total, uint32_t* q_base,
                           uint32_t* q_cnt, cudaStream_t st);
void launch_search_push(const uint32_t* acc_score, const uint32_t* acc_n, uint32_t n_queries, uint32_t n_entries,
                        uint32_t entry_base, uint32_t query_base, ResultHeader* hdr, CandDev* cands, cudaStream_t st);
void launch_search_last_push(const mnemo_lastgroup_dev* last, uint32_t n_qsigs, const uint32_t* sig_entry,
                             const uint32_t* entry_start, uint32_t entry_base, LastRec* dst, cudaStream_t st);
void launch_search_gather(const SearchParams& p, uint32_t* out, uint32_t cap, uint32_t* table_counts, cudaStream_t st);
uint32_t lsh_scan_blocks(uint64_t n_slots);

}  // namespace mnemo
