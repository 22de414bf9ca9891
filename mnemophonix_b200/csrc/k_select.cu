// k_select.cu — the final ranking of search() on the GPU, exact (search.c:65-107 comparator, :170 qsort, :173-193
// selection over the first ten records of the sorted array).
//
// The reference sorts ALL n_entries (entry, score, n_matches) records with libc qsort and a comparator that is not a
// strict weak order (integer abs() of a float difference), so "the first ten" depends on glibc's merge tree: a
// top-down merge sort that splits n into n/2 | n - n/2 and takes the left head when cmp <= 0.  Two facts make that
// cheap to reproduce without sorting anything completely:
//   * a merge emits its outputs from the heads of its inputs, so the first ten outputs of a node depend only on the
//     first ten of each child, whatever the comparator does: every node of the tree keeps a list of at most ten;
//   * an entry without matches loses to every entry with matches (averages are >= 30) and ties with its like, so it
//     never precedes a matched entry in any merge and can never be selected: such leaves are empty lists.
// One CTA ranks one query at a time: the leaves (tree depth where every node holds one or two entries) are read from
// the dense per-entry accumulators the probe kernel filled, then the levels are merged upwards in place — a node's
// list lives at the start of its own index range in a per-CTA scratch that stays in L2 — and thread 0 applies the
// thresholds to the root's list.  With a database sharded on node boundaries of this tree each shard ranks its own
// subtree the same way and emits the list of its root (kSelTop records per query) instead.
#include "common.cuh"
#include "search_internal.h"

namespace mnemo {

constexpr int kSelThreads = 256;

// node i of depth `depth` starts at leaf_lo[i]; leaf_lo[2^depth] = n_total
__global__ void __launch_bounds__(256) select_bounds_kernel(uint32_t n_total, uint32_t depth, uint32_t* __restrict__ leaf_lo) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > (1u << depth)) return;
  if (i == (1u << depth)) {
    leaf_lo[i] = n_total;
    return;
  }
  uint32_t lo = 0, n = n_total;
  for (int b = (int)depth - 1; b >= 0; --b) {   // msort splits n into n/2 | n - n/2
    const uint32_t n1 = n / 2;
    if ((i >> b) & 1u) {
      lo += n1;
      n -= n1;
    } else {
      n = n1;
    }
  }
  leaf_lo[i] = lo;
}

__global__ void __launch_bounds__(kSelThreads)
    search_select_kernel(const uint32_t* __restrict__ acc_score, const uint32_t* __restrict__ acc_n, uint32_t n_queries,
                         uint32_t n_local, uint32_t first_local, uint32_t n_total, uint32_t entry_offset, uint32_t depth,
                         const uint32_t* __restrict__ leaf_lo, SelRec* __restrict__ scratch, uint8_t* __restrict__ cnt_scratch,
                         mnemo_match_dev* __restrict__ out, SelRec* __restrict__ out_lists, uint32_t* __restrict__ out_counts) {
  const uint32_t n_leaves = 1u << depth;
  SelRec* __restrict__ S = scratch + (uint64_t)blockIdx.x * n_total;
  uint8_t* __restrict__ Cn = cnt_scratch + (uint64_t)blockIdx.x * n_leaves;
  const uint32_t tid = threadIdx.x;

  for (uint32_t q = blockIdx.x; q < n_queries; q += gridDim.x) {
    const uint32_t* __restrict__ rs = acc_score + (uint64_t)q * n_local;
    const uint32_t* __restrict__ rn = acc_n + (uint64_t)q * n_local;
    // leaves: one or two entries each; position p of the tree is accumulator column p - first_local
    for (uint32_t i = tid; i < n_leaves; i += kSelThreads) {
      const uint32_t lo = leaf_lo[i], hi = leaf_lo[i + 1];
      SelRec r[2];
      uint32_t c = 0;
      for (uint32_t p = lo; p < hi; ++p) {
        const uint32_t col = p - first_local;   // wraps for p < first_local: fails the range test
        if (col < n_local) {
          const uint32_t n = rn[col];
          if (n) {
            const float sc = (float)rs[col];
            r[c].entry = (int32_t)(entry_offset + p);
            r[c].score = sc;
            r[c].n = (int32_t)n;
            r[c].avg = __fdiv_rn(sc, (float)n);
            ++c;
          }
        }
      }
      if (c == 2 && cmp_entry_avg(r[0].avg, r[0].n, r[1].avg, r[1].n) > 0) {
        const SelRec t = r[0];
        r[0] = r[1];
        r[1] = t;
      }
      if (c > 0) S[lo] = r[0];
      if (c > 1) S[lo + 1] = r[1];
      Cn[i] = (uint8_t)c;
    }
    __syncthreads();
    for (int d = (int)depth - 1; d >= 0; --d) {
      const uint32_t nodes = 1u << d, sh = depth - d;
      for (uint32_t i = tid; i < nodes; i += kSelThreads) {
        const uint32_t a = i << sh, b = a + (1u << (sh - 1));
        const uint32_t ca = Cn[a], cb = Cn[b];
        if (cb == 0) continue;   // the left list is the node's list and already sits at the node's start
        const uint32_t lo_a = leaf_lo[a], lo_b = leaf_lo[b];
        if (ca == 0) {           // forward copy (destination below source)
          for (uint32_t k = 0; k < cb; ++k) S[lo_a + k] = S[lo_b + k];
          Cn[a] = (uint8_t)cb;
          continue;
        }
        SelRec A[kSelTop];       // outputs overwrite the left list's slots: buffer it
        for (uint32_t k = 0; k < ca; ++k) A[k] = S[lo_a + k];
        uint32_t ia = 0, ib = 0, w = 0;
        SelRec hb = S[lo_b];
        while (w < (uint32_t)kSelTop && ia < ca && ib < cb) {
          if (cmp_entry_avg(A[ia].avg, A[ia].n, hb.avg, hb.n) <= 0) {
            S[lo_a + w++] = A[ia++];
          } else {
            S[lo_a + w++] = hb;   // slot lo_a + w <= lo_b + ib: never ahead of the unread part of the right list
            if (++ib < cb) hb = S[lo_b + ib];
          }
        }
        while (w < (uint32_t)kSelTop && ia < ca) S[lo_a + w++] = A[ia++];
        while (w < (uint32_t)kSelTop && ib < cb) {
          S[lo_a + w++] = hb;
          if (++ib < cb) hb = S[lo_b + ib];
        }
        Cn[a] = (uint8_t)w;
      }
      __syncthreads();
    }
    if (tid == 0) {
      const uint32_t cnt = Cn[0];
      if (out_lists) {
        out_counts[q] = cnt;
        for (uint32_t k = 0; k < cnt; ++k) out_lists[(uint64_t)q * kSelTop + k] = S[k];
      } else {
        // search.c:173-185
        mnemo_match_dev best = {-7, 0.0f, 0};
        float best_avg = 0.0f;
        for (uint32_t k = 0; k < cnt; ++k) {
          const SelRec r = S[k];
          if ((r.n >= 10 || (r.avg >= 35.0f && r.n >= 5)) && r.avg >= 30.0f && r.avg > best_avg) {
            best_avg = r.avg;
            best.entry = r.entry;
            best.score = r.score;
            best.n_matches = r.n;
          }
        }
        out[q] = best;
      }
    }
    __syncthreads();   // S / Cn are rewritten by the next query's leaves
  }
}

// Sharded search: a shard's own last group (search.c:147-165 drops the greatest (entry, signature) candidate of a query
// signature) has to be scored after all whenever a HIGHER shard holds a candidate for that query signature, because
// then the globally greatest candidate lives there.  has_all[r][s] are the all-gathered flags of the n_shards shards.
// One thread per query signature.
__global__ void __launch_bounds__(256) search_fixup_kernel(const mnemo_lastgroup_dev* __restrict__ last,
                                                           const uint8_t* __restrict__ has_all, uint32_t n_shards,
                                                           uint32_t shard, uint32_t n_qsigs,
                                                           const uint32_t* __restrict__ query_start, uint32_t n_queries,
                                                           const uint32_t* __restrict__ sig_entry, uint32_t n_entries,
                                                           uint32_t* __restrict__ acc_score, uint32_t* __restrict__ acc_n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_qsigs) return;
  const mnemo_lastgroup_dev l = last[i];
  if (!l.has || l.hits < 2 || l.score < 30) return;   // search.c:11,16
  uint32_t higher = 0;
  for (uint32_t r = shard + 1; r < n_shards; ++r) higher |= has_all[(uint64_t)r * n_qsigs + i];
  if (!higher) return;
  const uint32_t q = find_segment<uint32_t>(query_start, n_queries, i);
  const uint64_t a = (uint64_t)q * n_entries + sig_entry[l.g];
  atomicAdd(&acc_score[a], l.score);
  atomicAdd(&acc_n[a], 1u);
}

// Top levels of the ranking tree, one thread per query: the lists of the n_nodes nodes (every shard's block of the
// all-gathered buffer holds [nodes_per_shard][n_queries][kSelTop] records followed by [nodes_per_shard][n_queries]
// counts; shard r owns nodes r * n_nodes / n_shards .. (r + 1) * n_nodes / n_shards - 1) are merged pairwise exactly as
// glibc's merge sort merges the nodes' sorted ranges, truncated to the first ten, then search.c:173-185.  The level
// buffers live in shared memory, record e of thread t at [e][t] (consecutive threads, consecutive 16 bytes: no bank
// conflicts); CTAs are one warp so that 10 000 queries spread over all SMs.
constexpr int kMergeThreads = 32;
size_t select_merge_smem(uint32_t n_nodes) { return sizeof(SelRec) * kSelTop * (n_nodes > 1 ? n_nodes / 2 : 1) * kMergeThreads; }

__global__ void __launch_bounds__(kMergeThreads) select_merge_kernel(const uint8_t* __restrict__ gathered, uint32_t n_shards,
                                                                     uint32_t n_nodes, uint32_t nodes_per_shard,
                                                                     uint32_t n_queries, mnemo_match_dev* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  SelRec* const sm = reinterpret_cast<SelRec*>(sm_raw);
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_queries) return;
  const uint64_t shard_bytes = (uint64_t)nodes_per_shard * n_queries * (sizeof(SelRec) * kSelTop + 4);
  const uint64_t lists_bytes = (uint64_t)nodes_per_shard * n_queries * sizeof(SelRec) * kSelTop;
  // list i of the current level: records sm[(i * kSelTop + k) * kMergeThreads + threadIdx.x]
  auto cur = [&](uint32_t i, uint32_t k) -> SelRec& { return sm[(i * kSelTop + k) * kMergeThreads + threadIdx.x]; };
  uint32_t ccnt[kMergeMaxNodes / 2];
  auto leaf = [&](uint32_t node, const SelRec** lst) -> uint32_t {
    uint32_t r = 0;
    while (r + 1 < n_shards && (uint64_t)(r + 1) * n_nodes / n_shards <= node) ++r;
    const uint32_t slot = node - (uint32_t)((uint64_t)r * n_nodes / n_shards);
    const uint8_t* blk = gathered + (uint64_t)r * shard_bytes;
    *lst = reinterpret_cast<const SelRec*>(blk) + ((uint64_t)slot * n_queries + q) * kSelTop;
    const uint32_t c = reinterpret_cast<const uint32_t*>(blk + lists_bytes)[(uint64_t)slot * n_queries + q];
    return c < (uint32_t)kSelTop ? c : (uint32_t)kSelTop;
  };
  mnemo_match_dev best = {-7, 0.0f, 0};
  float best_avg = 0.0f;
  auto consider = [&](const SelRec r) {   // search.c:173-185
    if ((r.n >= 10 || (r.avg >= 35.0f && r.n >= 5)) && r.avg >= 30.0f && r.avg > best_avg) {
      best_avg = r.avg;
      best.entry = r.entry;
      best.score = r.score;
      best.n_matches = r.n;
    }
  };
  if (n_nodes == 1) {
    const SelRec* root;
    const uint32_t cnt = leaf(0, &root);
    for (uint32_t k = 0; k < cnt; ++k) consider(root[k]);
    out[q] = best;
    return;
  }
  // first level: pairs of gathered lists -> shared memory
  for (uint32_t i = 0; i < n_nodes / 2; ++i) {
    const SelRec *A, *B;
    const uint32_t ca = leaf(2 * i, &A), cb = leaf(2 * i + 1, &B);
    uint32_t ia = 0, ib = 0, w = 0;
    while (w < (uint32_t)kSelTop && ia < ca && ib < cb) {
      const SelRec a = A[ia], b = B[ib];
      if (cmp_entry_avg(a.avg, a.n, b.avg, b.n) <= 0) {
        cur(i, w++) = a;
        ++ia;
      } else {
        cur(i, w++) = b;
        ++ib;
      }
    }
    while (w < (uint32_t)kSelTop && ia < ca) cur(i, w++) = A[ia++];
    while (w < (uint32_t)kSelTop && ib < cb) cur(i, w++) = B[ib++];
    ccnt[i] = w;
  }
  // further levels in place: list i <- merge(list 2i, list 2i+1); the left input is buffered because the output
  // overwrites list i (i <= 2i: list 0 is its own left input, the others lie ahead of what is still to be read)
  for (uint32_t m = n_nodes / 2; m > 1; m >>= 1)
    for (uint32_t i = 0; i < m / 2; ++i) {
      SelRec A[kSelTop];
      const uint32_t ca = ccnt[2 * i], cb = ccnt[2 * i + 1];
      for (uint32_t k = 0; k < ca; ++k) A[k] = cur(2 * i, k);
      uint32_t ia = 0, ib = 0, w = 0;
      SelRec B[kSelTop];
      for (uint32_t k = 0; k < cb; ++k) B[k] = cur(2 * i + 1, k);
      while (w < (uint32_t)kSelTop && ia < ca && ib < cb)
        cur(i, w++) = cmp_entry_avg(A[ia].avg, A[ia].n, B[ib].avg, B[ib].n) <= 0 ? A[ia++] : B[ib++];
      while (w < (uint32_t)kSelTop && ia < ca) cur(i, w++) = A[ia++];
      while (w < (uint32_t)kSelTop && ib < cb) cur(i, w++) = B[ib++];
      ccnt[i] = w;
    }
  for (uint32_t k = 0; k < ccnt[0]; ++k) consider(cur(0, k));
  out[q] = best;
}

__global__ void __launch_bounds__(256) search_has_kernel(const mnemo_lastgroup_dev* __restrict__ last, uint32_t n_qsigs,
                                                         uint8_t* __restrict__ has) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_qsigs) has[i] = last[i].has ? 1 : 0;
}

void launch_search_fixup(const mnemo_lastgroup_dev* last, const uint8_t* has_all, uint32_t n_shards, uint32_t shard,
                         uint32_t n_qsigs, const uint32_t* query_start, uint32_t n_queries, const uint32_t* sig_entry,
                         uint32_t n_entries, uint32_t* acc_score, uint32_t* acc_n, cudaStream_t st) {
  if (n_qsigs)
    search_fixup_kernel<<<(n_qsigs + 255) / 256, 256, 0, st>>>(last, has_all, n_shards, shard, n_qsigs, query_start, n_queries,
                                                               sig_entry, n_entries, acc_score, acc_n);
}
void launch_select_merge(const void* gathered, uint32_t n_shards, uint32_t n_nodes, uint32_t nodes_per_shard,
                         uint32_t n_queries, mnemo_match_dev* out, cudaStream_t st) {
  if (n_queries)
    select_merge_kernel<<<(n_queries + kMergeThreads - 1) / kMergeThreads, kMergeThreads, select_merge_smem(n_nodes), st>>>(
        reinterpret_cast<const uint8_t*>(gathered), n_shards, n_nodes, nodes_per_shard, n_queries, out);
}
void launch_search_has(const mnemo_lastgroup_dev* last, uint32_t n_qsigs, uint8_t* has, cudaStream_t st) {
  if (n_qsigs) search_has_kernel<<<(n_qsigs + 255) / 256, 256, 0, st>>>(last, n_qsigs, has);
}

uint32_t select_depth(uint32_t n_total) {   // depth at which every node holds one or two entries
  uint32_t d = 0;
  while (((uint64_t)2 << d) < n_total) ++d;   // 2^(d+1) >= n_total
  return d;
}
uint32_t select_grid(uint32_t n_queries, int sm_count) { return n_queries < 3u * sm_count ? n_queries : 3u * sm_count; }

void launch_select_bounds(uint32_t n_total, uint32_t depth, uint32_t* leaf_lo, cudaStream_t st) {
  const uint32_t n = (1u << depth) + 1;
  select_bounds_kernel<<<(n + 255) / 256, 256, 0, st>>>(n_total, depth, leaf_lo);
}

void launch_search_select(const uint32_t* acc_score, const uint32_t* acc_n, uint32_t n_queries, uint32_t n_local,
                          uint32_t first_local, uint32_t n_total, uint32_t entry_offset, uint32_t depth,
                          const uint32_t* leaf_lo, SelRec* scratch, uint8_t* cnt_scratch, uint32_t grid,
                          mnemo_match_dev* out, SelRec* out_lists, uint32_t* out_counts, cudaStream_t st) {
  if (!n_queries || !n_total) return;
  search_select_kernel<<<grid, kSelThreads, 0, st>>>(acc_score, acc_n, n_queries, n_local, first_local, n_total,
                                                     entry_offset, depth, leaf_lo, scratch, cnt_scratch, out, out_lists,
                                                     out_counts);
}

}  // namespace mnemo
