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
#include <algorithm>
#include <cstdio>
#include <cstdlib>
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
                         const uint32_t* __restrict__ todo, mnemo_match_dev* __restrict__ out, SelRec* __restrict__ out_lists,
                         uint32_t* __restrict__ out_counts) {
  const uint32_t n_leaves = 1u << depth;
  if (todo) n_queries = todo[0];   // only the queries the sparse pass left over: todo[1 + i]
  SelRec* __restrict__ S = scratch + (uint64_t)blockIdx.x * n_total;
  uint8_t* __restrict__ Cn = cnt_scratch + (uint64_t)blockIdx.x * n_leaves;
  const uint32_t tid = threadIdx.x;

  for (uint32_t qi = blockIdx.x; qi < n_queries; qi += gridDim.x) {
    const uint32_t q = todo ? todo[1 + qi] : qi;
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

// The same tree with everything in shared memory, for index ranges of up to kSmemSelMax positions (10 bytes each: the
// average, the match count and a 16-bit position per list slot): the merges are chains of dependent reads and writes,
// which cost a shared-memory round trip here instead of an L2 one.  The row is loaded once, coalesced; a node's list is
// a list of positions at the start of the node's own range; the score of the few records that leave the kernel is
// read again from the accumulator row.
constexpr int kSmemSelThreads = 512;
constexpr uint32_t kSmemSelMax = 20000;
size_t select_smem_bytes(uint32_t n_total, uint32_t depth) {
  return (size_t)n_total * 10 + ((size_t)1 << depth) + 32;
}

__global__ void __launch_bounds__(kSmemSelThreads)
    search_select_smem_kernel(const uint32_t* __restrict__ acc_score, const uint32_t* __restrict__ acc_n, uint32_t n_queries,
                              uint32_t n_local, uint32_t first_local, uint32_t n_total, uint32_t entry_offset, uint32_t depth,
                              const uint32_t* __restrict__ leaf_lo, const uint32_t* __restrict__ todo,
                              mnemo_match_dev* __restrict__ out, SelRec* __restrict__ out_lists,
                              uint32_t* __restrict__ out_counts) {
  extern __shared__ __align__(16) uint8_t sel_smem[];
  float* const avg = reinterpret_cast<float*>(sel_smem);                      // [n_total]
  uint32_t* const nn = reinterpret_cast<uint32_t*>(avg + n_total);            // [n_total]
  uint16_t* const L = reinterpret_cast<uint16_t*>(nn + n_total);              // [n_total] lists of positions
  uint8_t* const Cn = reinterpret_cast<uint8_t*>(L + n_total + (n_total & 1));   // [2^depth] list lengths
  const uint32_t n_leaves = 1u << depth;
  const uint32_t tid = threadIdx.x;
  if (todo) n_queries = todo[0];
  auto before = [&](uint32_t x, uint32_t y) { return entry_before(avg[x], (int32_t)nn[x], avg[y], (int32_t)nn[y]); };

  for (uint32_t qi = blockIdx.x; qi < n_queries; qi += gridDim.x) {
    const uint32_t q = todo ? todo[1 + qi] : qi;
    const uint32_t* __restrict__ rs = acc_score + (uint64_t)q * n_local;
    const uint32_t* __restrict__ rn = acc_n + (uint64_t)q * n_local;
    for (uint32_t p0 = 0; p0 < n_total; p0 += kSmemSelThreads * 4) {   // eight independent loads per thread in flight
      uint32_t n[4], sc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t col = p0 + u * kSmemSelThreads + tid - first_local;   // wraps below first_local: fails the range test
        const bool in = p0 + u * kSmemSelThreads + tid < n_total && col < n_local;
        n[u] = in ? __ldg(rn + col) : 0u;
        sc[u] = in ? __ldg(rs + col) : 0u;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t p = p0 + u * kSmemSelThreads + tid;
        if (p < n_total) {
          nn[p] = n[u];
          avg[p] = n[u] ? __fdiv_rn((float)sc[u], (float)n[u]) : 0.0f;
        }
      }
    }
    __syncthreads();
    for (uint32_t i = tid; i < n_leaves; i += kSmemSelThreads) {   // leaves: one or two positions each
      const uint32_t lo = __ldg(leaf_lo + i), hi = __ldg(leaf_lo + i + 1);
      uint32_t c = 0, r0 = 0, r1 = 0;
      for (uint32_t p = lo; p < hi; ++p)
        if (nn[p]) {
          if (c == 0) r0 = p;
          else r1 = p;
          ++c;
        }
      if (c == 2 && !before(r0, r1)) {
        const uint32_t t = r0;
        r0 = r1;
        r1 = t;
      }
      if (c > 0) L[lo] = (uint16_t)r0;
      if (c > 1) L[lo + 1] = (uint16_t)r1;
      Cn[i] = (uint8_t)c;
    }
    __syncthreads();
    for (int d = (int)depth - 1; d >= 0; --d) {
      const uint32_t nodes = 1u << d, sh = depth - d;
      for (uint32_t i = tid; i < nodes; i += kSmemSelThreads) {
        const uint32_t a = i << sh, b = a + (1u << (sh - 1));
        const uint32_t ca = Cn[a], cb = Cn[b];
        if (cb == 0) continue;   // the left list is the node's list and already sits at the node's start
        const uint32_t lo_a = __ldg(leaf_lo + a), lo_b = __ldg(leaf_lo + b);
        // The outputs stay in registers until the merge is over, so the left list is read in place although the
        // node's list overwrites it.  The heads' averages and counts are kept in registers: a step reloads one head.
        const uint32_t w = min((uint32_t)kSelTop, ca + cb);
        uint32_t O[kSelTop];
        uint32_t ia = 0, ib = 0;
        uint32_t ha = ca ? L[lo_a] : 0u, hb = L[lo_b];
        float ha_avg = avg[ha], hb_avg = avg[hb];
        int32_t ha_n = (int32_t)nn[ha], hb_n = (int32_t)nn[hb];
#pragma unroll
        for (int k = 0; k < kSelTop; ++k) {
          if ((uint32_t)k < w) {
            const bool take_a = ib >= cb || (ia < ca && entry_before(ha_avg, ha_n, hb_avg, hb_n));
            O[k] = take_a ? ha : hb;
            if (take_a) {
              if (++ia < ca) {
                ha = L[lo_a + ia];
                ha_avg = avg[ha];
                ha_n = (int32_t)nn[ha];
              }
            } else if (++ib < cb) {
              hb = L[lo_b + ib];
              hb_avg = avg[hb];
              hb_n = (int32_t)nn[hb];
            }
          }
        }
#pragma unroll
        for (int k = 0; k < kSelTop; ++k)
          if ((uint32_t)k < w) L[lo_a + k] = (uint16_t)O[k];
        Cn[a] = (uint8_t)w;
      }
      __syncthreads();
    }
    if (tid == 0) {
      const uint32_t cnt = Cn[0];
      mnemo_match_dev best = {-7, 0.0f, 0};
      float best_avg = 0.0f;
      if (out_lists) out_counts[q] = cnt;
      for (uint32_t k = 0; k < cnt; ++k) {
        const uint32_t p = L[k];
        SelRec r;
        r.entry = (int32_t)(entry_offset + p);
        r.score = (float)__ldg(rs + (p - first_local));
        r.n = (int32_t)nn[p];
        r.avg = avg[p];
        if (out_lists) {
          out_lists[(uint64_t)q * kSelTop + k] = r;
        } else if ((r.n >= 10 || (r.avg >= 35.0f && r.n >= 5)) && r.avg >= 30.0f && r.avg > best_avg) {   // search.c:173-185
          best_avg = r.avg;
          best.entry = r.entry;
          best.score = r.score;
          best.n_matches = r.n;
        }
      }
      if (!out_lists) out[q] = best;
    }
    __syncthreads();   // the arrays are rewritten by the next query's row
  }
}

// The usual case — a query matches far fewer entries than the database holds — does not need the dense tree.  One
// CTA per query scans the accumulator row and keeps the matched entries, in ascending position, in shared memory, each
// with its path in the merge tree (one bit per level: msort's n/2 | n - n/2 split, followed until every node holds one
// entry).  The paths are sorted, so the matched entries under any node are a contiguous run found by bisection; level
// by level, the first thread of every run merges the lists of the node's two children exactly as above — the node's
// list (at most ten records) sits at the start of its run, left first on cmp <= 0 — and nothing outside shared memory
// is touched.  Queries with more than `cap` matched entries are appended to todo[] for the dense kernel above.
constexpr int kCompactThreads = 256;
constexpr int kCompactCap = 2048;
constexpr int kCompactU = 8;   // columns per lane and chunk: a warp scans 256 consecutive columns, the CTA 2048

__global__ void __launch_bounds__(kCompactThreads)
    search_select_compact_kernel(const uint32_t* __restrict__ acc_score, const uint32_t* __restrict__ acc_n, uint32_t n_queries,
                                 uint32_t n_local, uint32_t first_local, uint32_t n_total, uint32_t entry_offset,
                                 uint32_t levels, uint32_t cap, uint32_t* __restrict__ todo, mnemo_match_dev* __restrict__ out,
                                 SelRec* __restrict__ out_lists, uint32_t* __restrict__ out_counts) {
  __shared__ SelRec R[kCompactCap];
  __shared__ uint32_t key[kCompactCap];
  __shared__ uint8_t cnt[kCompactCap];
  __shared__ uint32_t s_wcnt[2][kCompactThreads / 32];
  constexpr int kWarps = kCompactThreads / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t q = blockIdx.x;
  const uint32_t* __restrict__ rs = acc_score + (uint64_t)q * n_local;
  const uint32_t* __restrict__ rn = acc_n + (uint64_t)q * n_local;
  const uint32_t n_cols = first_local >= n_total ? 0u : min(n_local, n_total - first_local);
  uint32_t K = 0;
  int par = 0;
  for (uint32_t c0 = 0; c0 < n_cols; c0 += kCompactThreads * kCompactU, par ^= 1) {
    const uint32_t w0 = c0 + warp * (32 * kCompactU);
    uint32_t v[kCompactU];
    uint32_t mine = 0;
#pragma unroll
    for (int u = 0; u < kCompactU; ++u) {
      const uint32_t col = w0 + u * 32 + lane;
      v[u] = col < n_cols ? __ldg(rn + col) : 0u;
    }
#pragma unroll
    for (int u = 0; u < kCompactU; ++u) mine += __popc(__ballot_sync(0xffffffffu, v[u] != 0u));
    if (lane == 0) s_wcnt[par][warp] = mine;
    __syncthreads();   // one barrier per chunk: the counts alternate between two rows
    uint32_t pos = K, total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const uint32_t c = s_wcnt[par][w];
      pos += w < warp ? c : 0u;
      total += c;
    }
    K += total;
    if (K > cap) break;   // uniform: too many for this kernel, the rest of the row does not matter
    if (mine == 0) continue;
#pragma unroll
    for (int u = 0; u < kCompactU; ++u) {
      const uint32_t m = __ballot_sync(0xffffffffu, v[u] != 0u);
      const uint32_t at = pos + __popc(m & ((1u << lane) - 1u));
      if (v[u] && at < cap) {
        const uint32_t col = w0 + u * 32 + lane, p = first_local + col;
        const float sc = (float)__ldg(rs + col);
        SelRec r;
        r.entry = (int32_t)(entry_offset + p);
        r.score = sc;
        r.n = (int32_t)v[u];
        r.avg = __fdiv_rn(sc, (float)v[u]);
        R[at] = r;
        uint32_t path = 0, lo = 0, n = n_total;
        for (uint32_t l = 0; l < levels; ++l) {
          const uint32_t n1 = n >> 1;
          if (p >= lo + n1) {
            path = (path << 1) | 1u;
            lo += n1;
            n -= n1;
          } else {
            path <<= 1;
            n = n1;
          }
        }
        key[at] = path;
        cnt[at] = 1;
      }
      pos += __popc(m);
    }
  }
  if (K > cap) {   // uniform: every thread has the same K
    if (tid == 0) todo[1 + atomicAdd(&todo[0], 1u)] = q;
    return;
  }
  __syncthreads();
  for (uint32_t sh = 1; sh <= levels && K > 1; ++sh) {
    for (uint32_t a = tid; a < K; a += kCompactThreads) {
      const uint32_t ka = key[a], node = ka >> sh;
      if (a > 0 && (key[a - 1] >> sh) == node) continue;   // not the first matched entry under this node
      if ((ka >> (sh - 1)) & 1u) continue;                  // nothing under the left child: the list is in place
      // first matched entry under the right child: the left child holds at most 2^(sh-1) entries
      const uint32_t right = ((node << 1) | 1u) << (sh - 1);
      uint32_t lo = a + 1, hi = min(K, a + (1u << (sh - 1)));
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (key[mid] < right) lo = mid + 1;
        else hi = mid;
      }
      const uint32_t b = lo;
      if (b >= K || (key[b] >> sh) != node) continue;       // nothing under the right child
      const uint32_t ca = cnt[a], cb = cnt[b];
      SelRec A[kSelTop];   // outputs overwrite the left list's slots: buffer it
#pragma unroll
      for (int k = 0; k < kSelTop; ++k)
        if ((uint32_t)k < ca) A[k] = R[a + k];
      uint32_t ia = 0, ib = 0, w = 0;
      SelRec hb = R[b];
      while (w < (uint32_t)kSelTop && ia < ca && ib < cb) {
        SelRec x = A[0];
#pragma unroll
        for (int k = 1; k < kSelTop; ++k)
          if ((uint32_t)k == ia) x = A[k];
        if (cmp_entry_avg(x.avg, x.n, hb.avg, hb.n) <= 0) {
          R[a + w++] = x;
          ++ia;
        } else {
          R[a + w++] = hb;   // slot a + w <= b + ib: never ahead of the unread part of the right list
          if (++ib < cb) hb = R[b + ib];
        }
      }
      while (w < (uint32_t)kSelTop && ia < ca) {
        SelRec x = A[0];
#pragma unroll
        for (int k = 1; k < kSelTop; ++k)
          if ((uint32_t)k == ia) x = A[k];
        R[a + w++] = x;
        ++ia;
      }
      while (w < (uint32_t)kSelTop && ib < cb) {
        R[a + w++] = hb;
        if (++ib < cb) hb = R[b + ib];
      }
      cnt[a] = (uint8_t)w;
    }
    __syncthreads();
  }
  if (tid == 0) {
    const uint32_t c = K ? min((uint32_t)cnt[0], (uint32_t)kSelTop) : 0u;
    if (out_lists) {
      out_counts[q] = c;
      for (uint32_t k = 0; k < c; ++k) out_lists[(uint64_t)q * kSelTop + k] = R[k];
    } else {
      mnemo_match_dev best = {-7, 0.0f, 0};   // search.c:173-185
      float best_avg = 0.0f;
      for (uint32_t k = 0; k < c; ++k) {
        const SelRec r = R[k];
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

// todo: n_queries + 1 words of device scratch (count, then the queries left to the dense kernel)
void launch_search_select(const uint32_t* acc_score, const uint32_t* acc_n, uint32_t n_queries, uint32_t n_local,
                          uint32_t first_local, uint32_t n_total, uint32_t entry_offset, uint32_t depth,
                          const uint32_t* leaf_lo, SelRec* scratch, uint8_t* cnt_scratch, uint32_t grid, uint32_t* todo,
                          mnemo_match_dev* out, SelRec* out_lists, uint32_t* out_counts, cudaStream_t st) {
  if (!n_queries || !n_total) return;
  const uint32_t cap = [] {   // test hook (read per call): 0 sends every query with a match to the dense kernel
    const char* v = getenv("MNEMO_SELECT_CAP");
    const long c = v ? atol(v) : kCompactCap;
    return (uint32_t)(c < 0 ? 0 : c > kCompactCap ? kCompactCap : c);
  }();
  // Two stages: the compacting kernel ranks the rows with few matched entries (the usual case: 10 k queries against a
  // sparse 3 M-signature database rank 0.4 ms faster through it than through the tree kernels, 2 ms faster at 19 k
  // entries) and gives up on a row as soon as it has seen more than `cap` of them; a tree kernel ranks what is left.
  // Where the shared-memory tree fits, the first stage only keeps rows of up to 256 matched entries, so that a dense
  // row costs it a twentieth of a scan.
  const bool fits = n_total <= kSmemSelMax && getenv("MNEMO_SELECT_GLOBAL") == nullptr;
  const uint32_t cap1 = fits && getenv("MNEMO_SELECT_CAP") == nullptr ? 256u : cap;
  cudaMemsetAsync(todo, 0, 4, st);
  // levels: until every node of the tree holds one entry (one more than the dense kernel's leaves of one or two)
  search_select_compact_kernel<<<n_queries, kCompactThreads, 0, st>>>(
      acc_score, acc_n, n_queries, n_local, first_local, n_total, entry_offset, depth + 1, cap1, todo, out, out_lists,
      out_counts);
  if (fits) {
    const size_t smem = select_smem_bytes(n_total, depth);
    cudaFuncSetAttribute(search_select_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)select_smem_bytes(kSmemSelMax, 14));
    // as many CTAs as stay resident: one wave, every CTA takes the same number of queries (+-1)
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t per_sm = (uint32_t)std::max<size_t>(1, std::min<size_t>(2048 / kSmemSelThreads, (227 * 1024) / (smem + 1024)));
    grid = std::min<uint32_t>(n_queries, per_sm * (uint32_t)sms);
    search_select_smem_kernel<<<grid, kSmemSelThreads, smem, st>>>(acc_score, acc_n, n_queries, n_local, first_local, n_total,
                                                                   entry_offset, depth, leaf_lo, todo, out, out_lists,
                                                                   out_counts);
  } else {
    search_select_kernel<<<grid, kSelThreads, 0, st>>>(acc_score, acc_n, n_queries, n_local, first_local, n_total,
                                                       entry_offset, depth, leaf_lo, scratch, cnt_scratch, todo, out, out_lists,
                                                       out_counts);
  }

  if (getenv("MNEMO_TRACE_SELECT")) {   // diagnostic: how many queries the dense kernel had to rank
    uint32_t left = 0;
    cudaMemcpyAsync(&left, todo, 4, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    fprintf(stderr, "[mnemo] select: %u of %u queries ranked by the dense kernel (more than %u matched entries)\n", left,
            n_queries, cap1);
  }
}

}  // namespace mnemo
