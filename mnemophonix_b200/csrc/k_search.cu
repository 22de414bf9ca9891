// k_search.cu — LSH table build, batched bucket probing and candidate scoring.
//
// Replaces create_hash_tables / get_matches (lsh.c:55-112) and the per-signature body of
// search (search.c:122-168).
//
// Build.  The reference keeps 25 chained hash tables of `size` = total_signatures/2 heads
// (lsh.c:61) and inserts every signature under slot = big-endian-u32(bytes 4k..4k+3) % size
// (lsh.c:49-52,74).  Here each table is a CSR: one histogram pass (global reductions), one
// exclusive scan over the 25*size counters, and a STABLE least-significant-digit radix sort of the 25*n
// (bucket, id) pairs (8-bit digits, below), so that every bucket lists its members in ascending id order.  A bucket holds exactly the members of the reference's chain (collisions through the
// modulo included); the sorted order is what lets the probe kernel walk id ranges (below).
//
// Probe + score.  For one query signature the reference concatenates the 25 probed chains,
// sorts the (entry, signature) pairs, and deep-checks every pair that occurs at least twice,
// except the last group in sort order, which its loop never flushes (search.c:147-165).
// One CTA per query signature:
//   1. 25 lanes compute the probe ranges; L = total candidates;
//   2. all threads stream the candidate ids (coalesced within a bucket) through a shared-memory
//      bitmap with atomicOr: an id seen "again" (or colliding in the bitmap) goes to a short list;
//      the block also keeps the greatest id = the reference's never-flushed last group;
//   3. the short list is de-duplicated in a shared-memory hash set;
//   4. one warp per surviving id reads that database signature once (25 lanes x 4 bytes): the exact
//      number of shared buckets is recomputed from the signature itself (so bitmap collisions are
//      harmless) and the 100-byte equality count is a byte-wise compare; pairs with >= 2 shared
//      buckets and >= 30 equal bytes add (score, 1) to the query's per-entry accumulators.
// Traffic per query signature = 100 + 25*8 + 4*L + 100*D' bytes, D' = deep checks
// plus the few bitmap false positives.
#include <algorithm>
#include "common.cuh"
#include "search_internal.h"

namespace mnemo {

// ------------------------------------------------------------------ build ----

__device__ __forceinline__ uint32_t bucket_key(uint32_t word_le) { return __byte_perm(word_le, 0, 0x0123); }

// x % d for 32-bit x and d without a division (Lemire's fastmod; exact for every 32-bit pair):
// magic = 0xFFFFFFFFFFFFFFFF / d + 1, x % d = hi64((magic * x mod 2^64) * d).  `size` is fixed per database, so the
// host computes the magic once; the probe kernel takes 25 remainders per deep-checked signature.
__device__ __forceinline__ uint32_t fastmod(uint32_t x, uint64_t magic, uint32_t d) {
  return (uint32_t)__umul64hi(magic * (uint64_t)x, (uint64_t)d);
}
// key % size == slot (slot < size) without computing the remainder: key >= slot and size | (key - slot), and for
// size = 2^s * odd, n is a multiple of size <=> rotr(n * inverse(odd) mod 2^32, s) <= (2^32 - 1) / size
// (Hacker's Delight 10-17).  One multiply, one funnel shift and two compares per database word; the deep check
// does 25 of them per signature and the 64-bit fastmod was a quarter of the kernel's instructions.
__device__ __forceinline__ bool same_slot(uint32_t key, uint32_t slot, uint32_t inv, uint32_t shift, uint32_t thr) {
  const uint32_t y = (key - slot) * inv;
  return key >= slot && __funnelshift_r(y, y, shift) <= thr;
}

// one thread per (signature, table)
__global__ void __launch_bounds__(256) lsh_count_kernel(const uint32_t* __restrict__ sig_words, uint64_t n_sigs,
                                                        uint32_t size, uint32_t* __restrict__ counts) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sigs * kTables) return;
  const uint32_t slot = bucket_key(sig_words[i]) % size;   // word i = signature i/25, table i%25
  atomicAdd(&counts[(uint64_t)(i % kTables) * size + slot], 1u);
}

// sort input: key = global bucket index (table * size + slot), value = signature id; word order is id-major, so
// equal keys arrive in ascending id order and a stable sort keeps them so
__global__ void __launch_bounds__(256) lsh_pairs_kernel(const uint32_t* __restrict__ sig_words, uint64_t n_sigs,
                                                        uint32_t size, uint32_t* __restrict__ keys,
                                                        uint32_t* __restrict__ vals) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sigs * kTables) return;
  const uint32_t slot = bucket_key(sig_words[i]) % size;
  keys[i] = (uint32_t)(i % kTables) * size + slot;
  vals[i] = (uint32_t)(i / kTables);
}

__global__ void __launch_bounds__(256) sig_entry_kernel(const uint32_t* __restrict__ entry_start, uint32_t n_entries,
                                                        uint64_t n_sigs, uint32_t* __restrict__ sig_entry) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_sigs) return;
  sig_entry[g] = find_segment<uint32_t>(entry_start, n_entries, (uint32_t)g);
}

// ---- exclusive scan of a long u32 array (3 kernels) ----
constexpr int kScanT = 256, kScanE = 8, kScanB = kScanT * kScanE;

__device__ __forceinline__ uint32_t blk_excl(uint32_t v, uint32_t* ws, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) ws[warp] = incl;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kScanT / 32; ++w) {
    const uint32_t s = ws[w];
    base += w < warp ? s : 0u;
    tot += s;
  }
  *total = tot;
  __syncthreads();
  return base + incl - v;
}

__global__ void __launch_bounds__(kScanT) scan_reduce_kernel(const uint32_t* __restrict__ in, uint64_t n,
                                                             uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t ws[kScanT / 32];
  const uint64_t base = (uint64_t)blockIdx.x * kScanB + (uint64_t)threadIdx.x * kScanE;
  uint32_t c = 0;
#pragma unroll
  for (int k = 0; k < kScanE; ++k) c += (base + k < n) ? in[base + k] : 0u;
  uint32_t tot;
  blk_excl(c, ws, &tot);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanT) scan_sums_kernel(uint32_t* __restrict__ block_sums, uint32_t n_blocks) {
  __shared__ uint32_t ws[kScanT / 32];
  uint32_t carry = 0;
  for (uint32_t b0 = 0; b0 < n_blocks; b0 += kScanT) {
    const uint32_t i = b0 + threadIdx.x;
    const uint32_t v = i < n_blocks ? block_sums[i] : 0u;
    uint32_t tot;
    const uint32_t ex = blk_excl(v, ws, &tot);
    if (i < n_blocks) block_sums[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) block_sums[n_blocks] = carry;
}

// out[i] = exclusive prefix; also out2 (a second copy used as the scatter cursor); out[n] = total
__global__ void __launch_bounds__(kScanT) scan_apply_kernel(const uint32_t* __restrict__ in, uint64_t n,
                                                            const uint32_t* __restrict__ block_sums,
                                                            uint32_t* __restrict__ out, uint32_t* __restrict__ out2) {
  __shared__ uint32_t ws[kScanT / 32];
  const uint64_t base = (uint64_t)blockIdx.x * kScanB + (uint64_t)threadIdx.x * kScanE;
  uint32_t v[kScanE], c = 0;
#pragma unroll
  for (int k = 0; k < kScanE; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    c += v[k];
  }
  uint32_t tot;
  uint32_t p = block_sums[blockIdx.x] + blk_excl(c, ws, &tot);
#pragma unroll
  for (int k = 0; k < kScanE; ++k) {
    if (base + k < n) {
      out[base + k] = p;
      out2[base + k] = p;
    }
    p += v[k];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = block_sums[gridDim.x];
}

uint32_t lsh_scan_blocks(uint64_t n_slots) { return (uint32_t)((n_slots + kScanB - 1) / kScanB); }

static int key_bits(uint64_t n_slots) {
  int b = 1;
  while (b < 32 && (1ull << b) < n_slots) ++b;
  return b;
}

// ---- stable LSD radix sort of (key, value) pairs, 8 bits per pass ----
// Pass = histogram per tile (digit-major, tile-minor in global memory) -> exclusive scan of that table -> scatter.
// Stability inside a tile: a warp owns 256 consecutive elements and ranks them 32 at a time, in order, with
// match.any (lanes holding the same digit) and a per-warp counter per digit in shared memory; the warps' counters are
// then prefixed in warp order.  Equal keys therefore keep their input order: ids ascending inside every bucket.
constexpr int kSortThreads = 256, kSortItems = 8, kSortTile = kSortThreads * kSortItems;

__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint32_t* __restrict__ keys, uint64_t n, int shift,
                                                                  uint32_t n_tiles, uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * kSortTile;
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const uint64_t i = base + (uint64_t)j * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(uint64_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(kSortThreads)
    radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t n, int shift,
                         uint32_t n_tiles, const uint32_t* __restrict__ hist_scanned, uint32_t* __restrict__ keys_out,
                         uint32_t* __restrict__ vals_out) {
  __shared__ uint32_t cnt[kSortThreads / 32][256];
  __shared__ uint32_t goff[256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int w = 0; w < kSortThreads / 32; ++w) cnt[w][tid] = 0;
  goff[tid] = hist_scanned[(uint64_t)tid * n_tiles + blockIdx.x];
  __syncthreads();
  const uint64_t wbase = (uint64_t)blockIdx.x * kSortTile + (uint64_t)warp * (32 * kSortItems);
  uint32_t key[kSortItems], val[kSortItems], rank[kSortItems];
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const uint64_t i = wbase + (uint64_t)j * 32 + lane;
    key[j] = i < n ? keys_in[i] : 0u;
    val[j] = i < n ? vals_in[i] : 0u;
  }
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {   // in input order
    const bool valid = wbase + (uint64_t)j * 32 + lane < n;
    const uint32_t d = (key[j] >> shift) & 255u;
    const uint32_t act = __ballot_sync(0xffffffffu, valid);
    uint32_t r = 0;
    if (valid) {
      const uint32_t peers = __match_any_sync(act, d);
      const int leader = __ffs(peers) - 1;
      uint32_t c = 0;
      if (lane == leader) {
        c = cnt[warp][d];
        cnt[warp][d] = c + (uint32_t)__popc(peers);
      }
      c = __shfl_sync(peers, c, leader);
      r = c + (uint32_t)__popc(peers & ((1u << lane) - 1u));
    }
    rank[j] = r;
    __syncwarp();
  }
  __syncthreads();
  {
    uint32_t run = 0;   // digit tid: exclusive prefix of the warps' counts, in warp order
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w) {
      const uint32_t t = cnt[w][tid];
      cnt[w][tid] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    if (wbase + (uint64_t)j * 32 + lane < n) {
      const uint32_t d = (key[j] >> shift) & 255u;
      const uint32_t dst = goff[d] + cnt[warp][d] + rank[j];
      keys_out[dst] = key[j];
      vals_out[dst] = val[j];
    }
  }
}

static uint32_t sort_tiles(uint64_t n) { return (uint32_t)((n + kSortTile - 1) / kSortTile); }

// scratch of the sort: the digit table [256 * tiles + 1] and the block sums of its scan
size_t lsh_sort_temp_bytes(uint64_t n_sigs, uint32_t size) {
  (void)size;
  const uint64_t h = 256ull * sort_tiles(n_sigs * kTables);
  return sizeof(uint32_t) * (h + 1 + lsh_scan_blocks(h) + 1);
}

static void radix_sort_pairs(uint32_t*& ka, uint32_t*& va, uint32_t*& kb, uint32_t*& vb, uint64_t n, int bits, void* temp,
                             cudaStream_t st);

// keys_in / vals_in / keys_out: 25*n_sigs u32 each (the sort's two buffers); temp: lsh_sort_temp_bytes()
cudaError_t launch_lsh_build(const uint8_t* sigs, uint64_t n_sigs, const uint32_t* entry_start, uint32_t n_entries,
                             uint32_t size, uint32_t* counts, uint32_t* start, uint32_t* scan_copy, uint32_t* block_sums,
                             uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, void* temp, size_t temp_bytes,
                             uint32_t* ids, uint32_t* sig_entry, cudaStream_t st) {
  const uint64_t n_words = n_sigs * kTables;
  const uint64_t n_slots = (uint64_t)size * kTables;
  if (n_sigs == 0 || size == 0) return cudaSuccess;
  const uint32_t* words = reinterpret_cast<const uint32_t*>(sigs);
  cudaMemsetAsync(counts, 0, n_slots * sizeof(uint32_t), st);
  lsh_count_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, st>>>(words, n_sigs, size, counts);
  const uint32_t n_blocks = (uint32_t)((n_slots + kScanB - 1) / kScanB);
  scan_reduce_kernel<<<n_blocks, kScanT, 0, st>>>(counts, n_slots, block_sums);
  scan_sums_kernel<<<1, kScanT, 0, st>>>(block_sums, n_blocks);
  scan_apply_kernel<<<n_blocks, kScanT, 0, st>>>(counts, n_slots, block_sums, start, scan_copy);
  lsh_pairs_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, st>>>(words, n_sigs, size, keys_in, vals_in);
  // stable sort by bucket: A = (keys_in, vals_in), B = (keys_out, ids); passes alternate A -> B -> A ...
  cudaError_t e = cudaSuccess;
  {
    (void)temp_bytes;
    uint32_t *ka = keys_in, *va = vals_in, *kb = keys_out, *vb = ids;
    radix_sort_pairs(ka, va, kb, vb, n_words, key_bits(n_slots), temp, st);
    if (va != ids) e = cudaMemcpyAsync(ids, va, n_words * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st);
  }
  sig_entry_kernel<<<(unsigned)((n_sigs + 255) / 256), 256, 0, st>>>(entry_start, n_entries, n_sigs, sig_entry);
  return e;
}

// ---- pair index (probe side: k_pairprobe.cu) ----
// search.c:147-165 deep-checks the signatures that share at least TWO buckets with the query signature.  For every pair of
// tables i < j the pair index lists the database signatures ordered by (slot_i, slot_j, id) as 64-bit entries
// (slot_j << 32 | id): the members of bucket (i, s_i) occupy the same positions as in table i's CSR, and inside that
// slice the ones with slot_j == s_j are one contiguous run found by binary search.  A query signature's candidates with
// two or more shared buckets are then the union of 300 short runs instead of whatever repeats among the ~42 k members
// of its 25 buckets.  Built from what the CSR build leaves behind: table j's id list is sorted by (slot_j, id), so a
// STABLE radix sort of it by slot_i gives the (slot_i, slot_j, id) order; all j > i of one i are sorted together with
// the key (j - i - 1) << bits | slot_i.  300 * n entries of 8 bytes.

// slots[k * n + id] = bucket of signature id in table k
__global__ void __launch_bounds__(256) lsh_slot_table_kernel(const uint32_t* __restrict__ sig_words, uint64_t n_sigs,
                                                             uint32_t size, uint32_t* __restrict__ slots) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sigs * kTables) return;
  slots[(i % kTables) * n_sigs + i / kTables] = bucket_key(sig_words[i]) % size;
}

__global__ void __launch_bounds__(256) pair_keys_kernel(const uint32_t* __restrict__ ids_tail, uint64_t n_elems, uint64_t n_sigs,
                                                        const uint32_t* __restrict__ slots_i, int sbits,
                                                        uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_elems) return;
  const uint32_t id = ids_tail[t];
  keys[t] = ((uint32_t)(t / n_sigs) << sbits) | slots_i[id];
  vals[t] = id;
}

// entry = ((slot_j << fbits) | remaining) << 32 | id; `remaining` (entries that follow in the same run, saturated) is
// filled in by pair_remaining_kernel
__global__ void __launch_bounds__(256) pair_finalize_kernel(const uint32_t* __restrict__ sorted_ids, uint64_t n_elems,
                                                            uint64_t n_sigs, const uint32_t* __restrict__ slots, int i, int fbits,
                                                            unsigned long long* __restrict__ out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_elems) return;
  const uint32_t id = sorted_ids[t];
  const uint32_t j = (uint32_t)i + 1u + (uint32_t)(t / n_sigs);
  out[t] = ((unsigned long long)(slots[(uint64_t)j * n_sigs + id] << fbits) << 32) | id;
}

// Runs: maximal stretches of entries with the same (pair, slot_i, slot_j).  A run starts where bucket (i, slot_i) starts
// in the slice — table i's CSR offsets, the same in each of the row's slices — or where slot_j changes.
__global__ void __launch_bounds__(256) pair_bucket_marks_kernel(const uint32_t* __restrict__ start_i, uint32_t size,
                                                                uint64_t n_sigs, uint32_t base_i, uint32_t n_slices,
                                                                uint32_t* __restrict__ flags) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (uint64_t)size * n_slices) return;
  const uint32_t s = (uint32_t)(t % size), slice = (uint32_t)(t / size);
  const uint32_t b = start_i[s], e = start_i[s + 1];
  if (e > b) flags[(uint64_t)slice * n_sigs + (b - base_i)] = 1u;   // non-empty bucket: its first position
}
__global__ void __launch_bounds__(256) pair_flags_kernel(const unsigned long long* __restrict__ entries, uint64_t n_elems,
                                                         int fbits, uint32_t* __restrict__ flags) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_elems) return;
  if (t == 0 || ((uint32_t)(entries[t] >> 32) >> fbits) != ((uint32_t)(entries[t - 1] >> 32) >> fbits)) flags[t] = 1u;
}
// starts[r] = position of the first entry of run r; starts[n_runs] = n_elems
__global__ void __launch_bounds__(256) pair_run_starts_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ excl,
                                                              uint64_t n_elems, uint32_t* __restrict__ starts) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n_elems) return;
  if (t == n_elems) starts[excl[n_elems]] = (uint32_t)n_elems;
  else if (flags[t]) starts[excl[t]] = (uint32_t)t;
}
__global__ void __launch_bounds__(256) pair_remaining_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ excl,
                                                             const uint32_t* __restrict__ starts, uint64_t n_elems, int fbits,
                                                             unsigned long long* __restrict__ entries) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_elems) return;
  const uint32_t run = excl[t] + flags[t] - 1u;
  const uint32_t remaining = starts[run + 1] - (uint32_t)t - 1u;
  const uint32_t cap = (1u << fbits) - 1u;
  entries[t] |= (unsigned long long)(remaining < cap ? remaining : cap) << 32;
}

// stable LSD radix sort of n (key, value) pairs on the low `bits` bits of the key; the result is in (ka, va)
static void radix_sort_pairs(uint32_t*& ka, uint32_t*& va, uint32_t*& kb, uint32_t*& vb, uint64_t n, int bits, void* temp,
                             cudaStream_t st) {
  const uint32_t n_tiles = sort_tiles(n);
  const uint64_t h = 256ull * n_tiles;
  uint32_t* hist = reinterpret_cast<uint32_t*>(temp);
  uint32_t* hsums = hist + h + 1;
  const uint32_t h_blocks = lsh_scan_blocks(h);
  for (int shift = 0; shift < bits; shift += 8) {
    radix_hist_kernel<<<n_tiles, kSortThreads, 0, st>>>(ka, n, shift, n_tiles, hist);
    scan_reduce_kernel<<<h_blocks, kScanT, 0, st>>>(hist, h, hsums);
    scan_sums_kernel<<<1, kScanT, 0, st>>>(hsums, h_blocks);
    scan_apply_kernel<<<h_blocks, kScanT, 0, st>>>(hist, h, hsums, hist, hist);   // in place (each block reads its part first)
    radix_scatter_kernel<<<n_tiles, kSortThreads, 0, st>>>(ka, va, n, shift, n_tiles, hist, kb, vb);
    uint32_t* t = ka; ka = kb; kb = t;
    t = va; va = vb; vb = t;
  }
}

// sum over all buckets of len^2: divided by the number of signatures it is the list length a query signature that
// resembles the database meets in ONE table (a bucket is hit in proportion to its size)
__global__ void __launch_bounds__(256) lsh_bucket_load_kernel(const uint32_t* __restrict__ start, uint64_t n_slots,
                                                              unsigned long long* __restrict__ sum_sq) {
  unsigned long long acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long len = start[i + 1] - start[i];
    acc += len * len;
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(sum_sq, acc);
}
void launch_bucket_load(const uint32_t* start, uint64_t n_slots, unsigned long long* sum_sq, cudaStream_t st) {
  cudaMemsetAsync(sum_sq, 0, 8, st);
  if (n_slots) lsh_bucket_load_kernel<<<592, 256, 0, st>>>(start, n_slots, sum_sq);
}

// Run directory: one open-addressing table per row i (all slices j > i together), two slots per index entry, mapping
// (slice, slot_i, slot_j) -> position of the run's first entry in the row's array.  A slot holds only that position:
// the probe kernel recognises its run by the position lying inside bucket (i, slot_i) of its slice and by the slot_j
// of the entry found there (search_internal.h: pair_dir_hash).  Built from the run starts of the remaining-length pass.
__global__ void __launch_bounds__(256) pair_dir_insert_kernel(const unsigned long long* __restrict__ entries,
                                                              const uint32_t* __restrict__ flags, uint64_t n_elems,
                                                              uint64_t n_sigs, const uint32_t* __restrict__ slots_i, int fbits,
                                                              uint32_t* __restrict__ dir, uint32_t capacity) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_elems || !flags[t]) return;
  const unsigned long long e = entries[t];
  const uint32_t slice = (uint32_t)(t / n_sigs);
  uint32_t idx = pair_dir_slot(pair_dir_hash(slice, slots_i[(uint32_t)e], (uint32_t)(e >> 32) >> fbits), capacity);
  while (atomicCAS(&dir[idx], 0xFFFFFFFFu, (uint32_t)t) != 0xFFFFFFFFu) idx = idx + 1 == capacity ? 0 : idx + 1;
}

bool pair_index_possible(uint32_t size) { return size > 0 && key_bits(size) + 5 <= 32; }
// bits of an entry's high word left of slot_j for the saturated remaining-run-length field (5 .. 12)
int pair_index_fbits(uint32_t size) { return std::min(12, 32 - key_bits(size)); }
uint32_t pair_index_row_base(int i) { return (uint32_t)(i * (2 * (kTables - 1) - i + 1) / 2); }   // pairs (a, .) with a < i

// start: the CSR offsets (25 size + 1); ids: the CSR id lists (25 n); slots: 25 n scratch; ka / va / kb / vb: 25 n scratch
// each; temp: lsh_sort_temp_bytes(); pairs: 300 n entries
// dir: 600 n slots (or nullptr: no directory, the probe kernel searches the sorted slices)
cudaError_t launch_pair_build(const uint8_t* sigs, uint64_t n_sigs, uint32_t size, const uint32_t* start, const uint32_t* ids,
                              uint32_t* slots, uint32_t* ka, uint32_t* va, uint32_t* kb, uint32_t* vb, void* temp,
                              unsigned long long* pairs, uint32_t* dir, cudaStream_t st) {
  if (n_sigs == 0 || size == 0) return cudaSuccess;
  const uint64_t n_words = n_sigs * kTables;
  lsh_slot_table_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint32_t*>(sigs), n_sigs, size, slots);
  const int sbits = key_bits(size), fbits = pair_index_fbits(size);
  uint32_t* block_sums = reinterpret_cast<uint32_t*>(temp);
  for (int i = 0; i + 1 < kTables; ++i) {
    const uint32_t n_slices = (uint32_t)(kTables - 1 - i);
    const uint64_t n_elems = (uint64_t)n_slices * n_sigs;
    const unsigned grid = (unsigned)((n_elems + 255) / 256);
    uint32_t *a = ka, *b = va, *c = kb, *d = vb;
    unsigned long long* row = pairs + (uint64_t)pair_index_row_base(i) * n_sigs;
    pair_keys_kernel<<<grid, 256, 0, st>>>(ids + (uint64_t)(i + 1) * n_sigs, n_elems, n_sigs, slots + (uint64_t)i * n_sigs, sbits, a, b);
    radix_sort_pairs(a, b, c, d, n_elems, sbits + 5, temp, st);
    pair_finalize_kernel<<<grid, 256, 0, st>>>(b, n_elems, n_sigs, slots, i, fbits, row);
    // remaining run lengths: flags (a) -> exclusive scan (c) -> run starts (d) -> field
    cudaMemsetAsync(a, 0, n_elems * 4, st);
    const uint64_t n_marks = (uint64_t)size * n_slices;
    pair_bucket_marks_kernel<<<(unsigned)((n_marks + 255) / 256), 256, 0, st>>>(start + (uint64_t)i * size, size, n_sigs,
                                                                               (uint32_t)((uint64_t)i * n_sigs), n_slices, a);
    pair_flags_kernel<<<grid, 256, 0, st>>>(row, n_elems, fbits, a);
    const uint32_t n_blocks = lsh_scan_blocks(n_elems);
    scan_reduce_kernel<<<n_blocks, kScanT, 0, st>>>(a, n_elems, block_sums);
    scan_sums_kernel<<<1, kScanT, 0, st>>>(block_sums, n_blocks);
    scan_apply_kernel<<<n_blocks, kScanT, 0, st>>>(a, n_elems, block_sums, c, c);
    pair_run_starts_kernel<<<(unsigned)((n_elems + 256) / 256), 256, 0, st>>>(a, c, n_elems, d);
    pair_remaining_kernel<<<grid, 256, 0, st>>>(a, c, d, n_elems, fbits, row);
    if (dir) {
      uint32_t* dir_i = dir + 2ull * pair_index_row_base(i) * n_sigs;
      cudaMemsetAsync(dir_i, 0xFF, 8ull * n_elems, st);
      pair_dir_insert_kernel<<<grid, 256, 0, st>>>(row, a, n_elems, n_sigs, slots + (uint64_t)i * n_sigs, fbits, dir_i,
                                                   (uint32_t)(2 * n_elems));
    }
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------ probe + score ----

constexpr int kSearchThreads = 256;
constexpr uint32_t kDupCap = 2048;        // short list capacity (ids seen "again")
constexpr uint32_t kSetCap = 4096;        // hash set slots for de-duplication (power of two, >= 2 * kDupCap)
// 16 KB bitmap, 48 KB of dynamic shared memory per CTA, 64 registers: FOUR CTAs per SM.  (8192 words at three CTAs per SM
// was the round-1 shape; with the big databases on the pair index this kernel serves shards and small databases, where
// the extra CTA buys 4-9 %: 10.1 -> 9.2 / 14.2 -> 13.4 / 24.9 -> 23.8 ms per 10 k-query batch at 180 k / 380 k / 760 k signatures.)
#ifndef PROBE_BITMAP_WORDS
#define PROBE_BITMAP_WORDS 4096
#endif
#ifndef PROBE_MIN_BLOCKS
#define PROBE_MIN_BLOCKS 4
#endif
constexpr uint32_t kBitmapMaxWords = PROBE_BITMAP_WORDS;
constexpr uint32_t kEmpty = 0xFFFFFFFFu;
constexpr uint32_t kPassCap = 8192;          // candidates per id-space partition (>= 32 bitmap bits each)
constexpr uint32_t kRangeBits = kBitmapMaxWords * 32;   // ids per pass in range mode: one bitmap bit per id, exact
constexpr uint32_t kMaxRangePasses = 16;                // id ranges whose tables are held at a time (2.1 M signatures)
constexpr uint32_t kStageCap = kSetCap + kDupCap;       // ids of one id range staged in shared memory (the flush's set + uniq areas)

uint32_t search_range_bits() { return kRangeBits; }
size_t search_smem_bytes() { return sizeof(uint32_t) * (kBitmapMaxWords + kDupCap + kSetCap + kDupCap); }

struct Probe {
  uint32_t begin[kTables];
  uint32_t prefix[kTables + 1];
  uint32_t qslot[kTables];
  uint32_t qword[kTables];
};

// bytes in which two words differ (search.c:35-43 counts the equal ones)
__device__ __forceinline__ uint32_t differing_bytes(uint32_t a, uint32_t b) {
  const uint32_t x = a ^ b;
  return __popc((((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u);
}

// Deep check of one database signature against the query, by one warp.  Returns (hits, score).
__device__ __forceinline__ void check_pair(const uint32_t* __restrict__ db_words, uint32_t g, const Probe& pr,
                                           uint32_t size, uint64_t magic, int lane, uint32_t* hits, uint32_t* score) {
  uint32_t h = 0, eq = 0;
  if (lane < kTables) {
    const uint32_t w = __ldg(db_words + (uint64_t)g * kTables + lane);
    h = fastmod(bucket_key(w), magic, size) == pr.qslot[lane];
    eq = __popc(__vcmpeq4(w, pr.qword[lane])) >> 3;   // search.c:35-43: number of equal bytes
  }
  *hits = __popc(__ballot_sync(0xffffffffu, h));
  *score = __reduce_add_sync(0xffffffffu, eq);
}

__device__ __forceinline__ int table_of(const uint32_t* prefix, uint32_t i) {
  int k = 0;
#pragma unroll
  for (int s = 16; s; s >>= 1)
    if (k + s <= kTables - 1 && prefix[k + s] <= i) k += s;
  return k;   // largest k with prefix[k] <= i
}

// Candidate de-duplication has two modes, chosen per query signature:
//   hashed (L <= kPassCap, the usual case for small databases): one pass, bitmap indexed by a hash of the id with
//     >= 32 bits per candidate; a collision only costs one extra deep check (hits are recomputed exactly);
//   id ranges (long candidate lists): buckets are sorted by id, so the candidates with ids in
//     [pass * kRangeBits, (pass + 1) * kRangeBits) are one contiguous piece of each of the 25 lists, found by binary
//     search; each pass streams just those through a bitmap with one bit per id (exact, no false repeats).  Every
//     candidate id is read once however long the lists are, and passes without candidates cost nothing.
//   Shards of any size: the ranges are walked kMaxRangePasses at a time.
__device__ __forceinline__ void probe_query_signature(const SearchParams& p, const uint32_t qs) {
  extern __shared__ __align__(16) uint32_t sm[];
  uint32_t* bitmap = sm;
  uint32_t* dup = bitmap + kBitmapMaxWords;
  uint32_t* set = dup + kDupCap;
  uint32_t* uniq = set + kSetCap;
  __shared__ Probe pr;
  __shared__ uint32_t s_bound[kTables][kMaxRangePasses + 1];
  __shared__ uint32_t s_pref[kMaxRangePasses][kTables + 1];   // range mode: prefix of the piece lengths of every pass
  __shared__ uint32_t s_so[kMaxRangePasses][kTables + 1];     // the same for the pieces padded to 16-byte units (stage offsets)
  __shared__ uint32_t s_ndup, s_nuniq, s_gmax;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t* __restrict__ qwords = reinterpret_cast<const uint32_t*>(p.query_sigs) + (uint64_t)qs * kTables;

  if (warp == 0) {   // the 25 probed buckets and the prefix of their lengths (warp scan: one barrier, no serial loop)
    uint32_t len = 0;
    if (lane < kTables) {
      const uint32_t w = qwords[lane];
      const uint32_t slot = fastmod(bucket_key(w), p.size_magic, p.size);
      pr.qword[lane] = w;
      pr.qslot[lane] = slot;
      MN_CHECK(slot < p.size);
      const uint32_t b = p.start[(uint64_t)lane * p.size + slot], e = p.start[(uint64_t)lane * p.size + slot + 1];
      MN_CHECK(b <= e && (uint64_t)e <= (uint64_t)kTables * p.n_sigs);
      pr.begin[lane] = b;
      len = e - b;
    }
    uint32_t incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane < kTables) pr.prefix[lane + 1] = incl;
    if (lane == 0) {
      pr.prefix[0] = 0;
      s_gmax = 0;
    }
  }
  __syncthreads();
  const uint32_t L = pr.prefix[kTables];
  const uint32_t query = __ldg(p.qs_query + qs);   // (a bisection of query_start here was 14 dependent loads per CTA)
  mnemo_lastgroup_dev lg = {0, 0, 0, 0};
  if (L == 0) {
    if (tid == 0) p.last[qs] = lg;
    return;
  }
  const uint32_t range_ids = p.range_ids;   // ids per pass in range mode (kRangeBits unless a test shrinks it)
  const uint32_t range_passes = (p.n_sigs + range_ids - 1) / range_ids;
  const bool range_mode = L > kPassCap;     // block-uniform
  uint32_t bm_words = 64;   // hashed mode: ~64 bitmap bits per candidate, 256 B .. 32 KB
  while (bm_words < kBitmapMaxWords && bm_words < 2 * L) bm_words <<= 1;
  const uint32_t bm_shift = 32 - (31 - __clz(bm_words * 32));
  const uint32_t* __restrict__ dbw = reinterpret_cast<const uint32_t*>(p.db_sigs);
  uint32_t deep = 0, deep_t = 0, gmax = 0;   // deep checks counted per warp (fallback path) / per thread

  // Deep checks of the distinct repeated ids lst[0 .. n): warps take groups of 32.  Lane k (< 25) loads word k of each
  // of the group's 32 signatures — 32 coalesced 100-byte loads in flight per warp — then one warp reduction per
  // signature yields its shared-bucket count and its equal-byte count together, and lane r scores signature r.
  // (One thread per signature with 25 loads of its own cost 32 L1 sectors per load instruction and made the L1 the
  // bound of the kernel: 84 of 150 ms per batch at config #4; one warp per signature kept a single row in flight.)
  // Two equal words are two shared buckets (same key, same slot), so the reduction counts differing bytes and EQUAL
  // words; only a signature with fewer than two equal words can still share two buckets through key % size
  // collisions (lsh.c:74), and only for those the slots are recomputed from the signature (same_slot).
  auto check_list = [&](const uint32_t* lst, uint32_t n, uint32_t g_skip) {
    const uint32_t my_slot = lane < kTables ? pr.qslot[lane] : 0u, my_word = lane < kTables ? pr.qword[lane] : 0u;
    for (uint32_t base = (uint32_t)warp * 32u; base < n; base += kSearchThreads) {
      const uint32_t cnt = min(32u, n - base);
      const uint32_t my_g = lst[base + ((uint32_t)lane < cnt ? lane : 0)];
      MN_CHECK(my_g < p.n_sigs);
      const uint32_t my_entry = __ldg(p.sig_entry + my_g);   // in flight with the rows, not after them
      uint32_t w[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const uint32_t g = __shfl_sync(0xffffffffu, my_g, r);
        MN_CHECK(g < p.n_sigs);
        w[r] = lane < kTables ? __ldg(dbw + (uint64_t)g * kTables + lane) : 0u;   // lanes >= 25 hold "equal" words
      }
      uint32_t mine = 0;
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        uint32_t v = differing_bytes(w[r], my_word) | (w[r] == my_word ? 0x10000u : 0u);
        v = __reduce_add_sync(0xffffffffu, v);
        if (lane == r) mine = v;
      }
      const uint32_t score = 100u - (mine & 0xFFFFu);   // search.c:35-43
      uint32_t hits = (mine >> 16) - (32u - kTables);
      uint32_t slow = __ballot_sync(0xffffffffu, (uint32_t)lane < cnt && hits < 2u);
      while (slow) {
        const int r = __ffs(slow) - 1;
        slow &= slow - 1u;
        const uint32_t g = __shfl_sync(0xffffffffu, my_g, r);
        uint32_t h = 0;
        if (lane < kTables)
          h = same_slot(bucket_key(__ldg(dbw + (uint64_t)g * kTables + lane)), my_slot, p.div_inv, p.div_shift, p.div_thr);
        const uint32_t exact = __popc(__ballot_sync(0xffffffffu, h));
        if (lane == r) hits = exact;
      }
      if ((uint32_t)lane < cnt && my_g != g_skip && hits >= 2u) {   // search.c:147-165 (last group never flushed), MIN_BUCKET_MATCH_FOR_DEEP_CHECK
        ++deep_t;
        if (score >= 30u) {                                         // MIN_SCORE (search.c:16)
          const uint64_t a = (uint64_t)query * p.n_entries + my_entry;
          MN_CHECK(a < (uint64_t)p.n_queries * p.n_entries);
          atomicAdd(&p.acc_score[a], score);
          atomicAdd(&p.acc_n[a], 1u);
        }
      }
    }
  };

  if (range_mode) {
    // the greatest candidate id is the greatest last element of the sorted lists
    if (warp == 0) {
      uint32_t v = 0;
      if (lane < kTables) {
        const uint32_t len = pr.prefix[lane + 1] - pr.prefix[lane];
        if (len) v = __ldg(p.ids + pr.begin[lane] + len - 1);
      }
      v = __reduce_max_sync(0xffffffffu, v);
      if (lane == 0) s_gmax = v;
    }
    for (uint32_t i = tid; i < kBitmapMaxWords / 4; i += kSearchThreads) reinterpret_cast<uint4*>(bitmap)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
      s_ndup = 0;
      s_nuniq = 0;
    }
    uint32_t listed = 0;   // block-uniform: entries waiting in dup[]
    // the id ranges are walked kMaxRangePasses at a time (the per-pass tables live in shared memory)
    for (uint32_t c0 = 0; c0 < range_passes; c0 += kMaxRangePasses) {
    const uint32_t n_pass = min(kMaxRangePasses, range_passes - c0);
    // where each id range starts in each list
    for (uint32_t t = tid; t < kTables * (n_pass + 1); t += kSearchThreads) {
      const uint32_t k = t / (n_pass + 1), j = t - k * (n_pass + 1);
      const uint32_t len = pr.prefix[k + 1] - pr.prefix[k];
      const uint32_t* __restrict__ lst = p.ids + pr.begin[k];
      const uint32_t target = (c0 + j) * range_ids;   // <= n_sigs + range_ids: no overflow
      uint32_t lo = 0, hi = len;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(lst + mid) < target) lo = mid + 1;
        else hi = mid;
      }
      s_bound[k][j] = lo;
    }
    __syncthreads();
    for (uint32_t j = tid; j < n_pass; j += kSearchThreads) {
      uint32_t c = 0, cs = 0;
      for (int k = 0; k < kTables; ++k) {
        s_pref[j][k] = c;
        s_so[j][k] = cs;
        const uint32_t n = s_bound[k][j + 1] - s_bound[k][j];
        c += n;
        // staged form: the piece's source is aligned down to 16 bytes, its length rounded up to whole 16-byte units
        if (n) cs += ((((pr.begin[k] + s_bound[k][j]) & 3u) + n + 3u) >> 2) << 2;
      }
      s_pref[j][kTables] = c;
      s_so[j][kTables] = cs;
    }
    __syncthreads();

    // ---- id ranges: the CTA's latency chain is what bounds this kernel at long lists, so a pass is two barriers:
    // stream (mark + collect repeats) | clear the bitmap for the next pass.  Everything a pass needs was computed
    // above for all passes at once; repeats are scored together every few passes (flush).
    const uint32_t g_last = s_gmax;
    auto flush = [&]() {
      if (listed == 0) return;
      const uint32_t ndup = listed;
      uint32_t set_cap = 64;
      while (set_cap < 2 * ndup) set_cap <<= 1;   // <= kSetCap because ndup <= kDupCap
      for (uint32_t i = tid; i < set_cap; i += kSearchThreads) set[i] = kEmpty;
      __syncthreads();
      for (uint32_t i = tid; i < ndup; i += kSearchThreads) {
        const uint32_t g = dup[i];
        uint32_t h = (g * 2246822519u) & (set_cap - 1);
        while (true) {
          const uint32_t prev = atomicCAS(&set[h], kEmpty, g);
          if (prev == kEmpty) {
            uniq[atomicAdd(&s_nuniq, 1u)] = g;
            break;
          }
          if (prev == g) break;
          h = (h + 1) & (set_cap - 1);
        }
      }
      __syncthreads();
      const uint32_t nuniq = s_nuniq;
      check_list(uniq, nuniq, g_last);
      __syncthreads();
      if (tid == 0) {
        s_ndup = 0;
        s_nuniq = 0;
      }
      listed = 0;
      __syncthreads();
    };
    // A range costs two barriers: [mark] | [clear the bitmap, flush if the short list is about to fill, stage the NEXT
    // range] |.  The next range's ids are copied into shared memory (set + uniq, idle outside a flush) while the bitmap
    // of this one is cleared: every warp issues 16-byte asynchronous copies for the pieces of its lists — the whole
    // range in flight at once, no registers held.  A piece starts at any word, so its source is aligned down (up to 3
    // ids of the neighbouring bucket in front, up to 3 behind) and those strangers are blanked once the copy has
    // landed.  The marking loop then reads a flat array: no candidate -> (list, offset) walk, no global load.
    uint32_t* const stage = set;
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    auto next_pass = [&](uint32_t from) {   // first range at or after `from` in which something can repeat
      while (from < n_pass && s_pref[from][kTables] < 2) ++from;
      return from;
    };
    auto stage_pass = [&](uint32_t pass) {
      if (s_so[pass][kTables] > kStageCap) return;   // a crowded range is marked straight from global memory
      for (int k = warp; k < kTables; k += kSearchThreads / 32) {
        const uint32_t b = s_bound[k][pass], n = s_bound[k][pass + 1] - b;
        if (!n) continue;
        const uint32_t first = pr.begin[k] + b, units = ((first & 3u) + n + 3u) >> 2;
        const uint32_t dst = stage_s + (s_so[pass][k] << 2);
        const uint32_t* src = p.ids + (first & ~3u);
        MN_CHECK(s_so[pass][k] + 4 * units <= kStageCap && (uint64_t)(first & ~3u) + 4ull * units <= (uint64_t)kTables * p.n_sigs + 4);
        for (uint32_t u = lane; u < units; u += 32)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (u << 4)), "l"(src + (u << 2)) : "memory");
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncwarp();
      for (int k = warp; k < kTables; k += kSearchThreads / 32) {
        const uint32_t b = s_bound[k][pass], n = s_bound[k][pass + 1] - b;
        if (!n) continue;
        const uint32_t lead = (pr.begin[k] + b) & 3u, words = ((lead + n + 3u) >> 2) << 2;
        uint32_t* dst = stage + s_so[pass][k];
        if ((uint32_t)lane < lead) dst[lane] = kEmpty;
        if ((uint32_t)lane < words - lead - n) dst[lead + n + lane] = kEmpty;
      }
    };
    uint32_t pass = next_pass(0);
    if (pass < n_pass) stage_pass(pass);   // nothing is listed at the start of a group of ranges
    __syncthreads();
    while (pass < n_pass) {
      const uint32_t* __restrict__ pref = s_pref[pass];
      const uint32_t Lp = pref[kTables];
      const uint32_t id_base = (c0 + pass) * range_ids;
      auto visit = [&](uint32_t g) {
        const uint32_t h = g - id_base;
        const uint32_t bit = 1u << (h & 31u);
        MN_CHECK(g >= id_base && (h >> 5) < kBitmapMaxWords && g < p.n_sigs);
        if (atomicOr(&bitmap[h >> 5], bit) & bit) {
          const uint32_t pos = atomicAdd(&s_ndup, 1u);
          if (pos < kDupCap) dup[pos] = g;
        }
      };
      const uint32_t staged = s_so[pass][kTables];
      if (staged <= kStageCap) {
        uint32_t i = tid;
        for (; i + 3 * kSearchThreads < staged; i += 4 * kSearchThreads) {
          const uint32_t g0 = stage[i], g1 = stage[i + kSearchThreads], g2 = stage[i + 2 * kSearchThreads],
                         g3 = stage[i + 3 * kSearchThreads];
          if (g0 != kEmpty) visit(g0);
          if (g1 != kEmpty) visit(g1);
          if (g2 != kEmpty) visit(g2);
          if (g3 != kEmpty) visit(g3);
        }
        for (; i < staged; i += kSearchThreads) {
          const uint32_t g = stage[i];
          if (g != kEmpty) visit(g);
        }
      } else {
        // a crowded range (more ids than the stage holds): straight from global memory through the flat index
        uint32_t k = 0, seg_end = pref[1];
        const uint32_t* src = p.ids + pr.begin[0] + s_bound[0][pass];   // src + i addresses candidate i while i < seg_end
        auto locate = [&](uint32_t i) -> const uint32_t* {
          while (i >= seg_end) {
            ++k;
            seg_end = pref[k + 1];
            src = p.ids + pr.begin[k] + s_bound[k][pass] - pref[k];
          }
          return src + i;
        };
        uint32_t i = tid;
        for (; i + 3 * kSearchThreads < Lp; i += 4 * kSearchThreads) {
          const uint32_t* a0 = locate(i);
          const uint32_t* a1 = locate(i + kSearchThreads);
          const uint32_t* a2 = locate(i + 2 * kSearchThreads);
          const uint32_t* a3 = locate(i + 3 * kSearchThreads);
          const uint32_t g0 = __ldg(a0), g1 = __ldg(a1), g2 = __ldg(a2), g3 = __ldg(a3);
          visit(g0);
          visit(g1);
          visit(g2);
          visit(g3);
        }
        for (; i < Lp; i += kSearchThreads) visit(__ldg(locate(i)));
      }
      __syncthreads();   // every mark is in, s_ndup is final
      const uint32_t ndup = s_ndup;
      for (uint32_t i = tid; i < kBitmapMaxWords / 4; i += kSearchThreads) reinterpret_cast<uint4*>(bitmap)[i] = make_uint4(0, 0, 0, 0);
      if (ndup <= kDupCap) {
        listed = ndup;   // scored at the next flush
      } else {
        // more repeats than the list holds: this pass's entries are dropped from it and the pass is scored list-free
        // (one warp per candidate; the pair is scored only from its lowest matching list, i.e. exactly once)
        __syncthreads();             // everyone has read s_ndup
        if (tid == 0) s_ndup = listed;
        for (int k = 0; k < kTables; ++k) {
          const uint32_t b = s_bound[k][pass], n = s_bound[k][pass + 1] - b;
          const uint32_t* __restrict__ lst = p.ids + pr.begin[k] + b;
          for (uint32_t j = warp; j < n; j += kSearchThreads / 32) {
            const uint32_t g = __ldg(lst + j);
            if (g == g_last) continue;
            uint32_t h = 0, eq = 0;
            if (lane < kTables) {
              const uint32_t w = __ldg(dbw + (uint64_t)g * kTables + lane);
              h = same_slot(bucket_key(w), pr.qslot[lane], p.div_inv, p.div_shift, p.div_thr);
              eq = __popc(__vcmpeq4(w, pr.qword[lane])) >> 3;
            }
            const uint32_t mask = __ballot_sync(0xffffffffu, h);
            const uint32_t score = __reduce_add_sync(0xffffffffu, eq);
            if (__popc(mask) >= 2 && (__ffs(mask) - 1) == k) {
              ++deep;
              if (score >= 30 && lane == 0) {
                const uint64_t a = (uint64_t)query * p.n_entries + p.sig_entry[g];
                MN_CHECK(a < (uint64_t)p.n_queries * p.n_entries);
          atomicAdd(&p.acc_score[a], score);
                atomicAdd(&p.acc_n[a], 1u);
              }
            }
          }
        }
      }
      const uint32_t nxt = next_pass(pass + 1);
      if (nxt < n_pass) {
        if (listed && listed + s_pref[nxt][kTables] / 4 > kDupCap) flush();   // (uses the stage area: before staging)
        stage_pass(nxt);
      }
      __syncthreads();   // bitmap clean, next range staged, and nobody still reads s_ndup when the next range appends
      pass = nxt;
    }
    if (c0 + kMaxRangePasses < range_passes) {
      flush();
      __syncthreads();   // the per-pass tables are rewritten for the next group of ranges
    } else {
      flush();
    }
    }
  }

  // ---- hashed mode: every pass streams the whole lists (pr.begin / pr.prefix) and keeps its hash partition ----
  for (uint32_t pass = 0; !range_mode && pass < 1; ++pass) {   // hashed mode is one pass (L <= kPassCap)
    if (tid == 0) {
      s_ndup = 0;
      s_nuniq = 0;
    }
    const uint32_t Lp = L;
    for (uint32_t i = tid; i < bm_words / 4; i += kSearchThreads) reinterpret_cast<uint4*>(bitmap)[i] = make_uint4(0, 0, 0, 0);
    // stage offsets of the 25 lists padded to 16-byte units (as in the id-range mode; s_so[0] is free here)
    if (warp == 0) {
      uint32_t w = 0;
      if (lane < kTables) {
        const uint32_t n = pr.prefix[lane + 1] - pr.prefix[lane];
        if (n) w = (((pr.begin[lane] & 3u) + n + 3u) >> 2) << 2;
      }
      uint32_t incl = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      if (lane <= kTables) s_so[0][lane] = incl - w;   // lane 25: the total
    }
    __syncthreads();
    const uint32_t* const seg_prefix = pr.prefix;
    const uint32_t* const seg_begin = pr.begin;
    auto visit = [&](uint32_t g) {
      gmax = max(gmax, g);
      const uint32_t h = (g * 2654435761u) >> bm_shift;
      const uint32_t bit = 1u << (h & 31u);
      MN_CHECK((h >> 5) < bm_words && g < p.n_sigs);
      const uint32_t old = atomicOr(&bitmap[h >> 5], bit);
      if (old & bit) {
        const uint32_t pos = atomicAdd(&s_ndup, 1u);
        if (pos < kDupCap) dup[pos] = g;
      }
    };
    const uint32_t staged = s_so[0][kTables];
    if (staged <= kStageCap) {
      // the whole candidate list goes through shared memory (set + uniq are idle until the de-duplication below):
      // 16-byte asynchronous copies, strangers of the aligned-down pieces blanked, flat marking loop
      uint32_t* const stage = set;
      const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
      for (int k = warp; k < kTables; k += kSearchThreads / 32) {
        const uint32_t n = seg_prefix[k + 1] - seg_prefix[k];
        if (!n) continue;
        const uint32_t first = seg_begin[k], units = ((first & 3u) + n + 3u) >> 2;
        const uint32_t dst = stage_s + (s_so[0][k] << 2);
        const uint32_t* src = p.ids + (first & ~3u);
        MN_CHECK(s_so[0][k] + 4 * units <= kStageCap && (uint64_t)(first & ~3u) + 4ull * units <= (uint64_t)kTables * p.n_sigs + 4);
        for (uint32_t u = lane; u < units; u += 32)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (u << 4)), "l"(src + (u << 2)) : "memory");
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncwarp();
      for (int k = warp; k < kTables; k += kSearchThreads / 32) {
        const uint32_t n = seg_prefix[k + 1] - seg_prefix[k];
        if (!n) continue;
        const uint32_t lead = seg_begin[k] & 3u, words = ((lead + n + 3u) >> 2) << 2;
        uint32_t* dst = stage + s_so[0][k];
        if ((uint32_t)lane < lead) dst[lane] = kEmpty;
        if ((uint32_t)lane < words - lead - n) dst[lead + n + lane] = kEmpty;
      }
      __syncthreads();
      uint32_t i = tid;
      for (; i + 3 * kSearchThreads < staged; i += 4 * kSearchThreads) {
        const uint32_t g0 = stage[i], g1 = stage[i + kSearchThreads], g2 = stage[i + 2 * kSearchThreads],
                       g3 = stage[i + 3 * kSearchThreads];
        if (g0 != kEmpty) visit(g0);
        if (g1 != kEmpty) visit(g1);
        if (g2 != kEmpty) visit(g2);
        if (g3 != kEmpty) visit(g3);
      }
      for (; i < staged; i += kSearchThreads) {
        const uint32_t g = stage[i];
        if (g != kEmpty) visit(g);
      }
    } else {
      // flat index over the candidates; the thread's position in the list table only moves forward, and four loads
      // are in flight before the first id is used
      uint32_t k = 0, seg_end = seg_prefix[1];
      const uint32_t* src = p.ids + seg_begin[0];   // src + i addresses candidate i while i < seg_end
      auto locate = [&](uint32_t i) -> const uint32_t* {
        while (i >= seg_end) {
          ++k;
          seg_end = seg_prefix[k + 1];
          src = p.ids + seg_begin[k] - seg_prefix[k];
        }
        return src + i;
      };
      uint32_t i = tid;
      for (; i + 3 * kSearchThreads < Lp; i += 4 * kSearchThreads) {
        const uint32_t* a0 = locate(i);
        const uint32_t* a1 = locate(i + kSearchThreads);
        const uint32_t* a2 = locate(i + 2 * kSearchThreads);
        const uint32_t* a3 = locate(i + 3 * kSearchThreads);
        const uint32_t g0 = __ldg(a0), g1 = __ldg(a1), g2 = __ldg(a2), g3 = __ldg(a3);
        visit(g0);
        visit(g1);
        visit(g2);
        visit(g3);
      }
      for (; i < Lp; i += kSearchThreads) visit(__ldg(locate(i)));
    }
    if (pass == 0) {
      gmax = __reduce_max_sync(0xffffffffu, gmax);
      if (lane == 0) atomicMax(&s_gmax, gmax);
    }
    __syncthreads();
    gmax = s_gmax;
    const uint32_t ndup = s_ndup;
    if (ndup <= kDupCap) {
      // de-duplicate the short list (hash set sized to it), then one warp per distinct id
      uint32_t set_cap = 64;
      while (set_cap < 2 * ndup) set_cap <<= 1;   // <= kSetCap because ndup <= kDupCap
      for (uint32_t i = tid; i < set_cap; i += kSearchThreads) set[i] = kEmpty;
      __syncthreads();
      for (uint32_t i = tid; i < ndup; i += kSearchThreads) {
        const uint32_t g = dup[i];
        uint32_t h = (g * 2246822519u) & (set_cap - 1);
        while (true) {
          const uint32_t prev = atomicCAS(&set[h], kEmpty, g);
          if (prev == kEmpty) {
            uniq[atomicAdd(&s_nuniq, 1u)] = g;
            break;
          }
          if (prev == g) break;
          h = (h + 1) & (set_cap - 1);
        }
      }
      __syncthreads();
      const uint32_t nuniq = s_nuniq;
      check_list(uniq, nuniq, gmax);
    } else {
      // More repeats than the short list holds (pathological buckets): no list at all.  One warp
      // per candidate recomputes the exact shared-bucket mask from the signature, and the pair is
      // scored only from its lowest matching table, i.e. exactly once.
      for (uint32_t i = warp; i < Lp; i += kSearchThreads / 32) {
        const int k = table_of(seg_prefix, i);
        const uint32_t g = __ldg(p.ids + seg_begin[k] + (i - seg_prefix[k]));
        if (g == gmax) continue;
        uint32_t h = 0, eq = 0;
        if (lane < kTables) {
          const uint32_t w = __ldg(dbw + (uint64_t)g * kTables + lane);
          h = fastmod(bucket_key(w), p.size_magic, p.size) == pr.qslot[lane];
          eq = __popc(__vcmpeq4(w, pr.qword[lane])) >> 3;
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, h);
        const uint32_t score = __reduce_add_sync(0xffffffffu, eq);
        if (__popc(mask) >= 2 && (__ffs(mask) - 1) == k) {
          ++deep;
          if (score >= 30 && lane == 0) {
            const uint64_t a = (uint64_t)query * p.n_entries + p.sig_entry[g];
            MN_CHECK(a < (uint64_t)p.n_queries * p.n_entries);
          atomicAdd(&p.acc_score[a], score);
            atomicAdd(&p.acc_n[a], 1u);
          }
        }
      }
    }
    __syncthreads();
  }
  gmax = s_gmax;
  if (warp == 0) {
    uint32_t hits, score;
    check_pair(dbw, gmax, pr, p.size, p.size_magic, lane, &hits, &score);
    if (lane == 0) {
      lg.has = 1;
      lg.g = gmax;
      lg.hits = hits;
      lg.score = score;
      p.last[qs] = lg;
    }
  }
  deep += __reduce_add_sync(0xffffffffu, deep_t);
  if (lane == 0) {
    if (deep) atomicAdd(p.stat_deep, (unsigned long long)deep);
    if (warp == 0) atomicAdd(p.stat_nodes, (unsigned long long)L);
  }
}

// one CTA per query signature
__global__ void __launch_bounds__(kSearchThreads, PROBE_MIN_BLOCKS) search_probe_kernel(SearchParams p) { probe_query_signature(p, blockIdx.x); }

// Per query: compact the non-zero per-entry accumulators into candidate records, ascending entry.
__global__ void __launch_bounds__(256) search_compact_kernel(const uint32_t* __restrict__ acc_score,
                                                             const uint32_t* __restrict__ acc_n, uint32_t n_entries,
                                                             uint32_t entry_base, CandDev* __restrict__ cands,
                                                             uint32_t cap, uint32_t* __restrict__ total,
                                                             uint32_t* __restrict__ q_base, uint32_t* __restrict__ q_cnt) {
  __shared__ uint32_t ws[8];
  __shared__ uint32_t s_base;
  const uint32_t q = blockIdx.x;
  const uint32_t* __restrict__ rn = acc_n + (uint64_t)q * n_entries;
  const uint32_t* __restrict__ rs = acc_score + (uint64_t)q * n_entries;
  const uint32_t per = (n_entries + 255) / 256;
  const uint32_t lo = min(n_entries, threadIdx.x * per), hi = min(n_entries, lo + per);
  uint32_t c = 0;
  for (uint32_t e = lo; e < hi; ++e) c += rn[e] != 0;
  uint32_t tot;
  uint32_t off = blk_excl(c, ws, &tot);
  if (threadIdx.x == 0) {
    s_base = tot ? atomicAdd(total, tot) : 0;
    q_base[q] = s_base;
    q_cnt[q] = tot;
  }
  __syncthreads();
  off += s_base;
  for (uint32_t e = lo; e < hi; ++e)
    if (rn[e] != 0) {
      if (off < cap) {
        CandDev r;
        r.query = q;
        r.entry = (int32_t)(entry_base + e);
        r.score = (float)rs[e];
        r.n_matches = (int32_t)rn[e];
        cands[off] = r;
      }
      ++off;
    }
}

// get_matches for one signature: gather (global id) of every bucket member, table by table
__global__ void __launch_bounds__(256) search_gather_kernel(SearchParams p, uint32_t* __restrict__ out, uint32_t cap,
                                                            uint32_t* __restrict__ table_counts) {
  __shared__ Probe pr;
  const int tid = threadIdx.x;
  const uint32_t* __restrict__ qwords = reinterpret_cast<const uint32_t*>(p.query_sigs);
  if (tid < kTables) {
    const uint32_t slot = bucket_key(qwords[tid]) % p.size;
    MN_CHECK(slot < p.size);
    const uint32_t b = p.start[(uint64_t)tid * p.size + slot], e = p.start[(uint64_t)tid * p.size + slot + 1];
    MN_CHECK(b <= e && (uint64_t)e <= (uint64_t)kTables * p.n_sigs);
    pr.begin[tid] = b;
    pr.prefix[tid + 1] = e - b;
    table_counts[tid] = e - b;
  }
  if (tid == 0) pr.prefix[0] = 0;
  __syncthreads();
  if (tid == 0)
    for (int k = 0; k < kTables; ++k) pr.prefix[k + 1] += pr.prefix[k];
  __syncthreads();
  const uint32_t L = pr.prefix[kTables];
  for (uint32_t i = tid; i < L && i < cap; i += 256) {
    int k = 0;
    while (k < kTables - 1 && pr.prefix[k + 1] <= i) ++k;
    out[i] = p.ids[pr.begin[k] + (i - pr.prefix[k])];
  }
}

// query of every query signature, once per batch: the probe kernels read one word instead of bisecting query_start
__global__ void __launch_bounds__(256) qs_query_kernel(const uint32_t* __restrict__ query_start, uint32_t n_queries,
                                                       uint32_t n_query_sigs, uint32_t* __restrict__ qs_query) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_query_sigs) qs_query[i] = find_segment<uint32_t>(query_start, n_queries, i);
}
void launch_qs_query(const uint32_t* query_start, uint32_t n_queries, uint32_t n_query_sigs, uint32_t* qs_query, cudaStream_t st) {
  if (n_query_sigs) qs_query_kernel<<<(n_query_sigs + 255) / 256, 256, 0, st>>>(query_start, n_queries, n_query_sigs, qs_query);
}

void launch_search(const SearchParams& p, uint32_t n_query_sigs, cudaStream_t st) {
  if (!n_query_sigs) return;
  static unsigned long long configured = 0;   // the attribute is per function and device: set once per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((configured >> (dev & 63)) & 1ull)) {
    cudaFuncSetAttribute(search_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)search_smem_bytes());
    configured |= 1ull << (dev & 63);
  }
  search_probe_kernel<<<n_query_sigs, kSearchThreads, search_smem_bytes(), st>>>(p);
}
void launch_search_compact(const uint32_t* acc_score, const uint32_t* acc_n, uint32_t n_queries, uint32_t n_entries,
                           uint32_t entry_base, CandDev* cands, uint32_t cap, uint32_t* total, uint32_t* q_base,
                           uint32_t* q_cnt, cudaStream_t st) {
  if (n_queries) search_compact_kernel<<<n_queries, 256, 0, st>>>(acc_score, acc_n, n_entries, entry_base, cands, cap,
                                                                  total, q_base, q_cnt);
}
void launch_search_gather(const SearchParams& p, uint32_t* out, uint32_t cap, uint32_t* table_counts, cudaStream_t st) {
  search_gather_kernel<<<1, 256, 0, st>>>(p, out, cap, table_counts);
}

}  // namespace mnemo
