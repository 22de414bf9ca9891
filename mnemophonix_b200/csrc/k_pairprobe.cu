// k_pairprobe.cu — candidate generation through the pair index (k_search.cu: launch_pair_build), then the same deep
// check and scoring as the bucket-streaming kernel: get_matches (lsh.c:89-112) + the per-signature body of search
// (search.c:122-168) for databases large enough that streaming every member of the 25 probed buckets (L ~ 42 k ids per
// query signature at 3 M signatures) costs far more than looking at the ~1.5 k signatures that share two buckets.
//
// One CTA per query signature:
//   1. 25 lanes find the probed buckets (slot, begin, length) — L and the greatest candidate id (the group search.c
//      never scores, search.c:147-165) come from there, exactly as in the streaming kernel;
//   2. each of the 300 table pairs (i < j) is one thread's binary search: inside the slice of pair (i, j) that holds
//      bucket (i, slot_i) — same positions as in table i's CSR — the entries with slot_j equal to the query's are one
//      run [lower bound, upper bound);
//   3. a block-wide scan turns the 300 run lengths into a flat index; every thread takes entries of that flat index,
//      inserts their ids into a shared-memory hash set (a signature sharing h buckets shows up h(h-1)/2 times) and
//      lists the new ones;
//   4. warps deep-check the listed ids in groups of 32 (equal words, differing bytes; k_search.cu has the same code)
//      and add (score, 1) to the per-(query, entry) accumulators.
// The runs are de-duplicated one id range at a time: usually one range (all ids), narrower ones when a query signature
// has more distinct candidates than the shared-memory list holds (step 3 below), so there is no upper limit.
#include "common.cuh"
#include "search_internal.h"

namespace mnemo {

constexpr int kPairThreads = 320;                    // one thread per table pair (300), ten warps
constexpr int kPairs = kTables * (kTables - 1) / 2;   // 300
constexpr uint32_t kPairSetCap = 8192;                // hash set slots (power of two)
constexpr uint32_t kPairUniqCap = 4096;               // distinct candidates per query signature handled here
constexpr uint32_t kPairEmpty = 0xFFFFFFFFu;

__constant__ uint8_t c_pair_i[kPairs], c_pair_j[kPairs];

__device__ __forceinline__ uint32_t pp_bucket_key(uint32_t word_le) { return __byte_perm(word_le, 0, 0x0123); }
__device__ __forceinline__ uint32_t pp_fastmod(uint32_t x, uint64_t magic, uint32_t d) {
  return (uint32_t)__umul64hi(magic * (uint64_t)x, (uint64_t)d);
}
__device__ __forceinline__ uint32_t pp_same_slot(uint32_t key, uint32_t slot, uint32_t inv, uint32_t shift, uint32_t thr) {
  const uint32_t y = (key - slot) * inv;
  return (key >= slot && __funnelshift_r(y, y, shift) <= thr) ? 1u : 0u;
}
__device__ __forceinline__ uint32_t pp_differing_bytes(uint32_t a, uint32_t b) {
  const uint32_t x = a ^ b;
  return __popc((((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u);
}

__device__ __forceinline__ uint32_t pp_key(const unsigned long long* __restrict__ slice, uint32_t i) {
  return (uint32_t)(__ldg(slice + i) >> 32);
}
// First index in [0, len) whose slot_j is >= sj (len if none).  Slots are key % size of scrambled keys, close to uniform
// over [0, size), so the search starts at the interpolated position and gallops outwards before it bisects: ~5 probes of
// a slice that lives in DRAM instead of log2(len) ~ 13.
__device__ __forceinline__ uint32_t pp_lower_bound(const unsigned long long* __restrict__ slice, uint32_t len, uint32_t sj,
                                                   float inv_size) {
  uint32_t pos = (uint32_t)((float)sj * inv_size * (float)len);
  pos = min(pos, len - 1);
  uint32_t lo, hi;   // invariant: key(lo - 1) < sj (or lo == 0), key(hi) >= sj (or hi == len)
  if (pp_key(slice, pos) >= sj) {
    hi = pos;
    lo = 0;
    uint32_t step = 1;
    while (hi > 0) {
      const uint32_t probe = hi > step ? hi - step : 0;
      if (pp_key(slice, probe) < sj) {
        lo = probe + 1;
        break;
      }
      hi = probe;
      step <<= 1;
    }
  } else {
    lo = pos + 1;
    hi = len;
    uint32_t step = 1;
    while (lo < len) {
      const uint32_t probe = min(lo + step - 1, len - 1);
      if (pp_key(slice, probe) >= sj) {
        hi = probe;
        break;
      }
      lo = probe + 1;
      step <<= 1;
    }
  }
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (pp_key(slice, mid) < sj) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}
// length of the run of entries with slot_j == sj that starts at lb: gallop, then bisect (runs are short)
__device__ __forceinline__ uint32_t pp_run_length(const unsigned long long* __restrict__ slice, uint32_t len, uint32_t lb,
                                                  uint32_t sj) {
  if (lb >= len || pp_key(slice, lb) != sj) return 0;
  uint32_t lo = lb + 1, hi = len, step = 1;   // key(lo - 1) == sj; key(hi) > sj (or hi == len)
  while (lo < len) {
    const uint32_t probe = min(lo + step - 1, len - 1);
    if (pp_key(slice, probe) > sj) {
      hi = probe;
      break;
    }
    lo = probe + 1;
    step <<= 1;
  }
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (pp_key(slice, mid) <= sj) lo = mid + 1;
    else hi = mid;
  }
  return lo - lb;
}

__global__ void __launch_bounds__(kPairThreads, 3)
    search_pair_kernel(SearchParams p) {
  extern __shared__ __align__(16) uint32_t pp_dyn[];
  uint32_t* const set = pp_dyn;                  // [kPairSetCap]
  uint32_t* const uniq = pp_dyn + kPairSetCap;   // [kPairUniqCap]
  __shared__ uint32_t seg_pos[kPairs];          // first entry of the run, relative to the pair's array
  __shared__ uint32_t seg_pref[kPairs + 1];     // exclusive prefix of the run lengths
  __shared__ uint32_t q_begin[kTables], q_len[kTables], q_slot[kTables], q_word[kTables];
  __shared__ uint32_t s_warp_tot[kPairThreads / 32];
  __shared__ uint32_t s_nuniq, s_gmax, s_L, s_over;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t qs = blockIdx.x;
  const uint32_t n = p.n_sigs;
  const uint32_t* __restrict__ qwords = reinterpret_cast<const uint32_t*>(p.query_sigs) + (uint64_t)qs * kTables;
  const uint32_t* __restrict__ dbw = reinterpret_cast<const uint32_t*>(p.db_sigs);

  if (warp == 0) {
    uint32_t len = 0, last = 0;
    if (lane < kTables) {
      const uint32_t w = qwords[lane];
      const uint32_t slot = pp_fastmod(pp_bucket_key(w), p.size_magic, p.size);
      const uint32_t b = p.start[(uint64_t)lane * p.size + slot], e = p.start[(uint64_t)lane * p.size + slot + 1];
      len = e - b;
      q_word[lane] = w;
      q_slot[lane] = slot;
      q_begin[lane] = b - lane * n;   // position inside table `lane`: the CSR gives every table n consecutive ids
      q_len[lane] = len;
      if (len) last = __ldg(p.ids + e - 1);   // lists are sorted by id: the greatest member is the last one
    }
    const uint32_t L = __reduce_add_sync(0xffffffffu, len);
    const uint32_t gmax = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0) {
      s_L = L;
      s_gmax = gmax;
      s_nuniq = 0;
      s_over = 0;
    }
  }
  for (uint32_t i = tid; i < kPairSetCap; i += kPairThreads) set[i] = kPairEmpty;
  __syncthreads();
  const uint32_t L = s_L;
  if (L == 0) {
    if (tid == 0) p.last[qs] = mnemo_lastgroup_dev{0, 0, 0, 0};
    return;
  }
  const uint32_t gmax = s_gmax;

  // 2. one search (lower bound + run length) per table pair: thread `tid` owns pair `tid`
  const float inv_size = 1.0f / (float)p.size;
  const unsigned long long* run = nullptr;   // this thread's run of entries with (slot_i, slot_j) equal to the query's
  uint32_t run_len = 0;
  if (tid < kPairs) {
    const int i = c_pair_i[tid], j = c_pair_j[tid];
    const uint32_t len_i = q_len[i];
    if (len_i && q_len[j]) {
      const unsigned long long* __restrict__ slice = p.pairs + (uint64_t)tid * n + q_begin[i];
      const uint32_t sj = q_slot[j];
      const uint32_t lo = pp_lower_bound(slice, len_i, sj, inv_size);
      run_len = pp_run_length(slice, len_i, lo, sj);
      run = slice + lo;
    }
  }
  // block-wide exclusive scan of this round's piece lengths -> flat index over the entries (seg_pos: where the piece
  // starts, relative to p.pairs + pid * n; seg_pref: prefix); returns the total
  auto publish = [&](uint32_t pos, uint32_t cnt) -> uint32_t {
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    uint32_t base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kPairThreads / 32; ++w) {
      base += w < warp ? s_warp_tot[w] : 0u;
      total += s_warp_tot[w];
    }
    if (tid < kPairs) {
      seg_pos[tid] = pos;
      seg_pref[tid] = base + incl - cnt;
    }
    if (tid == 0) seg_pref[kPairs] = total;
    __syncthreads();
    return total;
  };
  const uint32_t run_pos = run ? (uint32_t)(run - (p.pairs + (uint64_t)tid * n)) : 0u;
  const uint32_t T = publish(run_pos, run_len);
  if (tid == 0 && p.stat_pairs) atomicAdd(p.stat_pairs, (unsigned long long)T);
  const uint32_t query = find_segment<uint32_t>(p.query_start, p.n_queries, qs);
  const uint32_t my_slot = lane < kTables ? q_slot[lane] : 0u, my_word = lane < kTables ? q_word[lane] : 0u;
  uint32_t deep = 0;

  // 3.-4. The entries are de-duplicated and deep-checked one id range at a time.  T bounds the number of distinct
  // candidates, so T <= kPairUniqCap (the usual case: T ~ 2000) is one round over all ids.  Otherwise the id space is cut
  // into ranges: a run is sorted by id, so its piece inside a range is found by two short binary searches and a round
  // only visits the entries of its range; a range whose distinct ids still do not fit the list is halved and taken
  // again — nothing of it has been scored at that point — so there is no limit on the number of candidates.
  uint32_t span = n;
  while (span > 1 && (uint64_t)T * span > (uint64_t)kPairUniqCap * n) span = (span + 1) >> 1;   // first guess: even spread
  uint32_t id_lo = 0;
  bool first = true;
  while (id_lo < n && T) {
    const uint32_t id_hi = (n - id_lo > span) ? id_lo + span : n;
    uint32_t Tr = T;
    if (!(id_lo == 0 && id_hi == n)) {
      // this round's piece of every run: ids in [id_lo, id_hi)
      uint32_t a = 0, b = 0;
      if (run_len) {
        uint32_t lo = 0, hi = run_len;
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if ((uint32_t)__ldg(run + mid) < id_lo) lo = mid + 1;
          else hi = mid;
        }
        a = lo;
        hi = run_len;
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if ((uint32_t)__ldg(run + mid) < id_hi) lo = mid + 1;
          else hi = mid;
        }
        b = lo;
      }
      Tr = publish(run_pos + a, b - a);
    }
    if (!first) {   // the set was filled by the previous round
      for (uint32_t i = tid; i < kPairSetCap; i += kPairThreads) set[i] = kPairEmpty;
      if (tid == 0) {
        s_nuniq = 0;
        s_over = 0;
      }
      __syncthreads();
    }
    first = false;
    for (uint32_t t = tid; t < Tr; t += kPairThreads) {
      // run that holds flat index t: last pid with seg_pref[pid] <= t
      uint32_t lo = 0, hi = kPairs;
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (seg_pref[mid] <= t) lo = mid;
        else hi = mid;
      }
      const uint32_t g = (uint32_t)__ldg(p.pairs + (uint64_t)lo * n + seg_pos[lo] + (t - seg_pref[lo]));
      uint32_t h = (g * 2246822519u) & (kPairSetCap - 1);
      while (true) {
        const uint32_t prev = atomicCAS(&set[h], kPairEmpty, g);
        if (prev == g) break;
        if (prev == kPairEmpty) {
          const uint32_t pos = atomicAdd(&s_nuniq, 1u);
          if (pos < kPairUniqCap) uniq[pos] = g;
          else s_over = 1;
          break;
        }
        if (*(volatile uint32_t*)&s_over) break;   // the set may be full by now: the round is taken again anyway
        h = (h + 1) & (kPairSetCap - 1);
      }
    }
    __syncthreads();
    if (s_over) {   // more distinct ids in this range than the list holds: halve it (a single id cannot overflow)
      span = (span + 1) >> 1;
      __syncthreads();   // everyone has read s_over before the next round clears it
      continue;
    }
    const uint32_t nuniq = s_nuniq;
    // deep checks, 32 signatures per warp at a time (see check_list in k_search.cu)
    for (uint32_t base = (uint32_t)warp * 32u; base < nuniq; base += kPairThreads) {
      const uint32_t cnt = min(32u, nuniq - base);
      const uint32_t my_g = uniq[base + ((uint32_t)lane < cnt ? lane : 0)];
      uint32_t w[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const uint32_t g = __shfl_sync(0xffffffffu, my_g, r);
        w[r] = lane < kTables ? __ldg(dbw + (uint64_t)g * kTables + lane) : 0u;   // lanes >= 25 hold "equal" words
      }
      uint32_t mine = 0;
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        uint32_t v = pp_differing_bytes(w[r], my_word) | (w[r] == my_word ? 0x10000u : 0u);
        v = __reduce_add_sync(0xffffffffu, v);
        if (lane == r) mine = v;
      }
      const uint32_t score = 100u - (mine & 0xFFFFu);   // search.c:35-43
      uint32_t hits = (mine >> 16) - (32u - kTables);
      uint32_t slow = __ballot_sync(0xffffffffu, (uint32_t)lane < cnt && hits < 2u);
      while (slow) {   // fewer than two equal words: the shared buckets come from key % size collisions, recount
        const int r = __ffs(slow) - 1;
        slow &= slow - 1u;
        const uint32_t g = __shfl_sync(0xffffffffu, my_g, r);
        uint32_t hh = 0;
        if (lane < kTables)
          hh = pp_same_slot(pp_bucket_key(__ldg(dbw + (uint64_t)g * kTables + lane)), my_slot, p.div_inv, p.div_shift, p.div_thr);
        const uint32_t exact = __popc(__ballot_sync(0xffffffffu, hh));
        if (lane == r) hits = exact;
      }
      if ((uint32_t)lane < cnt && my_g != gmax && hits >= 2u) {   // last group never flushed; MIN_BUCKET_MATCH_FOR_DEEP_CHECK
        ++deep;
        if (score >= 30u) {                                       // MIN_SCORE (search.c:16)
          const uint64_t a = (uint64_t)query * p.n_entries + p.sig_entry[my_g];
          atomicAdd(&p.acc_score[a], score);
          atomicAdd(&p.acc_n[a], 1u);
        }
      }
    }
    id_lo = id_hi;
    if (nuniq < kPairUniqCap / 4 && span < n) span <<= 1;   // a sparse range: widen the next one again
    __syncthreads();   // uniq and the set are rewritten by the next round
  }
  // this shard's last group: bucket hits and score of the greatest candidate
  if (warp == 0) {
    uint32_t h = 0, eq = 0;
    if (lane < kTables) {
      const uint32_t w = __ldg(dbw + (uint64_t)gmax * kTables + lane);
      h = pp_same_slot(pp_bucket_key(w), q_slot[lane], p.div_inv, p.div_shift, p.div_thr);
      eq = 4u - pp_differing_bytes(w, q_word[lane]);
    }
    const uint32_t hits = __popc(__ballot_sync(0xffffffffu, h));
    const uint32_t score = __reduce_add_sync(0xffffffffu, eq);
    if (lane == 0) p.last[qs] = mnemo_lastgroup_dev{1u, gmax, hits, score};
  }
  deep = __reduce_add_sync(0xffffffffu, deep);
  if (lane == 0) {
    if (deep) atomicAdd(p.stat_deep, (unsigned long long)deep);
    if (warp == 0) atomicAdd(p.stat_nodes, (unsigned long long)L);
  }
}

static void pair_upload_constants() {
  uint8_t pi[kPairs], pj[kPairs];
  int pid = 0;
  for (int i = 0; i < kTables; ++i)
    for (int j = i + 1; j < kTables; ++j, ++pid) {
      pi[pid] = (uint8_t)i;
      pj[pid] = (uint8_t)j;
    }
  cudaMemcpyToSymbol(c_pair_i, pi, sizeof pi);
  cudaMemcpyToSymbol(c_pair_j, pj, sizeof pj);
}

void launch_search_pairs(const SearchParams& p, uint32_t n_query_sigs, cudaStream_t st) {
  if (!n_query_sigs) return;
  const int smem = (int)(sizeof(uint32_t) * (kPairSetCap + kPairUniqCap));
  static unsigned long long configured = 0;   // per function and device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((configured >> (dev & 63)) & 1ull)) {
    cudaFuncSetAttribute(search_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    pair_upload_constants();
    configured |= 1ull << (dev & 63);
  }
  search_pair_kernel<<<n_query_sigs, kPairThreads, smem, st>>>(p);
}

}  // namespace mnemo
