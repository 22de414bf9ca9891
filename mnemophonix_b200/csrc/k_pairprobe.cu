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
//   4. warps deep-check the listed ids in groups of 32, four 128-byte rows per load instruction, and add (score, 1) to
//      the per-(query, entry) accumulators (the two shared buckets search.c:153 asks for hold by construction here).
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
constexpr uint32_t kWin = 4;                          // entries per probe of a search: one aligned 32-byte sector (8 = a whole 64-byte burst: 3.9 instead of 5.1 probes per search, yet 5 % slower)

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

__global__ void __launch_bounds__(kPairThreads, 3)
    search_pair_kernel(SearchParams p) {
  extern __shared__ __align__(16) uint32_t pp_dyn[];
  uint32_t* const set = pp_dyn;                  // [kPairSetCap]
  uint32_t* const uniq = pp_dyn + kPairSetCap;   // [kPairUniqCap]
  __shared__ uint32_t seg_pos[kPairs];          // first entry of the run, relative to the pair's array
  __shared__ uint32_t seg_pref[kPairs + 1];     // exclusive prefix of the run lengths
  __shared__ uint32_t q_begin[kTables], q_len[kTables], q_slot[kTables];
  __shared__ __align__(16) uint32_t q_word[32];   // the query's 25 words, then zeros like the padding of a 128-byte row
  __shared__ uint32_t s_warp_tot[kPairThreads / 32];
  __shared__ uint32_t s_nuniq, s_gmax, s_L, s_over, s_T;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t qs = blockIdx.x;
  const uint32_t n = p.n_sigs;
  const long long t_begin = p.stat_pairs ? clock64() : 0;   // diagnostics
  const uint32_t* __restrict__ qwords = reinterpret_cast<const uint32_t*>(p.query_sigs) + (uint64_t)qs * kTables;
  const uint32_t* __restrict__ dbw = p.rows128;   // 32 words per signature, 25 used

  uint32_t last = 0;   // warp 0: greatest member of the lane's bucket; the load stays in flight until after the searches
  if (warp == 0) {
    uint32_t len = 0;
    if (lane < kTables) {
      const uint32_t w = qwords[lane];
      const uint32_t slot = pp_fastmod(pp_bucket_key(w), p.size_magic, p.size);
      MN_CHECK(slot < p.size);
      const uint32_t b = p.start[(uint64_t)lane * p.size + slot], e = p.start[(uint64_t)lane * p.size + slot + 1];
      MN_CHECK(b <= e && e <= (uint32_t)(lane + 1) * n && b >= (uint32_t)lane * n);
      len = e - b;
      q_word[lane] = w;
      q_slot[lane] = slot;
      q_begin[lane] = b - lane * n;   // position inside table `lane`: the CSR gives every table n consecutive ids
      q_len[lane] = len;
      if (len) last = __ldg(p.ids + e - 1);   // lists are sorted by id: the greatest member is the last one
    } else {
      q_word[lane] = 0u;
    }
    const uint32_t L = __reduce_add_sync(0xffffffffu, len);
    if (lane == 0) {
      s_L = L;
      s_nuniq = 0;
      s_over = 0;
      s_T = 0;
    }
  }
  for (uint32_t i = tid; i < kPairSetCap; i += kPairThreads) set[i] = kPairEmpty;
  __syncthreads();
  const uint32_t L = s_L;
  if (L == 0) {
    if (tid == 0) p.last[qs] = mnemo_lastgroup_dev{0, 0, 0, 0};
    return;
  }

  // 2. one search per table pair: thread `tid` owns pair `tid`.  Probes are whole 32-byte sectors (four entries,
  // aligned), in coordinates x = position in the pair's array + r where r makes x = 0 a sector boundary.  A sector
  // holds the lower bound itself whenever its entries are not all on one side, which ends the search early; the run
  // is then read two sectors at a time and its first entries go straight into the hash set.
  const float inv_n = 1.0f / (float)n;
  const uint32_t cap = p.pair_cap;   // kPairUniqCap unless a test shrinks it
  const uint32_t fb = p.pair_fbits, rem_mask = (1u << fb) - 1u;
  const unsigned long long* run = nullptr;   // this thread's run of entries with (slot_i, slot_j) equal to the query's
  uint32_t run_len = 0, run_pos = 0;         // run_pos: relative to the pair's array
  uint32_t rest_pos = 0, rest_len = 0;       // what the inline insertion below left for the flat pass
  uint32_t d_probes = 0;                     // diagnostics
  auto insert = [&](uint32_t g) {
    uint32_t h = (g * 2246822519u) & (kPairSetCap - 1);
    while (true) {
      const uint32_t prev = atomicCAS(&set[h], kPairEmpty, g);
      if (prev == g) break;
      if (prev == kPairEmpty) {
        const uint32_t pos = atomicAdd(&s_nuniq, 1u);
        MN_CHECK(g < n && cap <= kPairUniqCap);
        if (pos < cap) uniq[pos] = g;
        else s_over = 1;
        break;
      }
      if (*(volatile uint32_t*)&s_over) break;   // the set may be full by now: the range is taken again anyway
      h = (h + 1) & (kPairSetCap - 1);
    }
  };
  if (tid < kPairs) {
    const int i = c_pair_i[tid], j = c_pair_j[tid];
    const uint32_t len_i = q_len[i];
    if (len_i && q_len[j]) {
      const uint64_t base = (uint64_t)tid * n;
      const uint32_t r = (uint32_t)(base & (kWin - 1));
      const unsigned long long* __restrict__ P4 = p.pairs + (base - r);   // burst-aligned origin of the x coordinates
      const uint32_t Bx = q_begin[i] + r, Ex = Bx + len_i, sj = q_slot[j];
      // probe = one aligned window of kWin entries
      auto load_win = [&](uint32_t w, unsigned long long (&e)[kWin]) {
        MN_CHECK((base - r) + (uint64_t)kWin * (w + 1) <= 300ull * n + 8);   // the index is allocated with 8 entries to spare
        const ulonglong2* q2 = reinterpret_cast<const ulonglong2*>(P4) + (uint64_t)(kWin / 2) * w;
#pragma unroll
        for (int k = 0; k < kWin / 2; ++k) {
          const ulonglong2 t = __ldg(q2 + k);
          e[2 * k] = t.x;
          e[2 * k + 1] = t.y;
        }
      };
      // number of entries of window w that sort before the lower bound (entries of other buckets count as -inf / +inf)
      auto less_in = [&](uint32_t w, const unsigned long long (&e)[kWin]) -> uint32_t {
        const uint32_t x = kWin * w;
        uint32_t c = 0;
#pragma unroll
        for (int k = 0; k < kWin; ++k) c += (x + k < Bx || (x + k < Ex && ((uint32_t)(e[k] >> 32) >> fb) < sj)) ? 1u : 0u;
        return c;
      };
      uint32_t w = 0, lbx = Ex, probes = 0;   // lbx: x of the first entry >= sj (sorted search) / of the run (directory); Ex: none
      bool found = false;
      unsigned long long ent[kWin];
      if (p.pair_dir) {
        // Run directory: the slot for (slice, slot_i, slot_j) holds the position of the run's first entry in row i's
        // array, if the run exists.  Linear probing; a stranger in the chain is told by its position (outside this
        // bucket of this slice) or, inside the bucket, by the slot_j of the entry it points to.
        const uint32_t slice = (uint32_t)(j - i - 1), n_slices = (uint32_t)(kTables - 1 - i);
        const uint32_t* __restrict__ dir_i = p.pair_dir + 2ull * (uint64_t)(tid - slice) * n;   // tid - slice = pairs before row i
        const uint32_t capacity = 2u * n_slices * n;
        const uint32_t t_lo = slice * n + q_begin[i];   // bucket (i, slot_i) of this slice in the row's array
        uint32_t idx = pair_dir_slot(pair_dir_hash(slice, q_slot[i], sj), capacity);
        while (true) {
          ++probes;
          MN_CHECK(idx < capacity && probes <= capacity);
          const uint32_t v = __ldg(dir_i + idx);
          if (v == 0xFFFFFFFFu) break;
          if (v - t_lo < len_i) {
            const uint32_t x = v - slice * n + r;
            ++probes;
            load_win(x / kWin, ent);
            unsigned long long e = ent[0];
#pragma unroll
            for (int k = 1; k < (int)kWin; ++k)
              if ((uint32_t)k == (x & (kWin - 1))) e = ent[k];
            if (((uint32_t)(e >> 32) >> fb) == sj) {
              lbx = x;
              w = x / kWin;
              found = true;
              break;
            }
          }
          idx = idx + 1 == capacity ? 0 : idx + 1;
        }
      } else {
        uint32_t lo = Bx / kWin, hi = (Ex - 1) / kWin;   // lower bound >= kWin lo, <= kWin (hi + 1)
        // first guess: the rank of bucket (j, slot_j) among all signatures — table j's CSR offset, already at hand — is the
        // empirical distribution of slot_j, a far better predictor than a uniform spread when popular slots hold thousands
        w = (Bx + min((uint32_t)((float)q_begin[j] * inv_n * (float)len_i), len_i - 1)) / kWin;
        uint32_t step = 1;
        lbx = 0;
        int dir = 0;   // 0 first probe, +-1 galloping, 2 bisecting
        while (lo <= hi) {
          ++probes;
          load_win(w, ent);
          const uint32_t c = less_in(w, ent);
          if (c > 0 && c < kWin) {
            lbx = kWin * w + c;
            found = true;
            break;
          }
          if (c == kWin) {
            lo = w + 1;
            dir = dir == 0 ? 1 : dir == -1 ? 2 : dir;
          } else {
            if (w == lo) {
              lbx = kWin * w;
              found = true;
              break;
            }
            hi = w - 1;
            dir = dir == 0 ? -1 : dir == 1 ? 2 : dir;
          }
          if (lo > hi) break;
          if (dir == 1) w = min(hi, lo + step - 1);
          else if (dir == -1) w = hi - lo >= step - 1 ? hi - (step - 1) : lo;
          else w = lo + ((hi - lo) >> 1);
          step <<= 1;
        }
        if (!found) lbx = kWin * lo;
        lbx = max(lbx, Bx);
      }
      d_probes = probes;
      if (lbx < Ex) {
        // the run: entries from lbx on whose slot_j equals sj.  Its first window is usually the one the search ended on
        // (else an L1 hit), and its first entry says how many follow (saturating field: a longer run is walked in hops
        // of rem_mask entries), so the length costs no second search.  The next window is read only when the run
        // reaches into it.
        const uint32_t w0 = lbx / kWin;
        if (!found || w0 != w) load_win(w0, ent);
        const uint32_t k0 = lbx & (kWin - 1);
        unsigned long long first = ent[0];
#pragma unroll
        for (int k = 1; k < kWin; ++k)
          if ((uint32_t)k == k0) first = ent[k];
        if (((uint32_t)(first >> 32) >> fb) == sj) {
          uint32_t pos = lbx, rem = (uint32_t)(first >> 32) & rem_mask;
          while (rem == rem_mask) {
            pos += rem_mask;
            MN_CHECK(pos < Ex);
            rem = (uint32_t)(__ldg(P4 + pos) >> 32) & rem_mask;
          }
          const uint32_t total = pos - lbx + rem + 1u, x_run = lbx + total;   // the run is [lbx, x_run)
          MN_CHECK(lbx >= Bx && x_run <= Ex);
          uint32_t m = 0;
#pragma unroll
          for (int k = 0; k < kWin; ++k) {
            const uint32_t x = kWin * w0 + k;
            if (x >= lbx && x < x_run) {
              insert((uint32_t)ent[k]);
              ++m;
            }
          }
          if (x_run > kWin * (w0 + 1)) {
            load_win(w0 + 1, ent);
#pragma unroll
            for (int k = 0; k < kWin; ++k) {
              if (kWin * (w0 + 1) + k < x_run) {
                insert((uint32_t)ent[k]);
                ++m;
              }
            }
          }
          run = P4 + lbx;
          run_pos = lbx - r;
          run_len = total;
          rest_len = total - m;   // longer than the windows read: the remainder goes through the flat pass
          rest_pos = lbx + m - r;
        }
      }
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
  // flat pass: every thread takes four CONSECUTIVE entries of the published pieces (one 32-byte sector when they lie in
  // one run, as they usually do): one bisection of the prefix table per four entries, four loads in flight
  auto insert_published = [&](uint32_t Tr) {
    for (uint32_t t0 = 4u * tid; t0 < Tr; t0 += kPairThreads * 4) {
      uint32_t lo = 0, hi = kPairs;   // piece that holds flat index t0: last pid with seg_pref[pid] <= t0
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (seg_pref[mid] <= t0) lo = mid;
        else hi = mid;
      }
      uint32_t g[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t t = t0 + u;
        g[u] = kPairEmpty;
        if (t < Tr) {
          while (seg_pref[lo + 1] <= t) ++lo;   // (empty pieces are skipped; seg_pref[kPairs] = Tr > t ends the walk)
          MN_CHECK(lo < (uint32_t)kPairs && seg_pos[lo] + (t - seg_pref[lo]) < n);
          g[u] = (uint32_t)__ldg(p.pairs + (uint64_t)lo * n + seg_pos[lo] + (t - seg_pref[lo]));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (g[u] != kPairEmpty) insert(g[u]);
    }
  };
  {
    const uint32_t t_warp = __reduce_add_sync(0xffffffffu, run_len);
    if (lane == 0 && t_warp) atomicAdd(&s_T, t_warp);
  }
  if (p.stat_pairs) {   // diagnostics: one atomic per warp and counter
    const uint32_t a = __reduce_add_sync(0xffffffffu, run_len ? 1u : 0u), b = __reduce_add_sync(0xffffffffu, rest_len ? 1u : 0u);
    const uint32_t c = __reduce_add_sync(0xffffffffu, d_probes ? 1u : 0u), d = __reduce_add_sync(0xffffffffu, d_probes);
    if (lane == 0) {
      atomicAdd(p.stat_pairs + 1, (unsigned long long)a);
      atomicAdd(p.stat_pairs + 2, (unsigned long long)b);
      atomicAdd(p.stat_pairs + 3, (unsigned long long)c);
      atomicAdd(p.stat_pairs + 4, (unsigned long long)d);
    }
  }
  uint32_t gmax_word = 0;   // warp 0: the greatest candidate's row, in flight until the last-group record at the end
  if (warp == 0) {   // the greatest candidate id: the group search.c:147-165 never scores
    const uint32_t g = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0) s_gmax = g;
    MN_CHECK(g < n);
    if (lane < kTables) gmax_word = __ldg(dbw + (uint64_t)g * 32u + lane);
  }
  const long long t_searched = p.stat_pairs ? clock64() : 0;
  const uint32_t T_rest = publish(rest_pos, rest_len);   // (its barriers also order the inline insertions and s_T)
  const long long t_published = p.stat_pairs ? clock64() : 0;
  const uint32_t T = s_T;
  const uint32_t gmax = s_gmax;
  // more entries than the list holds distinct ids: go straight to the id ranges below instead of filling the set in vain
  const bool one_pass = T <= cap;   // (measured: trying up to 2x and falling back on overflow costs more than it saves)
  if (one_pass) insert_published(T_rest);
  __syncthreads();
  const long long t_inserted = p.stat_pairs ? clock64() : 0;
  if (tid == 0 && p.stat_pairs) {
    atomicAdd(p.stat_pairs, (unsigned long long)T);
    if (!one_pass) atomicAdd(p.stat_pairs + 5, 1ull);
  }
  const uint32_t query = __ldg(p.qs_query + qs);   // (a bisection of query_start here was 14 dependent loads per CTA)
  const uint4 my_q = *reinterpret_cast<const uint4*>(&q_word[4 * (lane & 7)]);   // this lane's 16 bytes of the query's row
  uint32_t deep = 0;

  // 3.-4. Usually that was all: the distinct candidates (T ~ 2000 entries, ~1500 ids) fit the list and are deep-checked
  // in one go.  When they did not, the id space is cut into ranges: a run is sorted by id, so its piece inside a range
  // is found by two short binary searches and a round only visits the entries of its range; a range whose distinct ids
  // still do not fit the list is halved and taken again — nothing of it has been scored at that point — so there is
  // no limit on the number of candidates.
  uint32_t span = n;
  bool filled = one_pass && !s_over;   // the set and the list hold the whole id range
  if (!filled)
    while (span > 1 && (uint64_t)T * span > (uint64_t)cap * n) span = (span + 1) >> 1;   // first guess: even spread
  if (!filled && span == n) span = (n + 1) >> 1;
  __syncthreads();   // everyone has read s_over before a round clears it
  uint32_t id_lo = 0, cur = 0, next_cur = 0;   // cur: this thread's first run entry with id >= id_lo
  while (id_lo < n && T) {
    const uint32_t id_hi = (n - id_lo > span) ? id_lo + span : n;
    if (!filled) {
      for (uint32_t i = tid; i < kPairSetCap; i += kPairThreads) set[i] = kPairEmpty;
      if (tid == 0) {
        s_nuniq = 0;
        s_over = 0;
      }
      // this round's piece of every run: ids in [id_lo, id_hi).  The rounds walk the id space upwards, so the piece
      // starts where the previous one ended (cur) and its end is found by galloping from there
      uint32_t a = cur, b = cur;
      if (cur < run_len) {
        if (id_hi >= n) {
          b = run_len;
        } else {
          uint32_t lo = cur, hi = run_len, step = 1;   // id(lo - 1) < id_hi (or lo == cur), id(hi) >= id_hi (or hi == run_len)
          while (lo < run_len) {
            const uint32_t probe = min(lo + step - 1, run_len - 1);
            MN_CHECK(probe < run_len);
            if ((uint32_t)__ldg(run + probe) >= id_hi) {
              hi = probe;
              break;
            }
            lo = probe + 1;
            step <<= 1;
          }
          while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            MN_CHECK(mid < run_len);
            if ((uint32_t)__ldg(run + mid) < id_hi) lo = mid + 1;
            else hi = mid;
          }
          b = lo;
        }
      }
      next_cur = b;
      const uint32_t Tr = publish(run_pos + a, b - a);   // (barriers: the set is clear before anyone inserts)
      insert_published(Tr);
      __syncthreads();
      if (s_over) {   // more distinct ids in this range than the list holds: halve it (a single id cannot overflow)
        span = (span + 1) >> 1;
        __syncthreads();   // everyone has read s_over before the next round clears it
        continue;
      }
    }
    filled = false;
    const uint32_t nuniq = s_nuniq;
    // deep checks, 32 signatures per warp at a time, four per load instruction: eight lanes read the 128-byte row of one
    // signature as 16-byte pieces and compare them with the query's piece; the four byte counts travel through one warp
    // reduction in four 8-bit fields (at most 100 differing bytes).  Every candidate of the pair index shares the two
    // buckets of its pair with the query, so MIN_BUCKET_MATCH_FOR_DEEP_CHECK (search.c:153) holds by construction and
    // only the score (search.c:35-43) is computed.
    for (uint32_t base = (uint32_t)warp * 32u; base < nuniq; base += kPairThreads) {
      const uint32_t cnt = min(32u, nuniq - base);
      const uint32_t my_g = uniq[base + ((uint32_t)lane < cnt ? lane : 0)];
      MN_CHECK(my_g < n && base + ((uint32_t)lane < cnt ? lane : 0) < cap);
      const uint32_t my_entry = __ldg(p.sig_entry + my_g);   // in flight with the rows, not after them
      MN_CHECK(my_entry < p.n_entries);
      uint4 w[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const uint32_t g = __shfl_sync(0xffffffffu, my_g, 4 * r + (lane >> 3));
        w[r] = __ldg(reinterpret_cast<const uint4*>(dbw + (uint64_t)g * 32u) + (lane & 7));
      }
      uint32_t mine = 0;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const uint32_t d = pp_differing_bytes(w[r].x, my_q.x) + pp_differing_bytes(w[r].y, my_q.y) +
                           pp_differing_bytes(w[r].z, my_q.z) + pp_differing_bytes(w[r].w, my_q.w);
        const uint32_t v = __reduce_add_sync(0xffffffffu, d << (8 * (lane >> 3)));
        if ((lane >> 2) == r) mine = v;   // lane = candidate 4 r + (lane & 3)
      }
      const uint32_t score = 100u - ((mine >> (8 * (lane & 3))) & 0xFFu);   // search.c:35-43
      if ((uint32_t)lane < cnt && my_g != gmax) {   // the last group is never flushed (search.c:147-165)
        ++deep;
        if (score >= 30u) {                          // MIN_SCORE (search.c:16)
          const uint64_t a = (uint64_t)query * p.n_entries + my_entry;
          MN_CHECK(a < (uint64_t)p.n_queries * p.n_entries);
          atomicAdd(&p.acc_score[a], score);
          atomicAdd(&p.acc_n[a], 1u);
        }
      }
    }
    id_lo = id_hi;
    cur = next_cur;
    if (nuniq < cap / 4 && span < n) span <<= 1;   // a sparse range: widen the next one again
    __syncthreads();   // uniq and the set are rewritten by the next round
  }
  // this shard's last group: bucket hits and score of the greatest candidate
  if (warp == 0) {
    uint32_t h = 0, eq = 0;
    if (lane < kTables) {
      const uint32_t w = gmax_word;
      h = pp_same_slot(pp_bucket_key(w), q_slot[lane], p.div_inv, p.div_shift, p.div_thr);
      eq = 4u - pp_differing_bytes(w, q_word[lane]);
    }
    const uint32_t hits = __popc(__ballot_sync(0xffffffffu, h));
    const uint32_t score = __reduce_add_sync(0xffffffffu, eq);
    MN_CHECK(qs < gridDim.x && query < p.n_queries);
    if (lane == 0) p.last[qs] = mnemo_lastgroup_dev{1u, gmax, hits, score};
  }
  if (p.stat_pairs && tid == 0) {
    const long long t_end = clock64();
    atomicAdd(p.stat_pairs + (one_pass ? 6 : 7), (unsigned long long)(t_end - t_begin));
    if (one_pass) {
      atomicAdd(p.stat_pairs + 8, (unsigned long long)(t_searched - t_begin));      // thread 0: setup + its own search
      atomicAdd(p.stat_pairs + 9, (unsigned long long)(t_published - t_searched));  // waiting for the slowest search
      atomicAdd(p.stat_pairs + 10, (unsigned long long)(t_inserted - t_published));
      atomicAdd(p.stat_pairs + 11, (unsigned long long)(t_end - t_inserted));
    }
  }
  deep = __reduce_add_sync(0xffffffffu, deep);
  if (lane == 0) {
    if (deep) atomicAdd(p.stat_deep, (unsigned long long)deep);
    if (warp == 0) atomicAdd(p.stat_nodes, (unsigned long long)L);
  }
}

__global__ void __launch_bounds__(256) pad_rows_kernel(const uint32_t* __restrict__ words, uint64_t n_sigs, uint32_t* __restrict__ rows) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_sigs * 32u) return;
  const uint32_t k = (uint32_t)(t & 31u);
  rows[t] = k < (uint32_t)kTables ? words[(t >> 5) * kTables + k] : 0u;
}
void launch_pad_rows(const uint8_t* sigs, uint64_t n_sigs, uint32_t* rows, cudaStream_t st) {
  if (n_sigs) pad_rows_kernel<<<(unsigned)((n_sigs * 32u + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint32_t*>(sigs), n_sigs, rows);
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
