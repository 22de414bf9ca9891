// k_compact.cu — order-preserving compaction of the kept signatures (minhash.c:45-52).
//
// build_signatures drops silent / all-255 fingerprints and packs the survivors in order.
// Here: exclusive prefix sum of the keep flags over every image of the batch (three small
// kernels), then one warp per kept image copies its 100 bytes to its final slot.  Per-track
// counts are differences of the prefix at the track boundaries.
#include "common.cuh"

namespace mnemo {

constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 8;
constexpr int kScanBlock = kScanThreads * kScanPerThread;   // 2048 images per block

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    const uint32_t s = warp_sums[w];
    base += w < warp ? s : 0u;
    tot += s;
  }
  *total = tot;
  __syncthreads();
  return base + incl - v;
}

__global__ void __launch_bounds__(kScanThreads) compact_count_kernel(const uint8_t* __restrict__ keep, uint64_t n,
                                                                     uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t ws[kScanThreads / 32];
  const uint64_t base = (uint64_t)blockIdx.x * kScanBlock + (uint64_t)threadIdx.x * kScanPerThread;
  uint32_t c = 0;
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k) c += (base + k < n) ? keep[base + k] : 0;
  uint32_t tot;
  block_exclusive_scan(c, ws, &tot);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads) compact_blockscan_kernel(uint32_t* __restrict__ block_sums,
                                                                         uint32_t n_blocks) {
  __shared__ uint32_t ws[kScanThreads / 32];
  uint32_t carry = 0;
  for (uint32_t b0 = 0; b0 < n_blocks; b0 += kScanThreads) {
    const uint32_t i = b0 + threadIdx.x;
    const uint32_t v = i < n_blocks ? block_sums[i] : 0u;
    uint32_t tot;
    const uint32_t ex = block_exclusive_scan(v, ws, &tot);
    if (i < n_blocks) block_sums[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) block_sums[n_blocks] = carry;   // grand total
}

__global__ void __launch_bounds__(kScanThreads) compact_scatter_kernel(const uint8_t* __restrict__ keep, uint64_t n,
                                                                       const uint32_t* __restrict__ block_sums,
                                                                       const uint8_t* __restrict__ sig_dense,
                                                                       uint32_t* __restrict__ pos,
                                                                       uint8_t* __restrict__ sig_out) {
  __shared__ uint32_t ws[kScanThreads / 32];
  __shared__ uint32_t spos[kScanBlock];
  __shared__ uint8_t skeep[kScanBlock];
  const uint64_t blk0 = (uint64_t)blockIdx.x * kScanBlock;
  const uint64_t base = blk0 + (uint64_t)threadIdx.x * kScanPerThread;
  uint8_t f[kScanPerThread];
  uint32_t c = 0;
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k) {
    f[k] = (base + k < n) ? keep[base + k] : 0;
    c += f[k];
  }
  uint32_t tot;
  uint32_t p = block_sums[blockIdx.x] + block_exclusive_scan(c, ws, &tot);
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k) {
    if (base + k < n) pos[base + k] = p;
    spos[threadIdx.x * kScanPerThread + k] = p;
    skeep[threadIdx.x * kScanPerThread + k] = f[k];
    p += f[k];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) pos[n] = block_sums[gridDim.x];
  __syncthreads();
  // all threads sweep the block's 2048 x 25 words (signatures are 100 bytes, 4-byte aligned): coalesced
  // reads, and coalesced writes too because kept signatures land next to each other
  const uint32_t n_here = (uint32_t)min((uint64_t)kScanBlock, n - blk0);
  const uint32_t* __restrict__ src = reinterpret_cast<const uint32_t*>(sig_dense + blk0 * kSigLen);
  uint32_t* __restrict__ dst = reinterpret_cast<uint32_t*>(sig_out);
  for (uint32_t w = threadIdx.x; w < n_here * 25u; w += kScanThreads) {
    const uint32_t i = w / 25u, k = w - i * 25u;
    if (skeep[i]) dst[(uint64_t)spos[i] * 25u + k] = src[w];
  }
}

__global__ void compact_counts_kernel(const uint32_t* __restrict__ pos, const uint64_t* __restrict__ img_prefix,
                                      uint32_t n_tracks, uint32_t* __restrict__ n_sigs) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_tracks) n_sigs[t] = pos[img_prefix[t + 1]] - pos[img_prefix[t]];
}

void launch_compact(const uint8_t* sig_dense, const uint8_t* keep, uint64_t n_images, const uint64_t* img_prefix,
                    uint32_t n_tracks, uint32_t* block_sums, uint32_t* pos, uint8_t* sig_out, uint32_t* n_sigs,
                    cudaStream_t st, int* n_launches) {
  if (n_tracks == 0) return;
  if (n_images == 0) {
    cudaMemsetAsync(n_sigs, 0, sizeof(uint32_t) * n_tracks, st);
    return;
  }
  const uint32_t n_blocks = (uint32_t)((n_images + kScanBlock - 1) / kScanBlock);
  compact_count_kernel<<<n_blocks, kScanThreads, 0, st>>>(keep, n_images, block_sums);
  compact_blockscan_kernel<<<1, kScanThreads, 0, st>>>(block_sums, n_blocks);
  compact_scatter_kernel<<<n_blocks, kScanThreads, 0, st>>>(keep, n_images, block_sums, sig_dense, pos, sig_out);
  compact_counts_kernel<<<(n_tracks + 255) / 256, 256, 0, st>>>(pos, img_prefix, n_tracks, n_sigs);
  *n_launches += 4;
}

}  // namespace mnemo
