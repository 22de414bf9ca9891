// Dev tool: issue cost of Blackwell's packed FP32 instructions (FMUL2 / FADD2 / FFMA2) against scalar ones.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/_bin/fp32x2_probe tools/fp32x2_probe.cu
// Every variant runs the same number of warp instructions per thread (kIters x kUnroll), on 16 independent
// accumulators, with 1..16 warps per SM sub-partition; prints cycles per warp instruction per sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 2048;

__device__ __forceinline__ uint64_t pk(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}

// MODE 0: scalar FMUL   1: scalar FADD   2: FMUL2   3: FADD2   4: FFMA2   5: scalar FFMA
// MODE 6: FMUL2 + FADD (1:1)   7: FMUL2 + 2 FADD   8: FADD2 + FADD2 + FADD (2:1)   9: FMUL2,FMUL2,FADD,FADD,FADD2,FADD2 (the FFT butterfly mix)
template <int MODE>
__global__ void probe(float* out, float seed, long long* cycles) {
  float a[16];
  uint64_t p[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
  const float c = seed * 0.999f;
  const uint64_t cc = pk(c, c);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = __fmul_rn(a[i], c);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = __fadd_rn(a[i], c);
    } else if (MODE == 5) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(a[i], c, c);
    } else if (MODE == 2) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
    } else if (MODE == 3) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
    } else if (MODE == 4) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(cc));
    } else if (MODE == 6) {   // 8 FMUL2 + 8 FADD
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
        a[i] = __fadd_rn(a[i], c);
      }
    } else if (MODE == 7) {   // 5 FMUL2 + 10 FADD (+1 FADD)
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
        a[2 * i] = __fadd_rn(a[2 * i], c);
        a[2 * i + 1] = __fadd_rn(a[2 * i + 1], c);
      }
      a[10] = __fadd_rn(a[10], c);
    } else if (MODE == 8) {   // 10 FADD2 + 5 FADD (+1 FADD)
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[(i + 3) & 7]) : "l"(cc));
        a[i] = __fadd_rn(a[i], c);
      }
      a[10] = __fadd_rn(a[10], c);
    } else if (MODE == 9) {   // butterfly mix: 2 FMUL2, 2 FADD, 2 FADD2, repeated (18 of 16: slightly more, reported per instruction)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
        asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i + 3]) : "l"(cc));
        a[2 * i] = __fadd_rn(a[2 * i], c);
        a[2 * i + 1] = __fadd_rn(a[2 * i + 1], c);
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[6]) : "l"(cc));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[7]) : "l"(cc));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += (float)(p[i] & 0xff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
static void run(const char* name, int per_iter) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4 * sizeof(float));
  cudaMalloc(&cyc, 8);
  printf("%-46s", name);
  for (int wps = 1; wps <= 16; wps *= 2) {   // warps per sub-partition
    const int threads = 32 * 4 * wps > 1024 ? 1024 : 32 * 4 * wps;
    const int blocks_per_sm = (32 * 4 * wps + threads - 1) / threads;
    probe<MODE><<<148 * blocks_per_sm, threads>>>(out, 1.0001f, cyc);
    probe<MODE><<<148 * blocks_per_sm, threads>>>(out, 1.0001f, cyc);
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    // cycles per warp instruction per sub-partition
    printf("  w%-2d %.2f", wps, (double)c / ((double)kIters * per_iter * wps));
  }
  printf("\n");
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("scalar FMUL", 16);
  run<1>("scalar FADD", 16);
  run<5>("scalar FFMA", 16);
  run<2>("FMUL2", 16);
  run<3>("FADD2", 16);
  run<4>("FFMA2", 16);
  run<6>("8 FMUL2 + 8 FADD", 16);
  run<7>("5 FMUL2 + 11 FADD", 16);
  run<8>("10 FADD2 + 6 FADD", 16);
  run<9>("butterfly mix 6 FMUL2 + 6 FADD + 6 FADD2", 18);
  cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
