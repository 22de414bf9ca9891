// Dev tool: exhaustive search for FP32-only formulations of the path's double-precision divisions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/div_probe tools/div_f32_probe.cu && /tmp/div_probe
// For every float x compares a candidate against (float)((double)x / d) and prints the number of
// mismatches together with the smallest and largest |x| that mismatch.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

struct Split { float y, h, l; };   // y = RN(1/d), d ~ h + l

__device__ __forceinline__ float cand1(float x, Split s) {   // one residual step
  const float q0 = __fmul_rn(x, s.y);
  const float r = __fmaf_rn(-s.l, q0, __fmaf_rn(-s.h, q0, x));
  return __fmaf_rn(r, s.y, q0);
}
__device__ __forceinline__ float cand2(float x, Split s) {   // two residual steps
  const float q1 = cand1(x, s);
  const float r = __fmaf_rn(-s.l, q1, __fmaf_rn(-s.h, q1, x));
  return __fmaf_rn(r, s.y, q1);
}

__global__ void probe(int mode, double d, Split s, uint32_t lo, uint64_t n, unsigned long long* out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t bits = (uint32_t)(lo + i);
  const float x = __uint_as_float(bits);
  if (!(fabsf(x) <= 3.0e38f)) return;
  float want;
  if (mode >= 10) want = __fdiv_rn(x, s.h);                                   // f32 division by the f32 constant h
  else want = __double2float_rn(__ddiv_rn((double)x, d));
  const float got = (mode % 10) == 1 ? cand1(x, s) : cand2(x, s);
  if (__float_as_uint(got) != __float_as_uint(want)) {
    if (bits == 0x80000000u) { atomicAdd(out + 3, 1ull); return; }   // -0 -> +0
    atomicAdd(out, 1ull);
    atomicMin(out + 1, (unsigned long long)(bits & 0x7fffffffu));
    atomicMax(out + 2, (unsigned long long)(bits & 0x7fffffffu));
  }
}

static void run(const char* name, int mode, double d, uint32_t lo, uint64_t n) {
  Split s;
  s.y = (float)(1.0 / d);
  s.h = (float)d;
  s.l = (float)(d - (double)s.h);
  unsigned long long* dev;
  cudaMalloc(&dev, 32);
  unsigned long long init[4] = {0, ~0ull, 0, 0};
  cudaMemcpy(dev, init, 32, cudaMemcpyHostToDevice);
  const uint64_t chunk = 1ull << 30;
  for (uint64_t off = 0; off < n; off += chunk) {
    const uint64_t m = n - off < chunk ? n - off : chunk;
    probe<<<(unsigned)((m + 255) / 256), 256>>>(mode, d, s, (uint32_t)(lo + off), m, dev);
  }
  unsigned long long r[4];
  cudaMemcpy(r, dev, 32, cudaMemcpyDeviceToHost);
  cudaFree(dev);
  float fmin, fmax;
  uint32_t bmin = (uint32_t)r[1], bmax = (uint32_t)r[2];
  memcpy(&fmin, &bmin, 4);
  memcpy(&fmax, &bmax, 4);
  printf("%-28s mode %2d y=%a h=%a l=%a: mismatches %llu (|x| from %g [0x%08x] to %g [0x%08x]), -0 cases %llu\n", name, mode,
         s.y, s.h, s.l, r[0], r[0] ? fmin : 0.f, bmin, r[0] ? fmax : 0.f, bmax, r[3]);
}

int main() {
  const double sqrt2 = 1.41421356237309504880;
  run("x/sqrt2 1-step", 1, sqrt2, 0, 1ull << 32);
  run("x/sqrt2 2-step", 2, sqrt2, 0, 1ull << 32);
  run("x/32767 1-step", 1, 32767.0, 0, 1ull << 32);
  run("x/32767 2-step", 2, 32767.0, 0, 1ull << 32);
  const float log256 = logf(256.0f);
  run("x/logf(256) f32 1-step", 11, (double)log256, 0, 1ull << 32);
  run("x/logf(256) f32 2-step", 12, (double)log256, 0, 1ull << 32);
  cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
