// Micro-benchmark: the attention softmax instruction stream in isolation (no MMA, no barriers).
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ uint32_t sw128_off(int r, int c16) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4)); }
template <int MODE>   // 0: exp+sum+pack+sts ; 1: + max ; 2: exp only (no pack/sts) ; 3: exp+sum only
__global__ void __launch_bounds__(512, 1) k(float* out, const float* in, int tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = (warp & 3) * 32 + lane;
  uint8_t* sP = smem + (warp >> 2) * 32768;
  uint32_t sv[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) sv[i] = __float_as_uint(in[(threadIdx.x * 128 + i) & 4095]);
  float m = 1.0f, l = 0.f;
  for (int t = 0; t < tiles; ++t) {
    if (MODE == 1) {
      float mx4[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
      for (int i = 0; i < 128; i += 8)
#pragma unroll
        for (int u = 0; u < 4; ++u) mx4[u] = fmaxf(mx4[u], fmaxf(__uint_as_float(sv[i + 2 * u]), __uint_as_float(sv[i + 2 * u + 1])));
      m = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * 0.999f + 0.001f * m;
    }
    float sum4[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t pk[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      const float p0 = ex2(__uint_as_float(sv[2 * i]) - m), p1 = ex2(__uint_as_float(sv[2 * i + 1]) - m);
      if (MODE != 2) { sum4[i & 1] += p0; sum4[2 + (i & 1)] += p1; }
      if (MODE <= 1) pk[i] = pack_h2(p0, p1); else pk[i] = __float_as_uint(p0 + p1);
    }
    l += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
    if (MODE <= 1) {
#pragma unroll
      for (int c16 = 0; c16 < 16; ++c16)
        *reinterpret_cast<uint4*>(sP + (c16 >> 3) * 16384 + sw128_off(r, c16 & 7)) = make_uint4(pk[4 * c16], pk[4 * c16 + 1], pk[4 * c16 + 2], pk[4 * c16 + 3]);
    } else {
      uint32_t acc = 0;
#pragma unroll
      for (int i = 0; i < 64; ++i) acc ^= pk[i];
      if (acc == 0x12345) l += 1.f;
    }
    // make the next tile depend on this one a little (keeps the compiler honest)
#pragma unroll
    for (int i = 0; i < 128; i += 16) sv[i] = __float_as_uint(__uint_as_float(sv[i]) * 0.9999f + l * 1e-30f);
  }
  if (l == 123.f) out[threadIdx.x] = l + m;
}
template <int MODE> void run(const char* name) {
  float *d, *in; cudaMalloc(&d, 4096); cudaMalloc(&in, 4096 * 4);
  float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = -(float)(i % 37) * 0.25f; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768);
  for (int warps : {8, 16}) {
    int tiles = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, warps * 32, 4 * 32768>>>(d, in, 10);
    cudaEventRecord(e0);
    k<MODE><<<148, warps * 32, 4 * 32768>>>(d, in, tiles);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // per SM: warps/4 tile-streams of 128x128 per "tile step"
    double exps = (double)warps * 32 * 128 * tiles;
    printf("%-28s warps/SM %2d : %.3f us per tile step  -> %.1f ex2/ns/SM (%.1f/clk @1.9GHz)\n", name, warps, ms * 1e3 / tiles, exps / (ms * 1e6), exps / (ms * 1e6) / 1.9);
  }
}
int main() { run<2>("exp only"); run<3>("exp+sum"); run<0>("exp+sum+pack+sts"); run<1>("max+exp+sum+pack+sts"); return 0; }
