// Micro-benchmark: tensor-memory read / write bandwidth per SM on B200 — tcgen05.ld / tcgen05.st 32x32b.x32 issued by
// 4 / 8 / 16 warps of one CTA per SM (each warp its own lane quadrant), `cols` 32-bit columns per pass.
// The attention kernel reads 128 x 128 fp32 scores (64 KB) per (KV tile, head) out of tensor memory; this says how fast.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../hunyuan3d-2_b200/csrc -o tmem_rate tmem_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace tc;

template <int MODE>      // 0 ld, 1 st, 2 ld + 64 MUFU.EX2 per 64 columns (the softmax mix)
__global__ void __launch_bounds__(512, 1) k(long long* out, int iters, float seed) {
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tslot;
  const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64) % 512;
  uint32_t v[32], acc = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 1) {
      HY3D_TMEM_ST32(base, v); HY3D_TMEM_ST32(base + 32, v);
      tmem_wait_st();
    } else {
      uint32_t w[32];
      HY3D_TMEM_LD32(base, v); HY3D_TMEM_LD32(base + 32, w);
      tmem_wait_ld();
      if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) { acc += __float_as_uint(ex2(__uint_as_float(v[i]) * seed)); acc += __float_as_uint(ex2(__uint_as_float(w[i]) * seed)); }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += v[i] ^ w[i];
      }
    }
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) out[1] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int MODE>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 16);
  for (int warps : {4, 8, 16}) {
    const int iters = 4096;
    k<MODE><<<148, warps * 32>>>(d, 16, 0.f);
    k<MODE><<<148, warps * 32>>>(d, iters, 0.f);
    cudaDeviceSynchronize();
    long long clk; cudaMemcpy(&clk, d, 8, cudaMemcpyDeviceToHost);
    const double bytes = (double)warps * 32 * 64 * 4 * iters;      // per SM
    printf("%-14s warps/SM %2d : %.1f clk per 64-column pass, %.1f B/clk/SM  (128x128 fp32 tile = %.0f clk)\n", name, warps,
           (double)clk / iters, bytes / clk, 65536.0 / (bytes / clk));
  }
}
int main() {
  run<0>("tcgen05.ld"); run<1>("tcgen05.st"); run<2>("ld + 64 ex2");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
