// Micro-benchmark: issue-to-completion rate of tcgen05.mma (kind::f16, M = 128) on B200 for the
// shapes the attention kernel uses: A from shared memory (SS) or tensor memory (TS), N = 64 / 128 / 256.
// One CTA per SM, one warp issues `iters` batches of 8 x `per_commit` MMAs, each batch followed by a commit + wait.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../hunyuan3d-2_b200/csrc -o mma_rate mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace tc;

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters, int per_commit) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 ones
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tslot;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint64_t ad = make_desc_sw128(smem_u32(smem));
    const uint64_t bd = make_desc_sw128(smem_u32(smem + 16384));
    uint32_t ph = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {               // one commit + wait per `per_commit` MMAs x 8 groups
      if (elect_one()) {
        for (int g8 = 0; g8 < 8; ++g8)
          for (int k = 0; k < per_commit; ++k) {
            if (TS) mma_f16_ts(tmem, tmem + 256 + 8 * (k & 7), bd + 2 * (k & 3), idesc, 1);
            else mma_f16_ss(tmem, ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, 1);
          }
        mma_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), ph); ph ^= 1;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N, bool TS>
void run(const char* name, int per_commit) {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 512;
  const size_t smem = 1024 + 65536;
  cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<N, TS><<<148, 128, smem>>>(d, 64, per_commit);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<N, TS><<<148, 128, smem>>>(d, iters, per_commit);
  cudaEventRecord(e1);
  cudaError_t err = cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long clk; cudaMemcpy(&clk, d, 8, cudaMemcpyDeviceToHost);
  const double n = (double)iters * per_commit * 8;
  printf("%-22s per_commit %2d : %7.1f clk/MMA  (%.1f ns/MMA, nominal floor %d clk)  %.0f TFLOP/s chip  [%s]\n", name, per_commit,
         clk / n, ms * 1e6 / n, 128 * N / 256, 148.0 * n * 2.0 * 128 * N * 16 / (ms * 1e9), cudaGetErrorString(err));
  cudaFree(d);
}

int main() {
  for (int pc : {1, 4, 8}) {
    run<256, false>("SS M128 N256", pc);
    run<128, false>("SS M128 N128", pc);
    run<64, false>("SS M128 N64", pc);
    run<128, true>("TS M128 N128", pc);
    run<64, true>("TS M128 N64", pc);
  }
  return 0;
}
