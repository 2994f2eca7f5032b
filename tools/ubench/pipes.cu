// Micro-benchmark: per-SM throughput of MUFU.EX2, FADD, FFMA, F2FP, FMNMX3 on B200 vs warps per SM.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters, float seed) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
  unsigned pk = 0;
  if (OP >= 6) {                       // packed f32x2: the pairs stay in 64-bit registers for the whole loop
    unsigned long long v[8], sd;
    float b[8]; int ib[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { b[i] = a[i] * 0.5f; ib[i] = threadIdx.x + i; }
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(sd) : "f"(seed));
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (OP == 6) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v[i]) : "l"(sd));
        if (OP == 7) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(sd));
        // mixes: does scalar FP32 / integer work issue beside the packed stream (separate pipes) or share its pipe?
        if (OP == 8 || OP == 9 || OP == 11) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v[i]) : "l"(sd));
        if (OP == 8 || OP == 9) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(b[i]) : "f"(seed));
        if (OP == 9) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(b[(i + 3) & 7]) : "f"(seed));
        if (OP == 10 || OP == 11) asm volatile("mad.lo.s32 %0, %0, %1, %0;" : "+r"(ib[i]) : "r"(iters | 3));
        if (OP == 12) asm volatile("{.reg .b32 q; shl.b32 q, %0, 3; add.s32 %0, q, %1;}" : "+r"(ib[i]) : "r"(iters));
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { float x, y; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v[i])); a[i] = x + y + b[i] + __int_as_float(ib[i]); }
    iters = 0;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed));
      if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(seed));
      if (OP == 3) { unsigned r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7])); pk ^= r; }
      if (OP == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(seed));
      if (OP == 5) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[(i + 4) & 7]) : "f"(seed)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[(i + 5) & 7]) : "f"(seed)); }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 123.456f || pk == 77) out[0] = s;
}
template <int OP>
void run(const char* name, int opsPerIter) {
  float* d; cudaMalloc(&d, 4);
  for (int warps : {4, 8, 16, 32}) {
    int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148, warps * 32>>>(d, 16, 0.5f);
    cudaEventRecord(e0);
    k<OP><<<148, warps * 32>>>(d, iters, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)warps * 32 * iters * 8 * opsPerIter;       // per SM
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-12s warps/SM %2d : %.1f thread-ops/ns/SM  (= %.1f per clk at %.0f MHz nominal max)\n", name, warps, ops / (ms * 1e6), ops / (ms * 1e6) / (clk / 1e6), clk / 1e3);
  }
}
int main() {
  run<0>("MUFU.EX2", 1); run<1>("FADD", 1); run<2>("FFMA", 1); run<3>("F2FP", 1); run<4>("FMNMX3", 1); run<5>("EX2+2FADD", 3);
  run<6>("FFMA2(x2)", 2); run<7>("FADD2(x2)", 2);
  run<8>("FFMA2+FFMA", 3); run<9>("FFMA2+2FFMA", 4); run<10>("IMAD", 1); run<11>("FFMA2+IMAD", 3); run<12>("SHL+ADD", 1);
  return 0;
}
