// Coarse-to-fine octree refinement: replaces extract_near_surface_volume_fn + the Conv3d
// dilations + torch.where of the reference (volume_decoders.py:29-119, :245-260, :376-391;
// restated in SURVEY App. B).  All integer/byte work, HBM/L2-bound: no float grids, no host
// round trips except the final count.
//
//   coarse grid G[n^3] --k_coarse_active--> act (u8) --[k_dilate3 if not last]-->
//   fine mask evaluated on the fly from act (the x2 up-sampling followed by 1 or 2 box
//   dilations collapses to "any active coarse voxel c with |2c - f|_inf <= r", r = 1 or 2)
//   --k_fine_ballot--> bit words + per-block counts --k_scan--> offsets
//   --k_fine_emit--> ordered flat indices (torch.where order)
#include "common.cuh"

namespace {

__device__ __forceinline__ int sgn(float v) { return (v > 0.f) - (v < 0.f); }

__global__ void k_coarse_active(const float* __restrict__ g, int n, float alpha, uint8_t* __restrict__ act) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)n * n * n;
  if (t >= total) return;
  const unsigned tu = (unsigned)t;                       // grids are < 2^31 voxels: 32-bit index arithmetic
  int k = tu % n; unsigned r = tu / n; int j = r % n; int i = r / n;
  float gv = g[t];
  float val = __fadd_rn(gv, alpha);
  bool valid = val > -9000.f;
  int s = sgn(val);
  bool diff = false;
  const long long sn[3] = {(long long)n * n, n, 1};
  const int idx[3] = {i, j, k};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int d = -1; d <= 1; d += 2) {
      int c = idx[a] + d;
      c = c < 0 ? 0 : (c > n - 1 ? n - 1 : c);                 // replicate padding
      float nb = __fadd_rn(g[t + (long long)(c - idx[a]) * sn[a]], alpha);
      if (!(nb > -9000.f)) nb = val;                           // invalid neighbour -> own value
      diff |= sgn(nb) != s;
    }
  }
  act[t] = (uint8_t)((diff && valid) || (fabsf(gv) < HY3D_BAND));
}

__global__ void k_dilate3(const uint8_t* __restrict__ in, int n, uint8_t* __restrict__ out) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)n * n * n;
  if (t >= total) return;
  const unsigned tu = (unsigned)t;
  int k = tu % n; unsigned r = tu / n; int j = r % n; int i = r / n;
  uint8_t v = 0;
  for (int a = max(i - 1, 0); a <= min(i + 1, n - 1); ++a)
    for (int b = max(j - 1, 0); b <= min(j + 1, n - 1); ++b)
      for (int c = max(k - 1, 0); c <= min(k + 1, n - 1); ++c) v |= in[((long long)a * n + b) * n + c];
  out[t] = v;
}

__device__ __forceinline__ bool fine_active(const uint8_t* __restrict__ act, int n, int reach, int fi, int fj, int fk) {
  // coarse c with |2c - f| <= reach  <=>  ceil((f-reach)/2) <= c <= floor((f+reach)/2)
  int lo[3], hi[3];
  const int f[3] = {fi, fj, fk};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    int l = f[a] - reach, h = f[a] + reach;
    lo[a] = l <= 0 ? 0 : (l + 1) >> 1;
    hi[a] = h >> 1;
    if (hi[a] > n - 1) hi[a] = n - 1;
  }
  for (int a = lo[0]; a <= hi[0]; ++a)
    for (int b = lo[1]; b <= hi[1]; ++b)
      for (int c = lo[2]; c <= hi[2]; ++c)
        if (act[((long long)a * n + b) * n + c]) return true;
  return false;
}

constexpr int FB_WARPS = 8;
constexpr int FB_ITERS = 16;                      // 32-voxel words per warp
constexpr int FB_BLOCK = FB_WARPS * FB_ITERS * 32;  // 4096 fine voxels per block

// cand = 3x3x3 dilation of act: if cand[f >> 1] is clear no coarse voxel within reach of f is active
// (all candidates lie in [f>>1 - 1, f>>1 + 1] per axis), so most fine voxels cost one byte load.
__global__ void __launch_bounds__(FB_WARPS * 32) k_fine_ballot(const uint8_t* __restrict__ act, const uint8_t* __restrict__ cand,
                                                                int n, int nf, int reach, uint32_t* __restrict__ words,
                                                                int* __restrict__ blockcnt) {
  __shared__ int wc[FB_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total = (long long)nf * nf * nf;
  long long base = (long long)blockIdx.x * FB_BLOCK + (long long)warp * FB_ITERS * 32;
  int cnt = 0;
  for (int it = 0; it < FB_ITERS; ++it) {
    long long t = base + it * 32 + lane;
    bool on = false;
    if (t < total) {
      const unsigned tu = (unsigned)t;
      int k = tu % nf; unsigned r = tu / nf; int j = r % nf; int i = r / nf;
      if (cand[((long long)(i >> 1) * n + (j >> 1)) * n + (k >> 1)]) on = fine_active(act, n, reach, i, j, k);
    }
    uint32_t m = __ballot_sync(0xffffffffu, on);
    if (lane == 0 && base + it * 32 < total) words[(base >> 5) + it] = m;
    cnt += __popc(m);
  }
  if (lane == 0) wc[warp] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < FB_WARPS; ++w) s += wc[w];
    blockcnt[blockIdx.x] = s;
  }
}

// exclusive scan of int counts by ONE block (n up to a few hundred thousand); total -> out[n]
__global__ void __launch_bounds__(1024) k_scan_i32(const int* __restrict__ in, int n, long long* __restrict__ out) {
  __shared__ long long part[1024];
  const int tid = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int lo = tid * per, hi = min(lo + per, n);
  long long s = 0;
  for (int i = lo; i < hi; ++i) s += in[i];
  part[tid] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {                 // Hillis-Steele inclusive scan
    long long v = tid >= off ? part[tid - off] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  long long run = tid ? part[tid - 1] : 0;
  for (int i = lo; i < hi; ++i) { out[i] = run; run += in[i]; }
  if (tid == 1023) out[n] = part[1023];
}

__global__ void __launch_bounds__(FB_WARPS * 32) k_fine_emit(const uint32_t* __restrict__ words, long long nwords,
                                                              const long long* __restrict__ blockoff, int32_t* __restrict__ index,
                                                              long long cap) {
  __shared__ int wc[FB_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long w0 = (long long)blockIdx.x * (FB_BLOCK / 32) + (long long)warp * FB_ITERS;
  uint32_t my = (lane < FB_ITERS && w0 + lane < nwords) ? words[w0 + lane] : 0u;
  int c = __popc(my), incl = c;
  for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) wc[warp] = incl;
  __syncthreads();
  long long off = blockoff[blockIdx.x];
  for (int w = 0; w < warp; ++w) off += wc[w];
  int excl = incl - c;
  for (int it = 0; it < FB_ITERS; ++it) {
    uint32_t m = __shfl_sync(0xffffffffu, my, it);
    int e = __shfl_sync(0xffffffffu, excl, it);
    if ((m >> lane) & 1u) {
      long long pos = off + e + __popc(m & ((1u << lane) - 1u));
      if (pos < cap) index[pos] = (int32_t)((w0 + it) * 32 + lane);
    }
  }
}

__global__ void k_fill_f32(float* __restrict__ p, long long n, float v) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; t < n; t += stride) p[t] = v;
}

__global__ void k_scatter_f32(const int32_t* __restrict__ index, const float* __restrict__ vals, long long n, long long base,
                              float* __restrict__ grid) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int i = index[t];
  if (i < 0) return;
  float v = vals[t];
  if (v == HY3D_SENTINEL) v = __int_as_float(0x7fc00000);    // a logit equal to the sentinel is "unvisited" downstream (volume_decoders.py:275)
  grid[i - base] = v;
}

__global__ void k_sentinel_nan(float* __restrict__ p, long long n, float sentinel) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  const float nanv = __int_as_float(0x7fc00000);
  for (; t < n; t += stride) { float v = p[t]; if (v == sentinel) p[t] = nanv; }
}

}  // namespace

extern "C" {

int hy3d_refine_level(hy3d_ctx* ctx, const float* d_coarse, int32_t n, int32_t nf, float mc_level, int32_t last_level,
                      int32_t* d_index, int64_t cap, int64_t* h_count) {
  if (!ctx || !d_coarse || n < 2 || !h_count || cap < 0 || (cap > 0 && !d_index)) return HY3D_ERR_ARG;
  // the fine grid is (r+1)^3 with r = 2 r_c or 2 r_c + 1 (levels are built with r // 2, volume_decoders.py:202-208):
  // every up-sampled voxel 2c must exist (nf >= 2n - 1) and every fine voxel must have a coarse parent f >> 1 (nf <= 2n)
  if (nf < 2 * n - 1 || nf > 2 * n) return hy3d_fail(ctx, HY3D_ERR_ARG, "fine grid %d^3 does not refine a %d^3 grid (need 2n-1 or 2n)", nf, n);
  HY3D_CUDA(ctx, cudaSetDevice(ctx->device));
  const long long nc = (long long)n * n * n;
  const long long nfine = (long long)nf * nf * nf;
  if (nfine > 2147483647LL) return hy3d_fail(ctx, HY3D_ERR_UNSUPPORTED, "fine grid too large for int32 indices");
  const long long nwords = (nfine + 31) / 32;
  const int nblocks = (int)ceil_div64(nfine, FB_BLOCK);
  size_t need = (size_t)nc * 3 + 512;
  HY3D_CUDA(ctx, ctx->scratch.reserve(need));
  HY3D_CUDA(ctx, ctx->scratch2.reserve((size_t)nwords * 4 + (size_t)nblocks * 4 + (size_t)(nblocks + 1) * 8 + 1024));
  uint8_t* act = ctx->scratch.as<uint8_t>();
  uint8_t* act2 = act + ((nc + 127) / 128 * 128);
  uint8_t* cand = act2 + ((nc + 127) / 128 * 128);
  uint32_t* words = ctx->scratch2.as<uint32_t>();
  int* blockcnt = reinterpret_cast<int*>(words + ((nwords + 63) / 64 * 64));
  long long* blockoff = reinterpret_cast<long long*>(blockcnt + ((nblocks + 63) / 64 * 64));
  HY3D_PROF(ctx, FAM_OCTREE);
  k_coarse_active<<<(unsigned)ceil_div64(nc, 256), 256, 0, ctx->stream>>>(d_coarse, n, mc_level, act);
  HY3D_LAUNCH_CHECK(ctx);
  const uint8_t* mask = act;
  if (!last_level) {
    HY3D_PROF(ctx, FAM_OCTREE);
    k_dilate3<<<(unsigned)ceil_div64(nc, 256), 256, 0, ctx->stream>>>(act, n, act2);
    HY3D_LAUNCH_CHECK(ctx);
    mask = act2;
  }
  const int reach = last_level ? 2 : 1;
  HY3D_PROF(ctx, FAM_OCTREE);
  k_dilate3<<<(unsigned)ceil_div64(nc, 256), 256, 0, ctx->stream>>>(mask, n, cand);
  HY3D_LAUNCH_CHECK(ctx);
  HY3D_PROF(ctx, FAM_OCTREE);
  k_fine_ballot<<<nblocks, FB_WARPS * 32, 0, ctx->stream>>>(mask, cand, n, nf, reach, words, blockcnt);
  HY3D_LAUNCH_CHECK(ctx);
  HY3D_PROF(ctx, FAM_OCTREE);
  k_scan_i32<<<1, 1024, 0, ctx->stream>>>(blockcnt, nblocks, blockoff);
  HY3D_LAUNCH_CHECK(ctx);
  if (cap > 0) {
    HY3D_PROF(ctx, FAM_OCTREE);
    k_fine_emit<<<nblocks, FB_WARPS * 32, 0, ctx->stream>>>(words, nwords, blockoff, d_index, cap);
    HY3D_LAUNCH_CHECK(ctx);
  }
  long long* pinned = reinterpret_cast<long long*>(ctx->pinned);
  HY3D_CUDA(ctx, cudaMemcpyAsync(pinned, blockoff + nblocks, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  if (int rc = hy3d_watchdog_enqueue(ctx)) return rc;
  HY3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *h_count = pinned[0];
  return hy3d_watchdog_check(ctx);        // the coarse grid came from the tensor kernels: a barrier timeout there invalidates it
}

int hy3d_fill(hy3d_ctx* ctx, float* d_grid, int64_t n, float value) {
  if (!ctx || n < 0 || (n > 0 && !d_grid)) return HY3D_ERR_ARG;
  if (n == 0) return HY3D_OK;
  HY3D_CUDA(ctx, cudaSetDevice(ctx->device));
  int blocks = (int)(ceil_div64(n, 256) < (long long)ctx->num_sms * 16 ? ceil_div64(n, 256) : (long long)ctx->num_sms * 16);
  HY3D_PROF(ctx, FAM_OCTREE);
  k_fill_f32<<<blocks, 256, 0, ctx->stream>>>(d_grid, n, value);
  HY3D_LAUNCH_CHECK(ctx);
  return HY3D_OK;
}

int hy3d_scatter(hy3d_ctx* ctx, const int32_t* d_index, const float* d_values, int64_t n, int64_t base, float* d_grid) {
  if (!ctx || n < 0 || base < 0 || (n > 0 && (!d_index || !d_values || !d_grid))) return HY3D_ERR_ARG;
  if (n == 0) return HY3D_OK;
  HY3D_CUDA(ctx, cudaSetDevice(ctx->device));
  HY3D_PROF(ctx, FAM_OCTREE);
  k_scatter_f32<<<(unsigned)ceil_div64(n, 256), 256, 0, ctx->stream>>>(d_index, d_values, n, base, d_grid);
  HY3D_LAUNCH_CHECK(ctx);
  return HY3D_OK;
}

int hy3d_sentinel_to_nan(hy3d_ctx* ctx, float* d_grid, int64_t n, float sentinel) {
  if (!ctx || n < 0 || (n > 0 && !d_grid)) return HY3D_ERR_ARG;
  if (n == 0) return HY3D_OK;
  HY3D_CUDA(ctx, cudaSetDevice(ctx->device));
  int blocks = (int)(ceil_div64(n, 256) < (long long)ctx->num_sms * 16 ? ceil_div64(n, 256) : (long long)ctx->num_sms * 16);
  HY3D_PROF(ctx, FAM_OCTREE);
  k_sentinel_nan<<<blocks, 256, 0, ctx->stream>>>(d_grid, n, sentinel);
  HY3D_LAUNCH_CHECK(ctx);
  return HY3D_OK;
}

}  // extern "C"
