// FlashVDM query layout on the device: spatial bin of every active query, stable counting sort by bin,
// 128-query tile padding per bin, per-tile bin id and the strided sample list of each bin.
// Replaces reference volume_decoders.py:394-412 (bin ids of `index`, `index.sort()`, per-bin slices and
// `q[:, :, ::stride]`) plus the mini-grid regrouping of :343-356; one pass family, no host synchronisation.
//
// Layout (all int32, device): pidx[cap] flat grid indices, every bin starting on a multiple of 128, -1 = padding;
// tile_group[cap/128] bin of each tile; sidx[scap] every stride-th query of each bin (in bin order, -1 = padding);
// soff[G+1] bounds of each bin inside sidx.
#include "common.cuh"

namespace {

constexpr int NB = 216;            // 6^3 bins (query_grid_num = 6, reference :393)
constexpr int BIN_BLOCK = 256;
constexpr int BIN_PER_THREAD = 8;
constexpr int BIN_CHUNK = BIN_BLOCK * BIN_PER_THREAD;      // queries per block
constexpr int BIN_WARPS = BIN_BLOCK / 32;

struct BinGeom { int n1, n2; float cell[3], bmin[3]; };

// ---- extent of the active set: integer min/max per axis (pts = fl(fl(i*cell)+bmin) is non-decreasing in i)
__global__ void __launch_bounds__(256) k_bin_extent(const int32_t* __restrict__ index, long long nq, int n1, int n2, int* __restrict__ ext) {
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {-1, -1, -1};
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nq; e += (long long)gridDim.x * blockDim.x) {
    const int f = index[e];
    const int k = f % n2, ij = f / n2;
    const int j = ij % n1, i = ij / n1;
    lo[0] = min(lo[0], i); hi[0] = max(hi[0], i);
    lo[1] = min(lo[1], j); hi[1] = max(hi[1], j);
    lo[2] = min(lo[2], k); hi[2] = max(hi[2], k);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
    hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { atomicMin(ext + a, lo[a]); atomicMax(ext + 3 + a, hi[a]); }
  }
}

// reference :394-403 op for op in float32: q = floor((p - min) / (max - min) * (6 - 0.001)); a degenerate axis gives
// 0/0 = NaN, converted to int64 as the CPU reference does (INT64_MIN, wrapping products), the id clamped into [0, 215]
__device__ __forceinline__ int bin_of(int f, const BinGeom& g, const float* mn, const float* den) {
  const int k = f % g.n2, ij = f / g.n2;
  const int j = ij % g.n1, i = ij / g.n1;
  const int c[3] = {i, j, k};
  long long b = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float p = __fadd_rn(__fmul_rn((float)c[a], g.cell[a]), g.bmin[a]);
    const float q = floorf(__fmul_rn(__fdiv_rn(__fsub_rn(p, mn[a]), den[a]), 5.999f));
    const long long qi = (q == q) ? (long long)q : LLONG_MIN;
    b = (long long)((unsigned long long)b * 6ull + (unsigned long long)qi);
  }
  return (int)max(0LL, min((long long)(NB - 1), b));
}

__device__ __forceinline__ void bin_consts(const int* ext, const BinGeom& g, float* mn, float* den) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    mn[a] = __fadd_rn(__fmul_rn((float)ext[a], g.cell[a]), g.bmin[a]);
    const float mx = __fadd_rn(__fmul_rn((float)ext[3 + a], g.cell[a]), g.bmin[a]);
    den[a] = __fsub_rn(mx, mn[a]);
  }
}

// ---- pass 1: bin id per query (uint8) and the per-block histogram, stored bin-major [NB][nblocks]
__global__ void __launch_bounds__(BIN_BLOCK) k_bin_hist(const int32_t* __restrict__ index, long long nq, BinGeom g, const int* __restrict__ ext,
                                                       uint8_t* __restrict__ bins, int* __restrict__ blockhist, int nblocks) {
  __shared__ int hist[NB];
  __shared__ float mn[3], den[3];
  for (int b = threadIdx.x; b < NB; b += BIN_BLOCK) hist[b] = 0;
  if (threadIdx.x == 0) bin_consts(ext, g, mn, den);
  __syncthreads();
  const long long base = (long long)blockIdx.x * BIN_CHUNK;
#pragma unroll
  for (int it = 0; it < BIN_PER_THREAD; ++it) {
    const long long e = base + it * BIN_BLOCK + threadIdx.x;
    if (e < nq) {
      const int b = bin_of(index[e], g, mn, den);
      bins[e] = (uint8_t)b;
      atomicAdd(&hist[b], 1);
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < NB; b += BIN_BLOCK) blockhist[(size_t)b * nblocks + blockIdx.x] = hist[b];
}

// ---- pass 2: per bin, exclusive scan of its block counts in place; total -> counts[bin]
__global__ void __launch_bounds__(256) k_bin_scan_blocks(int* __restrict__ blockhist, int nblocks, int* __restrict__ counts) {
  __shared__ int wsum[8];
  __shared__ int carry;
  int* row = blockhist + (size_t)blockIdx.x * nblocks;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nblocks; b0 += 256) {
    const int i = b0 + threadIdx.x;
    const int v = i < nblocks ? row[i] : 0;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if ((threadIdx.x & 31) >= o) s += t; }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += wsum[w];
    const int c = carry;
    if (i < nblocks) row[i] = c + woff + s - v;
    __syncthreads();
    if (threadIdx.x == 255) carry = c + woff + s;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[blockIdx.x] = carry;
}

// ---- pass 3 (one block): padded starts, sample offsets.  meta: goff[G] | cum_pc[G] (inclusive) ; soff[G+1]
__global__ void __launch_bounds__(256) k_group_offsets(const int* __restrict__ counts, int G, int stride, int* __restrict__ goff,
                                                      int* __restrict__ cum_pc, int* __restrict__ soff) {
  __shared__ int pc[256], ns[256];
  const int t = threadIdx.x;
  const int c = t < G ? counts[t] : 0;
  pc[t] = (c + 127) / 128 * 128;
  ns[t] = (c + stride - 1) / stride;
  __syncthreads();
  if (t == 0) {
    int a = 0, s = 0;
    for (int g = 0; g < G; ++g) {
      goff[g] = a; a += pc[g]; cum_pc[g] = a;
      soff[g] = s; s += ns[g];
    }
    soff[G] = s;
  }
}

// tile -> group: number of groups whose padded end is <= tile start, clamped to G-1 (tiles past the last group are all padding)
__global__ void __launch_bounds__(256) k_tile_groups(const int* __restrict__ cum_pc, int G, long long ntiles, int* __restrict__ tile_group) {
  __shared__ int cp[256];
  if ((int)threadIdx.x < G) cp[threadIdx.x] = cum_pc[threadIdx.x];
  __syncthreads();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const long long start = t * 128;
  int lo = 0, hi = G;                        // first g with cp[g] > start
  while (lo < hi) { const int mid = (lo + hi) >> 1; if ((long long)cp[mid] > start) hi = mid; else lo = mid + 1; }
  tile_group[t] = min(lo, G - 1);
}

// ---- pass 4: stable placement.  A warp owns 256 consecutive queries (8 rounds of 32); rank inside the round by
// match.any, running per-warp counters per bin, then the warp prefix and the block's scanned offset.
__global__ void __launch_bounds__(BIN_BLOCK) k_bin_place(const int32_t* __restrict__ index, long long nq, const uint8_t* __restrict__ bins,
                                                        const int* __restrict__ blockhist, int nblocks, const int* __restrict__ goff,
                                                        const int* __restrict__ soff, int stride, int32_t* __restrict__ pidx,
                                                        int32_t* __restrict__ sidx) {
  __shared__ int wcnt[BIN_WARPS][NB];
  for (int i = threadIdx.x; i < BIN_WARPS * NB; i += BIN_BLOCK) (&wcnt[0][0])[i] = 0;
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long base = (long long)blockIdx.x * BIN_CHUNK + (long long)w * (32 * BIN_PER_THREAD);
  int bin[BIN_PER_THREAD], rank[BIN_PER_THREAD];
#pragma unroll
  for (int it = 0; it < BIN_PER_THREAD; ++it) {
    const long long e = base + it * 32 + lane;
    bin[it] = e < nq ? (int)bins[e] : 255;
    const unsigned peers = __match_any_sync(0xffffffffu, bin[it]);
    const int leader = __ffs(peers) - 1;
    int old = 0;
    if (lane == leader && bin[it] != 255) { old = wcnt[w][bin[it]]; wcnt[w][bin[it]] = old + __popc(peers); }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[it] = old + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();
  // per bin: exclusive prefix over the warps, plus the block's offset inside the bin
  for (int b = threadIdx.x; b < NB; b += BIN_BLOCK) {
    int a = blockhist[(size_t)b * nblocks + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < BIN_WARPS; ++ww) { const int c = wcnt[ww][b]; wcnt[ww][b] = a; a += c; }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < BIN_PER_THREAD; ++it) {
    if (bin[it] == 255) continue;
    const long long e = base + it * 32 + lane;
    const int within = wcnt[w][bin[it]] + rank[it];
    const int f = index[e];
    pidx[goff[bin[it]] + within] = f;
    if (within % stride == 0) sidx[soff[bin[it]] + within / stride] = f;
  }
}

// ---- level 0: m^3 mini-grids of s^3 voxels (reference :343-356); group g = (gi,gj,gk), queries in (a,b,c) order
__global__ void __launch_bounds__(256) k_minigrid_layout(int N, int m, int s, int padc, int nsamp, int stride, int32_t* __restrict__ pidx,
                                                        int32_t* __restrict__ tile_group, int32_t* __restrict__ sidx, int32_t* __restrict__ soff,
                                                        long long total, long long scap) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int G = m * m * m, s3 = s * s * s;
  if (t <= G) soff[t] = (int)(t * nsamp);
  if (t < scap && t >= (long long)G * nsamp) sidx[t] = -1;
  if (t >= total) return;
  const int g = (int)(t / padc), r = (int)(t % padc);
  if ((r & 127) == 0) tile_group[t >> 7] = g;
  int f = -1;
  if (r < s3) {
    const int gk = g % m, gj = (g / m) % m, gi = g / (m * m);
    const int c = r % s, b = (r / s) % s, a = r / (s * s);
    f = ((gi * s + a) * N + (gj * s + b)) * N + (gk * s + c);
    if (r % stride == 0) sidx[(long long)g * nsamp + r / stride] = f;
  }
  pidx[t] = f;
}

}  // namespace

extern "C" {

int hy3d_flash_layout_bins(hy3d_ctx* ctx, const int32_t* d_index, int64_t nq, int32_t n0, int32_t n1, int32_t n2,
                           const hy3d_coords* coords, int32_t stride, int32_t* d_pidx, int64_t cap, int32_t* d_tile_group,
                           int32_t* d_sidx, int64_t scap, int32_t* d_soff, int32_t* d_counts) {
  if (!ctx || !coords || nq < 0 || stride <= 0) return HY3D_ERR_ARG;
  if ((!d_index && nq > 0) || !d_pidx || !d_tile_group || !d_sidx || !d_soff) return hy3d_fail(ctx, HY3D_ERR_ARG, "null layout buffer");
  if (coords->mode != 2) return hy3d_fail(ctx, HY3D_ERR_ARG, "spatial bins need coords.mode 2 (idx*cell+bmin)");
  if ((cap % 128) != 0 || cap < (nq + (int64_t)NB * 127 + 127) / 128 * 128 || scap < nq / stride + NB)
    return hy3d_fail(ctx, HY3D_ERR_ARG, "layout capacity too small: cap >= round128(n + 216*127), scap >= n/stride + 216");
  if ((int64_t)n0 * n1 * n2 > INT32_MAX || cap > INT32_MAX) return hy3d_fail(ctx, HY3D_ERR_ARG, "grid too large for int32 indices");
  HY3D_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int nblocks = (int)((nq + BIN_CHUNK - 1) / BIN_CHUNK);
  // scratch: ext[8] | counts[256] | goff[256] | cum_pc[256] | blockhist[NB*nblocks] | bins[nq]
  const size_t ints = 8 + 256 * 3 + (size_t)NB * (size_t)(nblocks > 0 ? nblocks : 1);
  HY3D_CUDA(ctx, ctx->scratch.reserve(ints * 4 + (size_t)nq + 64));
  int* ext = ctx->scratch.as<int>();
  int* counts = ext + 8; int* goff = counts + 256; int* cum_pc = goff + 256; int* blockhist = cum_pc + 256;
  uint8_t* bins = reinterpret_cast<uint8_t*>(blockhist + (size_t)NB * (nblocks > 0 ? nblocks : 1));
  HY3D_CUDA(ctx, cudaMemsetAsync(d_pidx, 0xFF, (size_t)cap * 4, st));
  HY3D_CUDA(ctx, cudaMemsetAsync(d_sidx, 0xFF, (size_t)scap * 4, st));
  BinGeom g; g.n1 = n1; g.n2 = n2;
  for (int a = 0; a < 3; ++a) { g.cell[a] = coords->cell[a]; g.bmin[a] = coords->bmin[a]; }
  if (nq > 0) {
    const int init[8] = {INT_MAX, INT_MAX, INT_MAX, -1, -1, -1, 0, 0};
    HY3D_CUDA(ctx, cudaMemcpyAsync(ext, init, sizeof(init), cudaMemcpyHostToDevice, st));   // pageable 32 B: staged before return
    HY3D_PROF(ctx, FAM_SELECT);
    k_bin_extent<<<(unsigned)std::min<long long>(148 * 8, (nq + 255) / 256), 256, 0, st>>>(d_index, nq, n1, n2, ext);
    HY3D_LAUNCH_CHECK(ctx);
    HY3D_PROF(ctx, FAM_SELECT);
    k_bin_hist<<<nblocks, BIN_BLOCK, 0, st>>>(d_index, nq, g, ext, bins, blockhist, nblocks);
    HY3D_LAUNCH_CHECK(ctx);
    HY3D_PROF(ctx, FAM_SELECT);
    k_bin_scan_blocks<<<NB, 256, 0, st>>>(blockhist, nblocks, counts);
    HY3D_LAUNCH_CHECK(ctx);
  } else {
    HY3D_CUDA(ctx, cudaMemsetAsync(counts, 0, 256 * 4, st));
  }
  HY3D_PROF(ctx, FAM_SELECT);
  k_group_offsets<<<1, 256, 0, st>>>(counts, NB, stride, goff, cum_pc, d_soff);
  HY3D_LAUNCH_CHECK(ctx);
  const long long ntiles = cap / 128;
  HY3D_PROF(ctx, FAM_SELECT);
  k_tile_groups<<<(unsigned)((ntiles + 255) / 256), 256, 0, st>>>(cum_pc, NB, ntiles, d_tile_group);
  HY3D_LAUNCH_CHECK(ctx);
  if (nq > 0) {
    HY3D_PROF(ctx, FAM_SELECT);
    k_bin_place<<<nblocks, BIN_BLOCK, 0, st>>>(d_index, nq, bins, blockhist, nblocks, goff, d_soff, stride, d_pidx, d_sidx);
    HY3D_LAUNCH_CHECK(ctx);
  }
  if (d_counts) HY3D_CUDA(ctx, cudaMemcpyAsync(d_counts, counts, NB * 4, cudaMemcpyDeviceToDevice, st));
  return HY3D_OK;
}

int hy3d_flash_layout_minigrids(hy3d_ctx* ctx, int32_t N, int32_t mini_grid_num, int32_t stride, int32_t* d_pidx, int64_t cap,
                                int32_t* d_tile_group, int32_t* d_sidx, int64_t scap, int32_t* d_soff) {
  if (!ctx || N <= 0 || mini_grid_num <= 0 || stride <= 0) return HY3D_ERR_ARG;
  if (!d_pidx || !d_tile_group || !d_sidx || !d_soff) return hy3d_fail(ctx, HY3D_ERR_ARG, "null layout buffer");
  if (N % mini_grid_num) return hy3d_fail(ctx, HY3D_ERR_ARG, "level-0 grid must divide into mini_grid_num parts per axis");
  const int m = mini_grid_num, s = N / m, G = m * m * m;
  if (G > 255) return hy3d_fail(ctx, HY3D_ERR_UNSUPPORTED, "at most 255 mini-grids");
  const int s3 = s * s * s, padc = (s3 + 127) / 128 * 128, nsamp = (s3 + stride - 1) / stride;
  if (cap != (int64_t)G * padc || scap < (int64_t)G * nsamp) return hy3d_fail(ctx, HY3D_ERR_ARG, "cap must be G*round128(s^3), scap >= G*ceil(s^3/stride)");
  HY3D_CUDA(ctx, cudaSetDevice(ctx->device));
  const long long span = std::max<long long>(cap, std::max<long long>(scap, G + 1));
  HY3D_PROF(ctx, FAM_SELECT);
  k_minigrid_layout<<<(unsigned)((span + 255) / 256), 256, 0, ctx->stream>>>(N, m, s, padc, nsamp, stride, d_pidx, d_tile_group, d_sidx, d_soff, cap, scap);
  HY3D_LAUNCH_CHECK(ctx);
  return HY3D_OK;
}

}  // extern "C"
