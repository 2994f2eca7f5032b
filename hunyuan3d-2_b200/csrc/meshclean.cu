// Mesh clean-up on the device, before the mesh crosses PCIe: what the step right after the hot path does on the host —
// export_to_trimesh (reference hy3dgen/shapegen/pipelines.py:95-110): `mesh_f[:, ::-1]` (winding flip) followed by
// trimesh.Trimesh(v, f) whose default processing drops non-finite vertices together with every face that references one,
// and keeps only referenced vertices, re-indexing the faces.  The sparse volume decoders leave NaN at
// unvisited voxels, so their meshes carry NaN vertices along the band's rim; here they never leave the GPU.
//
//   k_face_flags   per face: all three vertices finite?  -> keep bit + marks its vertices referenced      (12 F read)
//   k_keep_counts  per 1024-element block: kept vertices (finite & referenced) / kept faces              (V + F bytes)
//   k_scan_counts  exclusive scan of the block counts (one block), totals
//   k_emit         order-preserving compaction: vertices copied, faces re-indexed and (optionally) flipped (12 V + 12 F in / out)
// Integer / byte work, HBM-bound: algorithmic bytes 24 V + 24 F.  (trimesh also merges vertices closer than 1e-8; the
// marching-cubes output is welded already, see INTEGRATION.md for the one corner case.)
#include "common.cuh"

namespace {

constexpr int CL_BLOCK = 1024;

__device__ __forceinline__ bool finite3(const float* __restrict__ v, long long i) {
  const float x = v[3 * i], y = v[3 * i + 1], z = v[3 * i + 2];
  return (fabsf(x) <= 3.402823466e38f) && (fabsf(y) <= 3.402823466e38f) && (fabsf(z) <= 3.402823466e38f);   // false for NaN / inf
}

__global__ void __launch_bounds__(256) k_face_flags(const float* __restrict__ verts, long long nV, const int32_t* __restrict__ faces,
                                                     long long nF, uint8_t* __restrict__ fkeep, uint8_t* __restrict__ vref) {
  const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nF) return;
  const int a = faces[3 * f], b = faces[3 * f + 1], c = faces[3 * f + 2];
  const bool ok = (unsigned)a < (unsigned long long)nV && (unsigned)b < (unsigned long long)nV && (unsigned)c < (unsigned long long)nV &&
                  finite3(verts, a) && finite3(verts, b) && finite3(verts, c);
  fkeep[f] = ok;
  if (ok) { vref[a] = 1; vref[b] = 1; vref[c] = 1; }          // same value from every writer: a benign race
}

// blockIdx.y = 0: vertices (keep = referenced; a referenced vertex is finite by construction), 1: faces
__global__ void __launch_bounds__(256) k_keep_counts(const uint8_t* __restrict__ vref, long long nV, const uint8_t* __restrict__ fkeep,
                                                      long long nF, int* __restrict__ cntV, int* __restrict__ cntF) {
  const uint8_t* src = blockIdx.y ? fkeep : vref;
  const long long n = blockIdx.y ? nF : nV;
  const long long base = (long long)blockIdx.x * CL_BLOCK;
  if (base >= n) return;
  int c = 0;
  for (int t = threadIdx.x; t < CL_BLOCK; t += 256) c += (base + t < n) ? src[base + t] : 0;
  __shared__ int red[8];
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < 8; ++w) s += red[w];
    (blockIdx.y ? cntF : cntV)[blockIdx.x] = s;
  }
}

// exclusive scan of n counts by one block per array (blockIdx.x selects it); out[n] = total
__global__ void __launch_bounds__(1024) k_scan_counts(const int* __restrict__ in0, int n0, long long* __restrict__ out0,
                                                       const int* __restrict__ in1, int n1, long long* __restrict__ out1) {
  const int* in = blockIdx.x ? in1 : in0;
  long long* out = blockIdx.x ? out1 : out0;
  const int n = blockIdx.x ? n1 : n0;
  __shared__ long long part[1024];
  const int tid = threadIdx.x, per = (n + 1023) / 1024, lo = tid * per, hi = min(lo + per, n);
  long long s = 0;
  for (int i = lo; i < hi; ++i) s += in[i];
  part[tid] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    long long v = tid >= off ? part[tid - off] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  long long run = tid ? part[tid - 1] : 0;
  for (int i = lo; i < hi; ++i) { out[i] = run; run += in[i]; }
  if (tid == 1023) out[n] = part[1023];
}

// block-local exclusive scan of 1024 keep flags (4 per thread) + the block's offset -> new position of every kept element
__device__ __forceinline__ void block_positions(const uint8_t* __restrict__ keep, long long base, long long n, long long blockoff,
                                                bool k[4], long long pos[4]) {
  __shared__ int wsum[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int c = 0;
#pragma unroll
  for (int u = 0; u < 4; ++u) { const long long e = base + threadIdx.x * 4 + u; k[u] = e < n && keep[e]; c += k[u]; }
  int incl = c;
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  long long p = blockoff + incl - c;
  for (int w = 0; w < warp; ++w) p += wsum[w];
#pragma unroll
  for (int u = 0; u < 4; ++u) { pos[u] = p; p += k[u]; }
  __syncthreads();
}

__global__ void __launch_bounds__(256) k_emit_verts(const float* __restrict__ verts, long long nV, const uint8_t* __restrict__ vref,
                                                     const long long* __restrict__ offV, float* __restrict__ out, int32_t* __restrict__ remap) {
  const long long base = (long long)blockIdx.x * CL_BLOCK;
  bool k[4]; long long pos[4];
  block_positions(vref, base, nV, offV[blockIdx.x], k, pos);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long e = base + threadIdx.x * 4 + u;
    if (e >= nV) continue;
    remap[e] = k[u] ? (int32_t)pos[u] : -1;
    if (k[u]) { out[3 * pos[u]] = verts[3 * e]; out[3 * pos[u] + 1] = verts[3 * e + 1]; out[3 * pos[u] + 2] = verts[3 * e + 2]; }
  }
}

__global__ void __launch_bounds__(256) k_emit_faces(const int32_t* __restrict__ faces, long long nF, const uint8_t* __restrict__ fkeep,
                                                     const long long* __restrict__ offF, const int32_t* __restrict__ remap, int flip,
                                                     int32_t* __restrict__ out) {
  const long long base = (long long)blockIdx.x * CL_BLOCK;
  bool k[4]; long long pos[4];
  block_positions(fkeep, base, nF, offF[blockIdx.x], k, pos);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long e = base + threadIdx.x * 4 + u;
    if (e >= nF || !k[u]) continue;
    const int a = remap[faces[3 * e]], b = remap[faces[3 * e + 1]], c = remap[faces[3 * e + 2]];
    out[3 * pos[u]] = flip ? c : a; out[3 * pos[u] + 1] = b; out[3 * pos[u] + 2] = flip ? a : c;
  }
}

}  // namespace

extern "C" int hy3d_mesh_clean(hy3d_ctx* ctx, const float* d_verts, int64_t nV, const int32_t* d_faces, int64_t nF, int32_t flip_winding,
                               float* d_verts_out, int32_t* d_faces_out, int64_t* h_nv_out, int64_t* h_nf_out) {
  if (!ctx || nV < 0 || nF < 0 || !h_nv_out || !h_nf_out) return HY3D_ERR_ARG;
  *h_nv_out = 0; *h_nf_out = 0;
  if (nV == 0 || nF == 0) return HY3D_OK;
  if (!d_verts || !d_faces || !d_verts_out || !d_faces_out) return HY3D_ERR_ARG;
  if (nV > 2147483647LL) return hy3d_fail(ctx, HY3D_ERR_UNSUPPORTED, "vertex ids exceed int32");
  HY3D_CUDA(ctx, cudaSetDevice(ctx->device));
  const int nbV = (int)ceil_div64(nV, CL_BLOCK), nbF = (int)ceil_div64(nF, CL_BLOCK);
  const size_t aV = ((size_t)nV + 255) / 256 * 256, aF = ((size_t)nF + 255) / 256 * 256;
  const size_t bytes = aV + aF + aV * 4 + ((size_t)nbV + nbF + 128) * 4 + ((size_t)nbV + nbF + 66) * 8;
  HY3D_CUDA(ctx, ctx->scratch.reserve(bytes));
  uint8_t* vref = ctx->scratch.as<uint8_t>();
  uint8_t* fkeep = vref + aV;
  int32_t* remap = reinterpret_cast<int32_t*>(fkeep + aF);
  int* cntV = remap + aV;
  int* cntF = cntV + (nbV + 63) / 64 * 64;
  long long* offV = reinterpret_cast<long long*>(cntF + (nbF + 63) / 64 * 64);      // 256-byte multiples keep the 8-byte alignment
  long long* offF = offV + (nbV + 1 + 31) / 32 * 32;
  HY3D_CUDA(ctx, cudaMemsetAsync(vref, 0, aV, ctx->stream));
  HY3D_PROF(ctx, FAM_MC_EMIT);
  k_face_flags<<<(unsigned)ceil_div64(nF, 256), 256, 0, ctx->stream>>>(d_verts, nV, d_faces, nF, fkeep, vref);
  HY3D_LAUNCH_CHECK(ctx);
  HY3D_PROF(ctx, FAM_MC_EMIT);
  k_keep_counts<<<dim3((unsigned)(nbV > nbF ? nbV : nbF), 2), 256, 0, ctx->stream>>>(vref, nV, fkeep, nF, cntV, cntF);
  HY3D_LAUNCH_CHECK(ctx);
  HY3D_PROF(ctx, FAM_MC_SCAN);
  k_scan_counts<<<2, 1024, 0, ctx->stream>>>(cntV, nbV, offV, cntF, nbF, offF);
  HY3D_LAUNCH_CHECK(ctx);
  HY3D_PROF(ctx, FAM_MC_EMIT);
  k_emit_verts<<<nbV, 256, 0, ctx->stream>>>(d_verts, nV, vref, offV, d_verts_out, remap);
  HY3D_LAUNCH_CHECK(ctx);
  HY3D_PROF(ctx, FAM_MC_EMIT);
  k_emit_faces<<<nbF, 256, 0, ctx->stream>>>(d_faces, nF, fkeep, offF, remap, flip_winding ? 1 : 0, d_faces_out);
  HY3D_LAUNCH_CHECK(ctx);
  long long* pl = reinterpret_cast<long long*>(ctx->pinned);
  HY3D_CUDA(ctx, cudaMemcpyAsync(pl, offV + nbV, 8, cudaMemcpyDeviceToHost, ctx->stream));
  HY3D_CUDA(ctx, cudaMemcpyAsync(pl + 1, offF + nbF, 8, cudaMemcpyDeviceToHost, ctx->stream));
  HY3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *h_nv_out = pl[0]; *h_nf_out = pl[1];
  return HY3D_OK;
}
