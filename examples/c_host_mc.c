/* A plain C host over the C-ABI (include/hy3dgeo.h): marching cubes of an analytic sphere, no Python, no torch.
 * Replaces the reference call  skimage.measure.marching_cubes(grid, mc_level, method="lewiner")  +  the rescale of
 * MCSurfaceExtractor.run (hy3dgen/shapegen/models/autoencoders/surface_extractors.py:68-76).
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/c_host_mc.c hunyuan3d-2_b200/libhy3dgeo.so \
 *       -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/hunyuan3d-2_b200 -o c_host_mc && ./c_host_mc
 *
 * Exit code 0: mesh extracted and closed (F = 2V - 4); 2: no usable CUDA device (the library refuses to run: there is no
 * CPU fallback); 1: anything else. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <cuda_runtime_api.h>

#include "hy3dgeo.h"

int main(void) {
  hy3d_ctx* ctx = NULL;
  int rc = hy3d_create(0, NULL, &ctx);              /* device 0, default stream */
  if (rc != 0 || !ctx) {
    fprintf(stderr, "hy3d_create: error %d (no sm_100 device?)\n", rc);
    return 2;
  }
  const int n = 65;                                  /* octree_resolution 64 -> 65^3 samples */
  const size_t total = (size_t)n * n * n;
  float* h = (float*)malloc(total * sizeof(float));
  if (!h) return 1;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      for (int k = 0; k < n; ++k) {
        const double x = i - 32.0, y = j - 32.0, z = k - 32.0;
        h[((size_t)i * n + j) * n + k] = (float)(20.3 - sqrt(x * x + y * y + z * z));   /* > 0 inside the sphere */
      }
  float* d_grid = NULL;
  if (cudaMalloc((void**)&d_grid, total * sizeof(float)) != cudaSuccess) return 1;
  cudaMemcpy(d_grid, h, total * sizeof(float), cudaMemcpyHostToDevice);
  free(h);

  int64_t nv = 0, nf = 0;
  float minmax[3];
  rc = hy3d_mc_count(ctx, d_grid, n, n, n, 0.0f, &nv, &nf, minmax);      /* the one synchronising read-back: sizes */
  if (rc != 0) { fprintf(stderr, "hy3d_mc_count: %d %s\n", rc, hy3d_last_error(ctx)); return 1; }
  float* d_verts = NULL; int32_t* d_faces = NULL;
  cudaMalloc((void**)&d_verts, (size_t)nv * 3 * sizeof(float));
  cudaMalloc((void**)&d_faces, (size_t)nf * 3 * sizeof(int32_t));
  /* vertices / grid_size * bbox_size + bbox_min (surface_extractors.py:75), bounds = 1.01 */
  const double div[3] = {n, n, n}, mul[3] = {2.02, 2.02, 2.02}, add[3] = {-1.01, -1.01, -1.01};
  rc = hy3d_mc_emit(ctx, div, mul, add, d_verts, d_faces);
  if (rc != 0) { fprintf(stderr, "hy3d_mc_emit: %d %s\n", rc, hy3d_last_error(ctx)); return 1; }
  cudaDeviceSynchronize();
  printf("sphere 65^3: V=%lld F=%lld (field min %.3f max %.3f)\n", (long long)nv, (long long)nf, minmax[0], minmax[1]);
  const int closed = nf == 2 * nv - 4;
  cudaFree(d_faces); cudaFree(d_verts); cudaFree(d_grid);
  hy3d_destroy(ctx);
  return closed ? 0 : 1;
}
