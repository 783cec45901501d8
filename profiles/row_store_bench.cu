// Microbenchmark: how fast can the observation rows of the block-per-env kernel leave an SM?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/rsb profiles/row_store_bench.cu && /tmp/rsb
// One 256-thread CTA per env writes U = 512 rows of F = 129 floats (264,192 B per env), nothing else:
//   mode 0  as the kernel does today: a warp walks 64 contiguous rows, lanes = columns, four 128-byte
//           STG.32 segments per row + the odd last column after the walk;
//   mode 1  the same rows staged in shared memory 8 at a time (4,128 B = 258 x 16 B, every tile 16-byte
//           aligned in global memory) and sent as ONE bulk async copy per tile, double buffered;
//   mode 2  flat: thread t of the CTA writes element t, t + 256, ... of the env's block (perfectly aligned
//           128-byte segments; the upper bound for plain stores).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

constexpr int U = 512, F = 129, THREADS = 256, WARPS = 8;

template <int MODE, int TILE_ROWS = 8, int NBUF = 2>
__global__ void __launch_bounds__(THREADS, 3) rows_kernel(float* __restrict__ out, float seed) {
  extern __shared__ __align__(128) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* obase = out + (size_t)blockIdx.x * U * F;
  const int rpw = U / WARPS, r0 = warp * rpw, r1 = r0 + rpw;
  if (MODE == 0) {
    for (int u = r0; u < r1; ++u) {
      float* row = obase + (size_t)u * F;
      const float v = seed + (float)u;
      row[lane] = v;
      row[32 + lane] = v + 1.0f;
      row[64 + lane] = v + 2.0f;
      row[96 + lane] = v + 3.0f;
    }
    for (int u = r0 + lane; u < r1; u += 32) obase[(size_t)u * F + 128] = seed;
  } else if (MODE == 1) {
    float* buf = smem + (size_t)warp * NBUF * TILE_ROWS * F;  // NBUF tiles per warp
    int which = 0;
    for (int t0 = r0; t0 < r1; t0 += TILE_ROWS, which = (which + 1) % NBUF) {
      float* tile = buf + which * TILE_ROWS * F;
      if (lane == 0) {  // the tile sent NBUF rounds ago has been read
        if (NBUF == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < TILE_ROWS; ++r) {
        const float v = seed + (float)(t0 + r);
        float* row = tile + r * F;
        row[lane] = v;
        row[32 + lane] = v + 1.0f;
        row[64 + lane] = v + 2.0f;
        row[96 + lane] = v + 3.0f;
        if (lane == 0) row[128] = seed;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(obase + (size_t)t0 * F),
                     "r"((uint32_t)__cvta_generic_to_shared(tile)), "r"((uint32_t)(TILE_ROWS * F * 4))
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  } else {
    for (int e = tid; e < U * F; e += THREADS) obase[e] = seed + (float)e;
  }
}

template <int MODE, int TILE_ROWS = 8, int NBUF = 2>
float run(float* out, int E, size_t smem) {
  auto kern = rows_kernel<MODE, TILE_ROWS, NBUF>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) kern<<<E, THREADS, smem>>>(out, (float)i);
  cudaEventRecord(a);
  for (int i = 0; i < 10; ++i) kern<<<E, THREADS, smem>>>(out, (float)i);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  if (cudaGetLastError() != cudaSuccess) printf("CUDA error in mode %d\n", MODE);
  return ms / 10;
}

int main(int argc, char** argv) {
  const int E = argc > 1 ? atoi(argv[1]) : 16384;
  float* out;
  const size_t bytes = (size_t)E * U * F * 4;
  if (cudaMalloc(&out, bytes) != cudaSuccess) return printf("alloc failed\n"), 1;
  const float m0 = run<0>(out, E, 0), m1 = run<1, 8, 2>(out, E, (size_t)WARPS * 2 * 8 * F * 4), m2 = run<2>(out, E, 0);
  const float m14 = run<1, 4, 2>(out, E, (size_t)WARPS * 2 * 4 * F * 4), m141 = run<1, 4, 1>(out, E, (size_t)WARPS * 4 * F * 4),
              m181 = run<1, 8, 1>(out, E, (size_t)WARPS * 8 * F * 4);
  printf("bulk tiles: 4 rows x 2 buffers %.1f us = %.0f GB/s | 4 rows x 1 buffer %.1f us = %.0f GB/s | 8 rows x 1 buffer %.1f us = %.0f GB/s\n",
         m14 * 1e3, bytes / m14 / 1e6, m141 * 1e3, bytes / m141 / 1e6, m181 * 1e3, bytes / m181 / 1e6);
  printf("E=%d (%.2f GB per launch): direct rows %.1f us = %.0f GB/s | bulk tiles %.1f us = %.0f GB/s | flat %.1f us = %.0f GB/s\n", E,
         bytes / 1e9, m0 * 1e3, bytes / m0 / 1e6, m1 * 1e3, bytes / m1 / 1e6, m2 * 1e3, bytes / m2 / 1e6);
  // correctness of mode 1 against mode 0 on a small sample
  float *h0 = (float*)malloc((size_t)U * F * 4), *h1 = (float*)malloc((size_t)U * F * 4);
  rows_kernel<0><<<1, THREADS>>>(out, 5.0f);
  cudaMemcpy(h0, out, (size_t)U * F * 4, cudaMemcpyDeviceToHost);
  cudaMemset(out, 0, (size_t)U * F * 4);
  rows_kernel<1, 4, 1><<<1, THREADS, (size_t)WARPS * 4 * F * 4>>>(out, 5.0f);
  cudaMemcpy(h1, out, (size_t)U * F * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < U * F; ++i) bad += h0[i] != h1[i];
  printf("bulk tiles vs direct rows: %d mismatches\n", bad);
  return 0;
}
