// tmem_rate.cu — microbenchmark: tcgen05.ld / tcgen05.st throughput (32x32b.x32 / .x16) with 4 or 8 warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I4d-facial-avatars_b200/csrc tools/tmem_rate.cu -o tools/tmem_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "nfb_ptx.cuh"
using namespace nfb;

__global__ void __launch_bounds__(256, 1) ld_kernel(int iters, int mode, long long* out, uint32_t* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tptr + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) {  // loads only
    for (int i = 0; i < iters; ++i) {
      uint32_t v[32];
      tmem_ld32(tb + (i & 3) * 32, v);
      tmem_wait_ld();
      acc += v[0] + v[31];
    }
  } else if (mode == 1) {  // two loads in flight
    for (int i = 0; i < iters; i += 2) {
      uint32_t v[32], w[32];
      tmem_ld32(tb + (i & 3) * 32, v);
      tmem_ld32(tb + ((i + 1) & 3) * 32, w);
      tmem_wait_ld();
      acc += v[0] + w[31];
    }
  } else {  // stores only (x16)
    uint32_t h[16];
    for (int j = 0; j < 16; ++j) h[j] = threadIdx.x + j;
    for (int i = 0; i < iters; ++i) {
      tmem_st16(tb + (i & 7) * 16, h);
      tmem_wait_st();
    }
  }
  long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[blockIdx.x * 8 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tptr, 512); }
}

int main() {
  long long* d; uint32_t* sink;
  cudaMalloc(&d, 148 * 8 * 8); cudaMalloc(&sink, 148 * 256 * 4);
  const int iters = 4000;
  for (int threads : {128, 256}) for (int mode = 0; mode < 3; ++mode) {
    ld_kernel<<<148, threads>>>(iters, mode, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < threads / 32; ++i) mx = h[i] > mx ? h[i] : mx;
    const double bytes = (double)(threads / 32) * iters * (mode == 2 ? 2048.0 : 4096.0);
    printf("%d warps %-28s: %.1f cycles per instruction per warp, %.1f B/cycle/SM\n", threads / 32,
           mode == 0 ? "ld.32x32b.x32 (1 in flight)" : mode == 1 ? "ld.32x32b.x32 (2 in flight)" : "st.32x32b.x16", (double)mx / iters, bytes / mx);
  }
  return 0;
}
