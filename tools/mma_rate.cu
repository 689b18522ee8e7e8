// mma_rate.cu — microbenchmark: cycles per tcgen05.mma (M=128, kind::f16) for A in TMEM vs SMEM and N in {16,64,128,256},
// issued back to back by one elected lane on every SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I4d-facial-avatars_b200/csrc tools/mma_rate.cu -o tools/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "nfb_ptx.cuh"
using namespace nfb;

__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int ts, int iters, int commit_every, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar = sb + 65536 + 8, bar2 = sb + 65536 + 16, tptr = sb + 65536;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = *reinterpret_cast<volatile uint32_t*>(smem + 65536);
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_f16(128, n);
    const uint64_t bdesc = umma_smem_desc_sw128(sb);            // B: up to 256 rows x 64 K = 32 KB at offset 0
    const uint64_t adesc = umma_smem_desc_sw128(sb + 32768);    // A: 128 rows x 64 K = 16 KB
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (ts) umma_ts(tb, tb + 256 + ks * 8, bdesc + ks * 2, idesc, 1);
          else umma_ss(tb, adesc + ks * 2, bdesc + ks * 2, idesc, 1);
        }
        if (commit_every && (i % commit_every) == commit_every - 1) umma_commit(bar2);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, ph);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  const int iters = 2000;
  for (int grid : {1, 148}) for (int ts = 0; ts < 2; ++ts) for (int n : {16, 64, 128, 256}) for (int ce : {0, 1}) {
    rate_kernel<<<grid, 128, 70000>>>(n, ts, iters, ce, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("grid %3d  %s  N=%3d commit_each_unit=%d : %.1f cycles per MMA (ideal %d)\n", grid, ts ? "A=TMEM" : "A=SMEM", n, ce,
           (double)mx / (iters * 4), 128 * n / 256 < 8 ? 8 : 128 * n / 256);
  }
  return 0;
}
