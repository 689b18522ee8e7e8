// mma_mix.cu — microbenchmark for the two-tile kernel's steady state: N=128 tcgen05.mma (A in TMEM) issued in groups that
// alternate between two accumulator / operand regions, with and without four "row" warps doing epilogue-like work
// (tcgen05.ld x2, FADD2 + cvt, tcgen05.st x2 per iteration) on the same SM, and with 4 or 8 MMAs per elected block.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I4d-facial-avatars_b200/csrc tools/mma_mix.cu -o tools/mma_mix
#include <cstdio>
#include <cuda_runtime.h>
#include "nfb_ptx.cuh"
using namespace nfb;

__global__ void __launch_bounds__(320, 1) mix_kernel(int iters, int per_block, int load, int n, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar = sb + 65536 + 8, bar2 = sb + 65536 + 16, tptr = sb + 65536;
  volatile int* done = reinterpret_cast<volatile int*>(smem + 65536 + 32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); mbar_fence_init(); *done = 0; }
  if (warp == 0) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = *reinterpret_cast<volatile uint32_t*>(smem + 65536);
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_f16(128, n);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t x = (i / (16 / per_block)) & 1;  // 16 MMAs per group, then switch stream
      const uint32_t p = tb + x * 256, q = p + 128;
      const uint64_t bdesc = umma_smem_desc_sw128(sb + (i & 3) * 16384);
      if (elect_one()) {
        for (int u = 0; u < per_block / 4; ++u) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ts(q, p + ((i + u) & 3) * 32 + ks * 8, bdesc + ks * 2, idesc, 1);
          umma_commit(bar2);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x] = t1 - t0; *done = 1; }
  } else if (warp >= 2 && load) {
    const int qd = warp & 3;
    const uint32_t tl = tb + ((uint32_t)(qd * 32) << 16);
    float acc = 0.f;
    int it = 0;
    while (!*done) {
      const uint32_t x = (it++) & 1;
      uint32_t va[32], vb[32], ha[16], hb[16];
      tmem_ld32(tl + x * 256 + 128 + ((warp >= 6) ? 64 : 0), va);
      tmem_ld32(tl + x * 256 + 128 + 32 + ((warp >= 6) ? 64 : 0), vb);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        ha[j / 2] = pack_relu_f16x2(__uint_as_float(va[j]) + 0.5f, __uint_as_float(va[j + 1]) + 0.25f);
        hb[j / 2] = pack_relu_f16x2(__uint_as_float(vb[j]) + 0.5f, __uint_as_float(vb[j + 1]) + 0.25f);
      }
      acc += __uint_as_float(ha[3]) + __uint_as_float(hb[5]);
      tmem_st16(tl + (1 - x) * 256 + 64 + ((warp >= 6) ? 32 : 0), ha);   // into a K atom the MMAs of the other stream do not read now
      tmem_st16(tl + (1 - x) * 256 + 64 + 16 + ((warp >= 6) ? 32 : 0), hb);
      tmem_wait_st();
    }
    if (acc == 123.f) sink[0] = acc;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* d; float* sink;
  cudaMalloc(&d, 148 * 8); cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  const int total_mmas = 16000;
  for (int n : {128, 256}) for (int load = 0; load < 2; ++load) for (int pb : {4, 8, 16}) {
    const int iters = total_mmas / pb;
    mix_kernel<<<148, 320, 70000>>>(iters, pb, load, n, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("N=%d  epilogue-load=%d  MMAs per elected block=%2d : %.1f cycles per MMA\n", n, load, pb, (double)mx / total_mmas);
  }
  return 0;
}
