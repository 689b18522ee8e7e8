// cvt_rate.cu — issue rate of the FP32->FP16x2 pack conversions used by the epilogue, 8 warps per SM (2 per scheduler).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int MODE>
__device__ __forceinline__ uint32_t cvt(float a, float b) {
  uint32_t r;
  if (MODE == 0) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
  if (MODE == 1) asm volatile("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
  if (MODE == 2) asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
  if (MODE == 3) asm volatile("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
  if (MODE == 4) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
  if (MODE == 5) { float t; asm volatile("add.f32 %0, %1, %2;" : "=f"(t) : "f"(a), "f"(b)); r = __float_as_uint(t); }
  if (MODE == 6) { asm volatile("max.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); }
  return r;
}
template <int MODE>
__global__ void k(int iters, long long* out, uint32_t* sink, float seed) {
  float x[16];
  for (int j = 0; j < 16; ++j) x[j] = seed + threadIdx.x + j;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) acc ^= cvt<MODE>(x[j], x[(j + 1) & 15]);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  long long* d; uint32_t* s; cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 148 * 256 * 4);
  const int iters = 2000;
  const char* names[] = {"cvt.rn.f16x2.f32", "cvt.rn.relu.f16x2.f32", "cvt.rn.satfinite.f16x2.f32", "cvt.rn.relu.satfinite.f16x2.f32", "cvt.rn.bf16x2.f32", "add.f32", "max.f32"};
  void (*ks[])(int, long long*, uint32_t*, float) = {k<0>, k<1>, k<2>, k<3>, k<4>, k<5>, k<6>};
  for (int m = 0; m < 7; ++m) {
    ks[m]<<<148, 256>>>(iters, d, s, 1.5f);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-34s: %.2f cycles per warp-instruction per scheduler (2 warps interleaved; includes 1 xor each)\n", names[m], (double)h / (iters * 16 * 2));
  }
  return 0;
}
